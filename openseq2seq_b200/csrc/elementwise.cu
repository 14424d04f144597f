// HBM-bound elementwise / reduction kernels of the Jasper path.
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace os2s {

// ---------------------------------------------------------------- weight cast + transpose
// w fp32 [K][R][C] -> w_bf16 [K][R][C] and wt_bf16 [K][C][R]; 32x32 smem tile, coalesced both ways.
__device__ __forceinline__ uint16_t to_half16(float v, int f16) {
  return f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16(v));
}
__global__ void weight_cast_transpose_kernel(const float* __restrict__ w, uint16_t* __restrict__ wb,
                                             uint16_t* __restrict__ wt, int R, int C, int f16) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* src = w + (size_t)k * R * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = src[(size_t)r * C + c];
      if (wb) wb[(size_t)k * R * C + (size_t)r * C + c] = to_half16(v, f16);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (wt) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) wt[(size_t)k * R * C + (size_t)c * R + r] = to_half16(tile[threadIdx.x][i], f16);
    }
  }
}

int weight_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int K, int C_in, int C_out,
                          cudaStream_t st, int f16) {
  if (K <= 0 || C_in <= 0 || C_out <= 0) return fail(ERR_INVALID, "weight_cast_transpose: bad shape");
  dim3 grid((C_out + 31) / 32, (C_in + 31) / 32, K), block(32, 8);
  weight_cast_transpose_kernel<<<grid, block, 0, st>>>(w, (uint16_t*)w_bf16, (uint16_t*)wt_bf16, C_in, C_out, f16);
  return check_launch("weight_cast_transpose");
}


// ================================================================== batch norm (training)
// Reference: tf.layers.batch_normalization(training=True, axis=-1, momentum, epsilon) on the conv
// output viewed as [B,T,1,C] (open_seq2seq/parts/cnns/conv_blocks.py:208-227 and :91-101).
// Statistics are taken over ALL B*T rows (masked rows included), biased variance for normalising,
// Bessel-corrected variance into the moving average (fused-BN behaviour).

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// conv outputs (the BN inputs "y") are stored as fp16, see OUT_F16 in conv_tc.cu
__device__ __forceinline__ void f16x8_to_float(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 float_to_f16x8(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// One 8-channel vector of a conv output y: fp16 (16 bytes) or fp32 (32 bytes, OS2S_CONV_F32).
template <bool YF32>
struct YVec {
  uint4 v[YF32 ? 2 : 1];
};
// vec = index of the 8-channel vector (row * (ld / 8) + column vector)
template <bool YF32>
__device__ __forceinline__ YVec<YF32> ld_yvec(const void* y, size_t vec) {
  YVec<YF32> r;
  if (YF32) {
    const uint4* b = reinterpret_cast<const uint4*>(y) + vec * 2;
    r.v[0] = __ldg(b);
    r.v[YF32 ? 1 : 0] = __ldg(b + 1);
  } else {
    r.v[0] = __ldg(reinterpret_cast<const uint4*>(y) + vec);
  }
  return r;
}
template <bool YF32>
__device__ __forceinline__ YVec<YF32> zero_yvec() {
  YVec<YF32> r;
  r.v[0] = make_uint4(0, 0, 0, 0);
  r.v[YF32 ? 1 : 0] = make_uint4(0, 0, 0, 0);
  return r;
}
template <bool YF32>
__device__ __forceinline__ void yvec_to_float(const YVec<YF32>& y, float (&f)[8]) {
  if (YF32) {
    f[0] = __uint_as_float(y.v[0].x); f[1] = __uint_as_float(y.v[0].y);
    f[2] = __uint_as_float(y.v[0].z); f[3] = __uint_as_float(y.v[0].w);
    const uint4 w = y.v[YF32 ? 1 : 0];
    f[4] = __uint_as_float(w.x); f[5] = __uint_as_float(w.y);
    f[6] = __uint_as_float(w.z); f[7] = __uint_as_float(w.w);
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&y.v[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// ------------------------------------------------------------------------------ strided copies
// blockIdx.y = copy, grid-stride over its 16-byte vectors (row-major); pitches in bytes.
__global__ void multi_copy_2d_kernel(const Copy2dTable tab) {
  const int i = blockIdx.y;
  const int rv = tab.row_vecs[i];
  const long long total = (long long)tab.rows[i] * rv;
  const char* src = tab.src[i];
  char* dst = tab.dst[i];
  const long long sp = tab.src_pitch[i], dp = tab.dst_pitch[i];
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / rv;
    const int c = (int)(v - r * rv);
    *reinterpret_cast<uint4*>(dst + r * dp + (long long)c * 16) =
        __ldg(reinterpret_cast<const uint4*>(src + r * sp + (long long)c * 16));
  }
}
int multi_copy_2d(const Copy2dTable& tab, cudaStream_t st) {
  if (tab.n <= 0) return 0;
  long long most = 0;
  for (int i = 0; i < tab.n; ++i) most = most > (long long)tab.rows[i] * tab.row_vecs[i] ? most : (long long)tab.rows[i] * tab.row_vecs[i];
  int bx = (int)((most + 256 * 4 - 1) / (256 * 4));
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  multi_copy_2d_kernel<<<dim3(bx, tab.n), 256, 0, st>>>(tab);
  return check_launch("multi_copy_2d");
}

// ---------------------------------------------------------------------------------------------
// Row tiling shared by the four BN kernels.  The [M, C] matrix is cut into row chunks, one CTA per
// chunk; inside a CTA thread t owns ONE 8-channel vector column cv = t % CV (16-byte accesses,
// a warp covers 512 contiguous bytes of a row) and walks rows r, r+RP, r+2RP, ... with RP = 512/CV
// rows in flight per pass (the CTA has exactly CV*RP threads).  Single-branch layers (43 of 53
// launches) are a separate instantiation that fits 64 registers, i.e. two 512-thread CTAs per SM.  Loops are unrolled 4 rows deep with all loads issued before use (4
// independent 16-byte loads in flight per thread and tensor), there is no integer division in any
// loop, and per-channel coefficients live in registers (single branch) or are fetched from shared
// memory once per 4 rows (dense-residual layers).
constexpr int kEwMaxThreads = 512;
constexpr int kUnroll = 4;

struct RowTile {
  int CV, RP, cv, r, row0, row1;
};
// CTA size: RP whole rows of CV vector columns, as many as fit in 512 threads (every thread is active)
static int ew_threads(int C) {
  const int CV = C >> 3;
  const int RP = kEwMaxThreads / CV > 0 ? kEwMaxThreads / CV : 1;
  return CV * RP;
}
__device__ __forceinline__ RowTile make_row_tile(int M, int C, int rows_per_block) {
  RowTile t;
  t.CV = C >> 3;
  t.RP = blockDim.x / t.CV;
  t.cv = threadIdx.x % t.CV;
  t.r = threadIdx.x / t.CV;
  t.row0 = blockIdx.x * rows_per_block;
  t.row1 = min(M, t.row0 + rows_per_block);
  return t;
}
// rows per CTA for a grid of (SM count x resident CTAs per SM): one full wave, no tail wave
static int rows_per_block_for(int M, int C, int ctas_per_sm) {
  const int RP = ew_threads(C) / (C >> 3);
  const int target = device_sm_count() * ctas_per_sm;
  int rpb = (M + target - 1) / target;
  const int min_rows = RP * kUnroll;  // at least one unrolled pass per thread
  if (rpb < min_rows) rpb = min_rows;
  return rpb;
}

// Per-channel sum and sum of squares of y [M, C] (fp16). stats = [2][C] fp32, pre-zeroed.
__global__ void __launch_bounds__(kEwMaxThreads, 2)
bn_stats_kernel(const __half* __restrict__ y, float* __restrict__ stats, int M, int C, int rows_per_block) {
  extern __shared__ float sh[];  // [2][C]
  const RowTile t = make_row_tile(M, C, rows_per_block);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  {
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const uint4* base = reinterpret_cast<const uint4*>(y) + t.cv;
    const size_t rs = (size_t)t.CV;  // row stride in uint4
    int row = t.row0 + t.r;
    for (; row + (kUnroll - 1) * t.RP < t.row1; row += kUnroll * t.RP) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u] = __ldg(base + (size_t)(row + u * t.RP) * rs);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float f[8];
        f16x8_to_float(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          q[i] += f[i] * f[i];
        }
      }
    }
    for (; row < t.row1; row += t.RP) {
      float f[8];
      f16x8_to_float(__ldg(base + (size_t)row * rs), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] += f[i] * f[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&sh[t.cv * 8 + i], s[i]);
      atomicAdd(&sh[C + t.cv * 8 + i], q[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], sh[i]);
}

int bn_stats(const void* y, float* stats, int M, int C, cudaStream_t st) {
  if (C % 8 != 0 || C > 2048 || C < 8) return fail(ERR_UNSUPPORTED, "bn_stats: C must be a multiple of 8, <= 2048");
  const int rpb = rows_per_block_for(M, C, 2);
  const int grid = (M + rpb - 1) / rpb;
  bn_stats_kernel<<<grid, ew_threads(C), 2 * C * sizeof(float), st>>>((const __half*)y, stats, M, C, rpb);
  return check_launch("bn_stats");
}

// --------------------------------------------------------------------------- forward apply
// out = mask_rows( dropout( relu( sum_j gamma_j * (y_j - mean_j) * invstd_j + beta_j ) ) )
// One launch covers the main branch plus all dense-residual branches of a block-ending layer
// (conv_blocks.py:61-168); plain layers have n_branch = 1 (conv_blocks.py:170-232).
// Block 0 additionally finalises mean / invstd (saved for backward) and the moving averages.
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  *hi = __umulhi(a, b);
  return a * b;
}
// Philox4x32-7 (Salmon et al.; 7 rounds pass BigCrush), counter = element-vector index, key = seed.
// One call yields 128 random bits = eight 16-bit uniforms, one per element of a 16-byte vector, so the
// dropout mask costs ~60 integer instructions per vector and the kernel stays HBM-bound.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo32(0xD2511F53u, ctr.x, &hi0);
    const uint32_t lo1 = mulhilo32(0xCD9E8D57u, ctr.z, &hi1);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// scale / shift of the thread's own 8 channels of branch b (batch or moving statistics); the owner
// thread (first row slot of CTA 0) also publishes mean / invstd and updates the moving averages.
__device__ __forceinline__ void bn_coef(const BnFwdParams& p, const BnBranchFwd& b, int c0, float inv_n, bool owner,
                                        float (&sc)[8], float (&sf)[8]) {
  const int C = p.C;
  float m[8], v[8], g[8], be[8];
  const float* src_m = p.use_moving ? b.moving : b.stats;
  *reinterpret_cast<float4*>(&m[0]) = __ldg(reinterpret_cast<const float4*>(src_m + c0));
  *reinterpret_cast<float4*>(&m[4]) = __ldg(reinterpret_cast<const float4*>(src_m + c0) + 1);
  const int sq = p.use_moving ? C : b.stats_ld;   // second row: moving variance / sum of squares
  *reinterpret_cast<float4*>(&v[0]) = __ldg(reinterpret_cast<const float4*>(src_m + sq + c0));
  *reinterpret_cast<float4*>(&v[4]) = __ldg(reinterpret_cast<const float4*>(src_m + sq + c0) + 1);
  *reinterpret_cast<float4*>(&g[0]) = __ldg(reinterpret_cast<const float4*>(b.gamma + c0));
  *reinterpret_cast<float4*>(&g[4]) = __ldg(reinterpret_cast<const float4*>(b.gamma + c0) + 1);
  *reinterpret_cast<float4*>(&be[0]) = __ldg(reinterpret_cast<const float4*>(b.beta + c0));
  *reinterpret_cast<float4*>(&be[4]) = __ldg(reinterpret_cast<const float4*>(b.beta + c0) + 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // training: batch statistics; inference (use_moving): the moving averages
    const float mean = p.use_moving ? m[i] : m[i] * inv_n;
    const float var = p.use_moving ? v[i] : fmaxf(v[i] * inv_n - mean * mean, 0.f);
    const float invstd = rsqrtf(var + p.eps);
    sc[i] = g[i] * invstd;
    sf[i] = be[i] - mean * sc[i];
    if (owner && !p.use_moving) {
      b.mean_invstd[c0 + i] = mean;
      b.mean_invstd[C + c0 + i] = invstd;
      if (b.moving) {
        const float n = 1.f / inv_n;
        const float unbiased = var * (n / fmaxf(n - 1.f, 1.f));
        b.moving[c0 + i] = b.moving[c0 + i] * p.momentum + mean * (1.f - p.momentum);
        b.moving[C + c0 + i] = b.moving[C + c0 + i] * p.momentum + unbiased * (1.f - p.momentum);
      }
    }
  }
}

template <bool ONE, bool YF32, bool OF16>
__global__ void __launch_bounds__(kEwMaxThreads, (ONE && !YF32) ? 2 : 1)
bn_apply_fwd_kernel(const BnFwdParams p, int rows_per_block) {
  const int C = p.C;
  const int M = p.B * p.T;
  const float inv_n = 1.f / (float)M;
  const RowTile t = make_row_tile(M, C, rows_per_block);
  const bool owner = blockIdx.x == 0 && t.r == 0;
  const int c0 = t.cv * 8;
  // single-branch layers keep their coefficients in registers for the whole kernel; dense-residual
  // layers recompute them per branch and 4-row group from L1-resident vectors (no shared memory,
  // no per-CTA prologue over all channels)
  float sc0[8], sf0[8];
  bn_coef(p, p.br[0], c0, inv_n, owner, sc0, sf0);
  if (!ONE && owner) {
    for (int j = 1; j < p.n_branch; ++j) {
      float a_[8], b_[8];
      bn_coef(p, p.br[j], c0, inv_n, true, a_, b_);
    }
  }
  const float inv_keep = 1.f / p.keep;
  const uint32_t keep16 = (uint32_t)(p.keep * 65536.f + 0.5f);
  unsigned long long seed = p.seed;
  if (p.step_ctr) seed += (unsigned long long)(*p.step_ctr) * 0x9E3779B97F4A7C15ull;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const size_t rs = (size_t)t.CV;
  // activation, dropout, row mask and the 16-byte store of one row vector
  auto finish = [&](float (&o)[8], int ru, bool live) {
    if (live) {
      if (p.apply_relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = fmaxf(o[i], 0.f);
          if (p.relu_clip > 0.f) o[i] = fminf(o[i], p.relu_clip);
        }
      }
      if (p.keep < 1.f) {
        const unsigned long long idx = (unsigned long long)ru * (unsigned)t.CV + (unsigned)t.cv;
        const uint4 r0 = philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u), key);
        const uint32_t rr[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t u16 = (i & 1) ? (rr[i >> 1] >> 16) : (rr[i >> 1] & 0xFFFFu);
          o[i] = (u16 < keep16) ? o[i] * inv_keep : 0.f;  // P(keep) = keep16 / 65536
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    reinterpret_cast<uint4*>(p.out)[(size_t)ru * rs + t.cv] = OF16 ? float_to_f16x8(o) : float_to_bf16x8(o);
  };
  // (b, tt) of the thread's first row, advanced incrementally (no division in the loop)
  int row = t.row0 + t.r;
  int b = row / p.T, tt = row - b * p.T;
  for (; row < t.row1; row += kUnroll * t.RP) {
    bool live[kUnroll];
    int bb = b, t2 = tt;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int ru = row + u * t.RP;
      live[u] = ru < t.row1 && (p.lens == nullptr || t2 < __ldg(p.lens + min(bb, p.B - 1)));
      t2 += t.RP;
      while (t2 >= p.T) { t2 -= p.T; ++bb; }
    }
    if (ONE) {
      const size_t ys = (size_t)(p.br[0].ld >> 3);
      YVec<YF32> v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        v[u] = live[u] ? ld_yvec<YF32>(p.br[0].y, (size_t)(row + u * t.RP) * ys + t.cv) : zero_yvec<YF32>();
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int ru = row + u * t.RP;
        if (ru >= t.row1) break;
        float f[8];
        yvec_to_float<YF32>(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = f[i] * sc0[i] + sf0[i];
        finish(f, ru, live[u]);
      }
    } else {
      float acc[kUnroll][8];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[u][i] = 0.f;
      for (int j = 0; j < p.n_branch; ++j) {
        const size_t ys = (size_t)(p.br[j].ld >> 3);
        YVec<YF32> v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          v[u] = live[u] ? ld_yvec<YF32>(p.br[j].y, (size_t)(row + u * t.RP) * ys + t.cv) : zero_yvec<YF32>();
        float sc[8], sf[8];
        if (j == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { sc[i] = sc0[i]; sf[i] = sf0[i]; }
        } else {
          bn_coef(p, p.br[j], c0, inv_n, false, sc, sf);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          float f[8];
          yvec_to_float<YF32>(v[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[u][i] += f[i] * sc[i] + sf[i];
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int ru = row + u * t.RP;
        if (ru >= t.row1) break;
        finish(acc[u], ru, live[u]);
      }
    }
    // advance (b, tt) by kUnroll * RP rows
    tt += kUnroll * t.RP;
    while (tt >= p.T) { tt -= p.T; ++b; }
  }
}

int bn_apply_fwd(const BnFwdParams& p, cudaStream_t st) {
  if (p.n_branch < 1 || p.n_branch > kMaxBranches) return fail(ERR_INVALID, "bn_apply_fwd: bad branch count");
  if (p.C % 8 != 0 || p.C > 2048) return fail(ERR_UNSUPPORTED, "bn_apply_fwd: C must be a multiple of 8, <= 2048");
  for (int j = 0; j < p.n_branch; ++j)
    if (p.br[j].ld < p.C || p.br[j].ld % 8 != 0 || p.br[j].stats_ld < p.C || p.br[j].stats_ld % 4 != 0)
      return fail(ERR_INVALID, "bn_apply_fwd: bad leading dimension");
  const int M = p.B * p.T;
  const bool one = p.n_branch == 1;
  const int rpb = rows_per_block_for(M, p.C, (one && !p.y_f32) ? 2 : 1);
  const int grid = (M + rpb - 1) / rpb;
  const int nt = ew_threads(p.C);
#define OS2S_BN_FWD(ONE_, Y_, O_) bn_apply_fwd_kernel<ONE_, Y_, O_><<<grid, nt, 0, st>>>(p, rpb)
  if (one) {
    if (p.y_f32) { if (p.out_f16) OS2S_BN_FWD(true, true, true); else OS2S_BN_FWD(true, true, false); }
    else { if (p.out_f16) OS2S_BN_FWD(true, false, true); else OS2S_BN_FWD(true, false, false); }
  } else {
    if (p.y_f32) { if (p.out_f16) OS2S_BN_FWD(false, true, true); else OS2S_BN_FWD(false, true, false); }
    else { if (p.out_f16) OS2S_BN_FWD(false, false, true); else OS2S_BN_FWD(false, false, false); }
  }
#undef OS2S_BN_FWD
  return check_launch("bn_apply_fwd");
}

// ------------------------------------------------------------------------------- backward
// dz = dA * [a != 0] / keep           (relu + dropout + row mask folded into "a != 0")
// pass 1: dbeta = sum dz (shared by all branches), S_j = sum dz * y_j
//         (dgamma_j = invstd_j * (S_j - mean_j * dbeta), finalised in pass 2's prologue)
// pass 2: dy_j = gamma_j * invstd_j * (dz - dbeta/N - xhat_j * dgamma_j/N) = A_j*dz + B_j*y_j + C_j
template <bool F32>
__device__ __forceinline__ void load_dz_raw(const BnBwdParams& p, size_t vec_index, uint4& a_raw, float (&dz)[8]) {
  // vec_index = row * CV + cv
  if (F32) {
    const float4* src = reinterpret_cast<const float4*>(p.dA) + vec_index * 2;
    const float4 a0 = __ldg(src), a1 = __ldg(src + 1);
    dz[0] = a0.x; dz[1] = a0.y; dz[2] = a0.z; dz[3] = a0.w;
    dz[4] = a1.x; dz[5] = a1.y; dz[6] = a1.z; dz[7] = a1.w;
  } else {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p.dA) + vec_index);
    if (p.h_f16) f16x8_to_float(raw, dz);
    else bf16x8_to_float(raw, dz);
  }
  if (p.apply_relu) a_raw = __ldg(reinterpret_cast<const uint4*>(p.a) + vec_index);
}
__device__ __forceinline__ void gate_dz(const BnBwdParams& p, const uint4& a_raw, float (&dz)[8]) {
  if (p.apply_relu) {
    // "a != 0" as a bit test: valid for bf16 and fp16 layer outputs alike (and immune to flush-to-zero)
    const uint32_t w[4] = {a_raw.x, a_raw.y, a_raw.z, a_raw.w};
    const float inv_keep = 1.f / p.keep;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dz[2 * i] = (w[i] & 0x7FFFu) ? dz[2 * i] * inv_keep : 0.f;
      dz[2 * i + 1] = (w[i] & 0x7FFF0000u) ? dz[2 * i + 1] * inv_keep : 0.f;
    }
  }
}

// G = branches reduced per sweep over the rows: 1 for single-branch layers (64 registers, two CTAs
// per SM), 4 for dense-residual layers.
template <bool F32, int G, bool YF32>
__global__ void __launch_bounds__(kEwMaxThreads, (G == 1 && !YF32) ? 2 : 1)
bn_bwd_reduce_kernel(const BnBwdParams p, int rows_per_block) {
  extern __shared__ float sh[];  // [1 + n_branch][C]
  const int C = p.C;
  const int nred = (1 + p.n_branch) * C;
  for (int i = threadIdx.x; i < nred; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const RowTile t = make_row_tile(p.M, C, rows_per_block);
  {
    const size_t rs = (size_t)t.CV;
    for (int j0 = 0; j0 < p.n_branch; j0 += G) {
      const int nj = min(G, p.n_branch - j0);
      float db[8], S[G][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        db[i] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) S[g][i] = 0.f;
      }
      for (int row = t.row0 + t.r; row < t.row1; row += 2 * t.RP) {
        const bool two = row + t.RP < t.row1;
        float dz0[8], dz1[8];
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
        YVec<YF32> y0[G], y1[G];
        const size_t v0 = (size_t)row * rs + t.cv, v1 = v0 + (size_t)t.RP * rs;
        load_dz_raw<F32>(p, v0, a0, dz0);
        if (two) load_dz_raw<F32>(p, v1, a1, dz1);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < nj) {
            const size_t ys = (size_t)(p.br[j0 + g].ld >> 3);
            y0[g] = ld_yvec<YF32>(p.br[j0 + g].y, (size_t)row * ys + t.cv);
            if (two) y1[g] = ld_yvec<YF32>(p.br[j0 + g].y, (size_t)(row + t.RP) * ys + t.cv);
          }
        }
        gate_dz(p, a0, dz0);
        if (two) gate_dz(p, a1, dz1);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) dz1[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) db[i] += dz0[i] + dz1[i];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < nj) {
            float f[8];
            yvec_to_float<YF32>(y0[g], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) S[g][i] += dz0[i] * f[i];
            if (two) {
              yvec_to_float<YF32>(y1[g], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) S[g][i] += dz1[i] * f[i];
            }
          }
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (g < nj) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(&sh[(1 + j0 + g) * C + t.cv * 8 + i], S[g][i]);
        }
      }
      if (j0 == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&sh[t.cv * 8 + i], db[i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nred; i += blockDim.x) atomicAdd(&p.red[i], sh[i]);
}

// dy = A*dz + Bc*y + Cc for the thread's own 8 channels of branch b (from the pass-1 sums).
__device__ __forceinline__ void bn_bwd_coef(const BnBwdParams& p, int j, int c0, float inv_n, bool owner,
                                            float (&A)[8], float (&Bc)[8], float (&Cc)[8]) {
  const int C = p.C;
  const BnBranchBwd& b = p.br[j];
  float mean[8], invstd[8], g[8], dbeta[8], S[8];
  auto ld8 = [](const float* src, float (&dst)[8]) {
    *reinterpret_cast<float4*>(&dst[0]) = __ldg(reinterpret_cast<const float4*>(src));
    *reinterpret_cast<float4*>(&dst[4]) = __ldg(reinterpret_cast<const float4*>(src) + 1);
  };
  ld8(b.mean_invstd + c0, mean);
  ld8(b.mean_invstd + C + c0, invstd);
  ld8(b.gamma + c0, g);
  ld8(p.red + c0, dbeta);
  ld8(p.red + (size_t)(1 + j) * C + c0, S);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float dgamma = invstd[i] * (S[i] - mean[i] * dbeta[i]);
    A[i] = g[i] * invstd[i];
    Bc[i] = -A[i] * invstd[i] * dgamma * inv_n;
    Cc[i] = -A[i] * dbeta[i] * inv_n - Bc[i] * mean[i];
    if (owner) {
      b.dgamma[c0 + i] = dgamma;
      b.dbeta[c0 + i] = dbeta[i];
    }
  }
}

template <bool F32, bool ONE, bool YF32>
__global__ void __launch_bounds__(kEwMaxThreads, (ONE && !YF32) ? 2 : 1)
bn_bwd_apply_kernel(const BnBwdParams p, int rows_per_block) {
  constexpr int U = (ONE || YF32) ? 2 : kUnroll;   // rows in flight per thread (3-4 loads each when ONE)
  const int C = p.C;
  const float inv_n = 1.f / (float)p.M;
  const RowTile t = make_row_tile(p.M, C, rows_per_block);
  const bool owner = blockIdx.x == 0 && t.r == 0;
  const int c0 = t.cv * 8;
  float A0[8], B0[8], C0[8];
  bn_bwd_coef(p, 0, c0, inv_n, owner, A0, B0, C0);
  if (!ONE && owner) {
    for (int j = 1; j < p.n_branch; ++j) {
      float a_[8], b_[8], c_[8];
      bn_bwd_coef(p, j, c0, inv_n, true, a_, b_, c_);
    }
  }
  const size_t rs = (size_t)t.CV;
  for (int row = t.row0 + t.r; row < t.row1; row += U * t.RP) {
    float dz[U][8];
    uint4 araw[U];
    bool live[U];
    if (ONE) {
      // all loads of the U rows (dA, a, y) are issued before the first use
      uint4* db = reinterpret_cast<uint4*>(p.br[0].dy) + t.cv;
      const size_t ys = (size_t)(p.br[0].ld >> 3);
      YVec<YF32> v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        live[u] = row + u * t.RP < t.row1;
        araw[u] = make_uint4(0, 0, 0, 0);
        if (live[u]) {
          load_dz_raw<F32>(p, (size_t)(row + u * t.RP) * rs + t.cv, araw[u], dz[u]);
          v[u] = ld_yvec<YF32>(p.br[0].y, (size_t)(row + u * t.RP) * ys + t.cv);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (live[u]) {
          gate_dz(p, araw[u], dz[u]);
          float f[8];
          yvec_to_float<YF32>(v[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = A0[i] * dz[u][i] + B0[i] * f[i] + C0[i];
          db[(size_t)(row + u * t.RP) * ys] = p.h_f16 ? float_to_f16x8(f) : float_to_bf16x8(f);
        }
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      live[u] = row + u * t.RP < t.row1;
      araw[u] = make_uint4(0, 0, 0, 0);
      if (live[u]) load_dz_raw<F32>(p, (size_t)(row + u * t.RP) * rs + t.cv, araw[u], dz[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (live[u]) gate_dz(p, araw[u], dz[u]);
    for (int j = 0; j < p.n_branch; ++j) {
      uint4* db = reinterpret_cast<uint4*>(p.br[j].dy) + t.cv;
      const size_t ys = (size_t)(p.br[j].ld >> 3);
      YVec<YF32> v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) v[u] = ld_yvec<YF32>(p.br[j].y, (size_t)(row + u * t.RP) * ys + t.cv);
      float A[8], Bc[8], Cc[8];
      if (j == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { A[i] = A0[i]; Bc[i] = B0[i]; Cc[i] = C0[i]; }
      } else {
        bn_bwd_coef(p, j, c0, inv_n, false, A, Bc, Cc);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (live[u]) {
          float f[8], o[8];
          yvec_to_float<YF32>(v[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = A[i] * dz[u][i] + Bc[i] * f[i] + Cc[i];
          db[(size_t)(row + u * t.RP) * ys] = p.h_f16 ? float_to_f16x8(o) : float_to_bf16x8(o);
        }
      }
    }
  }
}

int bn_bwd(const BnBwdParams& p, cudaStream_t st, bool reduce) {
  if (p.n_branch < 1 || p.n_branch > kMaxBranches) return fail(ERR_INVALID, "bn_bwd: bad branch count");
  if (p.C % 8 != 0 || p.C > 2048) return fail(ERR_UNSUPPORTED, "bn_bwd: C must be a multiple of 8, <= 2048");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<true, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_done = true;
  }
  for (int j = 0; j < p.n_branch; ++j)
    if (p.br[j].ld < p.C || p.br[j].ld % 8 != 0) return fail(ERR_INVALID, "bn_bwd: bad leading dimension");
  const size_t smem_r = (size_t)(1 + p.n_branch) * p.C * sizeof(float);
  if (smem_r > 100 * 1024) return fail(ERR_UNSUPPORTED, "bn_bwd: too many branches x channels");
  const bool one = p.n_branch == 1;
  const int rpb = rows_per_block_for(p.M, p.C, (one && !p.y_f32) ? 2 : 1);
  const int grid = (p.M + rpb - 1) / rpb;
  const int nt = ew_threads(p.C);
#define OS2S_BN_RED(F_, G_) do { if (p.y_f32) bn_bwd_reduce_kernel<F_, (G_ == 4 ? 2 : G_), true><<<grid, nt, smem_r, st>>>(p, rpb); \
                                 else bn_bwd_reduce_kernel<F_, G_, false><<<grid, nt, smem_r, st>>>(p, rpb); } while (0)
#define OS2S_BN_APP(F_, ONE_) do { if (p.y_f32) bn_bwd_apply_kernel<F_, ONE_, true><<<grid, nt, 0, st>>>(p, rpb); \
                                   else bn_bwd_apply_kernel<F_, ONE_, false><<<grid, nt, 0, st>>>(p, rpb); } while (0)
  if (!reduce) {
    // the reductions were accumulated by the data-gradient kernel that produced dA (single branch, bf16)
    if (!one || p.dA_is_f32) return fail(ERR_INVALID, "bn_bwd: apply-only needs one branch and a bf16 dA");
    OS2S_BN_APP(false, true);
  } else if (one) {
    if (p.dA_is_f32) {
      OS2S_BN_RED(true, 1);
      OS2S_BN_APP(true, true);
    } else {
      OS2S_BN_RED(false, 1);
      OS2S_BN_APP(false, true);
    }
  } else if (p.dA_is_f32) {
    OS2S_BN_RED(true, 4);
    OS2S_BN_APP(true, false);
  } else {
    OS2S_BN_RED(false, 4);
    OS2S_BN_APP(false, false);
  }
#undef OS2S_BN_RED
#undef OS2S_BN_APP
  return check_launch("bn_bwd");
}

}  // namespace os2s
