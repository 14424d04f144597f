// HBM-bound elementwise / reduction kernels of the Jasper path.
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace os2s {

// ---------------------------------------------------------------- weight cast + transpose
// w fp32 [K][R][C] -> w_bf16 [K][R][C] and wt_bf16 [K][C][R]; 32x32 smem tile, coalesced both ways.
__global__ void weight_cast_transpose_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb,
                                             __nv_bfloat16* __restrict__ wt, int R, int C) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* src = w + (size_t)k * R * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = src[(size_t)r * C + c];
      if (wb) wb[(size_t)k * R * C + (size_t)r * C + c] = __float2bfloat16(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (wt) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) wt[(size_t)k * R * C + (size_t)c * R + r] = __float2bfloat16(tile[threadIdx.x][i]);
    }
  }
}

int weight_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int K, int C_in, int C_out,
                          cudaStream_t st) {
  if (K <= 0 || C_in <= 0 || C_out <= 0) return fail(ERR_INVALID, "weight_cast_transpose: bad shape");
  dim3 grid((C_out + 31) / 32, (C_in + 31) / 32, K), block(32, 8);
  weight_cast_transpose_kernel<<<grid, block, 0, st>>>(w, (__nv_bfloat16*)w_bf16, (__nv_bfloat16*)wt_bf16,
                                                       C_in, C_out);
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
__device__ __forceinline__ uint4 float_to_bf16x8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// Per-channel sum and sum of squares of y [M, C] (bf16). stats = [2][C] fp32, pre-zeroed.
// Thread layout: each thread owns one 8-channel vector (16-byte loads) and walks rows.
constexpr int kStatThreads = 256;
__global__ void __launch_bounds__(kStatThreads)
bn_stats_kernel(const __half* __restrict__ y, float* __restrict__ stats, int M, int C,
                int rows_per_block) {
  extern __shared__ float sh[];  // [2][C]
  const int CV = C >> 3;
  const int RP = kStatThreads / CV;  // rows processed per pass
  const int cv = threadIdx.x % CV;
  const int r = threadIdx.x / CV;
  for (int i = threadIdx.x; i < 2 * C; i += kStatThreads) sh[i] = 0.f;
  __syncthreads();
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (r < RP) {
    const int row0 = blockIdx.x * rows_per_block;
    const int row1 = min(M, row0 + rows_per_block);
    for (int row = row0 + r; row < row1; row += RP) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(y + (size_t)row * C) + cv);
      float f[8];
      f16x8_to_float(v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] += f[i] * f[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&sh[cv * 8 + i], s[i]);
      atomicAdd(&sh[C + cv * 8 + i], q[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kStatThreads) atomicAdd(&stats[i], sh[i]);
}

int bn_stats(const void* y, float* stats, int M, int C, cudaStream_t st) {
  if (C % 8 != 0 || C > 2048 || C < 8) return fail(ERR_UNSUPPORTED, "bn_stats: C must be a multiple of 8, <= 2048");
  const int target_blocks = device_sm_count() * 4;
  int rows_per_block = (M + target_blocks - 1) / target_blocks;
  if (rows_per_block < 8) rows_per_block = 8;
  const int grid = (M + rows_per_block - 1) / rows_per_block;
  bn_stats_kernel<<<grid, kStatThreads, 2 * C * sizeof(float), st>>>((const __half*)y, stats, M, C,
                                                                    rows_per_block);
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
// Philox4x32-10 (Salmon et al.), counter = element-vector index, key = seed.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo32(0xD2511F53u, ctr.x, &hi0);
    const uint32_t lo1 = mulhilo32(0xCD9E8D57u, ctr.z, &hi1);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

constexpr int kApplyThreads = 256;
__global__ void __launch_bounds__(kApplyThreads)
bn_apply_fwd_kernel(const BnFwdParams p) {
  extern __shared__ float sh[];  // [n_branch][2][C]: scale, shift
  const int C = p.C;
  const float inv_n = 1.f / (float)((long long)p.B * p.T);
  for (int i = threadIdx.x; i < p.n_branch * C; i += kApplyThreads) {
    const int j = i / C, c = i - j * C;
    const BnBranchFwd& b = p.br[j];
    // training: batch statistics; inference (use_moving): the moving averages
    const float mean = p.use_moving ? b.moving[c] : b.stats[c] * inv_n;
    const float var = p.use_moving ? b.moving[C + c] : fmaxf(b.stats[C + c] * inv_n - mean * mean, 0.f);
    const float invstd = rsqrtf(var + p.eps);
    const float g = b.gamma[c];
    sh[(j * 2) * C + c] = g * invstd;
    sh[(j * 2 + 1) * C + c] = b.beta[c] - mean * g * invstd;
    if (blockIdx.x == 0 && !p.use_moving) {
      b.mean_invstd[c] = mean;
      b.mean_invstd[C + c] = invstd;
      if (b.moving) {
        const float n = (float)((long long)p.B * p.T);
        const float unbiased = var * (n / fmaxf(n - 1.f, 1.f));
        b.moving[c] = b.moving[c] * p.momentum + mean * (1.f - p.momentum);
        b.moving[C + c] = b.moving[C + c] * p.momentum + unbiased * (1.f - p.momentum);
      }
    }
  }
  __syncthreads();
  const int CV = C >> 3;
  const long long total = (long long)p.B * p.T * CV;
  const float inv_keep = 1.f / p.keep;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  for (long long idx = (long long)blockIdx.x * kApplyThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kApplyThreads) {
    const long long row = idx / CV;
    const int cv = (int)(idx - row * CV);
    const int b = (int)(row / p.T);
    const int t = (int)(row - (long long)b * p.T);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const bool valid = (p.lens == nullptr) || (t < p.lens[b]);
    if (valid) {
      for (int j = 0; j < p.n_branch; ++j) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.br[j].y + row * C) + cv);
        float f[8];
        f16x8_to_float(v, f);
        const float* sc = &sh[(j * 2) * C + cv * 8];
        const float* sf = &sh[(j * 2 + 1) * C + cv * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i] * sc[i] + sf[i];
      }
      if (p.apply_relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] = fmaxf(acc[i], 0.f);
          if (p.relu_clip > 0.f) acc[i] = fminf(acc[i], p.relu_clip);
        }
      }
      if (p.keep < 1.f) {
        const uint4 r0 = philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u), key);
        const uint4 r1 = philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 1u, 0u), key);
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float u = (float)(rr[i] >> 8) * (1.f / 16777216.f);  // [0,1)
          acc[i] = (u < p.keep) ? acc[i] * inv_keep : 0.f;
        }
      }
    }
    reinterpret_cast<uint4*>(p.out + row * C)[cv] = float_to_bf16x8(acc);
  }
}

int bn_apply_fwd(const BnFwdParams& p, cudaStream_t st) {
  if (p.n_branch < 1 || p.n_branch > kMaxBranches) return fail(ERR_INVALID, "bn_apply_fwd: bad branch count");
  if (p.C % 8 != 0) return fail(ERR_UNSUPPORTED, "bn_apply_fwd: C must be a multiple of 8");
  const size_t smem = (size_t)p.n_branch * 2 * p.C * sizeof(float);
  if (smem > 200 * 1024) return fail(ERR_UNSUPPORTED, "bn_apply_fwd: too many branches x channels");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(bn_apply_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const long long total = (long long)p.B * p.T * (p.C / 8);
  long long blocks = (total + kApplyThreads - 1) / kApplyThreads;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  bn_apply_fwd_kernel<<<(int)blocks, kApplyThreads, smem, st>>>(p);
  return check_launch("bn_apply_fwd");
}

// ------------------------------------------------------------------------------- backward
// dz = dA * [a != 0] / keep           (relu + dropout + row mask folded into "a != 0")
// pass 1: dbeta = sum dz (shared by all branches), dgamma_j = sum dz * xhat_j
// pass 2: dy_j = gamma_j * invstd_j * (dz - dbeta/N - xhat_j * dgamma_j/N)

template <bool F32>
__device__ __forceinline__ void load_dz(const BnBwdParams& p, long long row, int cv, float (&dz)[8]) {
  const int C = p.C;
  if (F32) {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.dA) + row * C) + cv * 2;
    const float4 a0 = __ldg(src), a1 = __ldg(src + 1);
    dz[0] = a0.x; dz[1] = a0.y; dz[2] = a0.z; dz[3] = a0.w;
    dz[4] = a1.x; dz[5] = a1.y; dz[6] = a1.z; dz[7] = a1.w;
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dA) + row * C) + cv);
    bf16x8_to_float(v, dz);
  }
  if (p.apply_relu) {
    const uint4 av = __ldg(reinterpret_cast<const uint4*>(p.a + row * C) + cv);
    float af[8];
    bf16x8_to_float(av, af);
    const float inv_keep = 1.f / p.keep;
#pragma unroll
    for (int i = 0; i < 8; ++i) dz[i] = (af[i] != 0.f) ? dz[i] * inv_keep : 0.f;
  }
}

template <bool F32>
__global__ void __launch_bounds__(kStatThreads)
bn_bwd_reduce_kernel(const BnBwdParams p, int rows_per_block) {
  extern __shared__ float sh[];  // [1 + n_branch][C]
  const int C = p.C;
  const int CV = C >> 3;
  const int RP = kStatThreads / CV;
  const int cv = threadIdx.x % CV;
  const int r = threadIdx.x / CV;
  const int nred = (1 + p.n_branch) * C;
  for (int i = threadIdx.x; i < nred; i += kStatThreads) sh[i] = 0.f;
  __syncthreads();
  if (r < RP) {
    const int row0 = blockIdx.x * rows_per_block;
    const int row1 = min(p.M, row0 + rows_per_block);
    float db[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) db[i] = 0.f;
    // branch loop outside the row loop would re-read dz; keep per-branch accumulators in smem atomics
    // only at the end: process branches one at a time over the row range (dz is L2/L1 resident).
    for (int j = 0; j < p.n_branch; ++j) {
      float dg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dg[i] = 0.f;
      const float* mi = p.br[j].mean_invstd;
      float mean[8], invstd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mean[i] = mi[cv * 8 + i];
        invstd[i] = mi[C + cv * 8 + i];
      }
      for (int row = row0 + r; row < row1; row += RP) {
        float dz[8];
        load_dz<F32>(p, row, cv, dz);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.br[j].y + (size_t)row * C) + cv);
        float f[8];
        f16x8_to_float(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dg[i] += dz[i] * (f[i] - mean[i]) * invstd[i];
          if (j == 0) db[i] += dz[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sh[(1 + j) * C + cv * 8 + i], dg[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sh[cv * 8 + i], db[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nred; i += kStatThreads) atomicAdd(&p.red[i], sh[i]);
}

template <bool F32>
__global__ void __launch_bounds__(kApplyThreads)
bn_bwd_apply_kernel(const BnBwdParams p) {
  extern __shared__ float sh[];  // per branch: k1 = gamma*invstd, k2 = dbeta/N, k3 = dgamma/N*invstd, mean
  const int C = p.C;
  const float inv_n = 1.f / (float)p.M;
  for (int i = threadIdx.x; i < p.n_branch * C; i += kApplyThreads) {
    const int j = i / C, c = i - j * C;
    const BnBranchBwd& b = p.br[j];
    const float mean = b.mean_invstd[c], invstd = b.mean_invstd[C + c];
    const float dbeta = p.red[c], dgamma = p.red[(1 + j) * C + c];
    float* s = &sh[(size_t)j * 4 * C];
    s[c] = b.gamma[c] * invstd;
    s[C + c] = dbeta * inv_n;
    s[2 * C + c] = dgamma * inv_n * invstd;
    s[3 * C + c] = mean;
    if (blockIdx.x == 0) {
      b.dgamma[c] = dgamma;
      b.dbeta[c] = dbeta;
    }
  }
  __syncthreads();
  const int CV = C >> 3;
  const long long total = (long long)p.M * CV;
  for (long long idx = (long long)blockIdx.x * kApplyThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kApplyThreads) {
    const long long row = idx / CV;
    const int cv = (int)(idx - row * CV);
    float dz[8];
    load_dz<F32>(p, row, cv, dz);
    for (int j = 0; j < p.n_branch; ++j) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.br[j].y + row * C) + cv);
      float f[8], o[8];
      f16x8_to_float(v, f);
      const float* s = &sh[(size_t)j * 4 * C + cv * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = s[i] * (dz[i] - s[C + i] - (f[i] - s[3 * C + i]) * s[2 * C + i]);
      reinterpret_cast<uint4*>(p.br[j].dy + row * C)[cv] = float_to_bf16x8(o);
    }
  }
}

int bn_bwd(const BnBwdParams& p, cudaStream_t st) {
  if (p.n_branch < 1 || p.n_branch > kMaxBranches) return fail(ERR_INVALID, "bn_bwd: bad branch count");
  if (p.C % 8 != 0 || p.C > 2048) return fail(ERR_UNSUPPORTED, "bn_bwd: C must be a multiple of 8, <= 2048");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_done = true;
  }
  const size_t smem_r = (size_t)(1 + p.n_branch) * p.C * sizeof(float);
  const size_t smem_a = (size_t)p.n_branch * 4 * p.C * sizeof(float);
  if (smem_a > 200 * 1024 || smem_r > 100 * 1024) return fail(ERR_UNSUPPORTED, "bn_bwd: too many branches x channels");
  const int target_blocks = device_sm_count() * 4;
  int rows_per_block = (p.M + target_blocks - 1) / target_blocks;
  if (rows_per_block < 8) rows_per_block = 8;
  const int grid_r = (p.M + rows_per_block - 1) / rows_per_block;
  const long long total = (long long)p.M * (p.C / 8);
  long long blocks = (total + kApplyThreads - 1) / kApplyThreads;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (p.dA_is_f32) {
    bn_bwd_reduce_kernel<true><<<grid_r, kStatThreads, smem_r, st>>>(p, rows_per_block);
    bn_bwd_apply_kernel<true><<<(int)blocks, kApplyThreads, smem_a, st>>>(p);
  } else {
    bn_bwd_reduce_kernel<false><<<grid_r, kStatThreads, smem_r, st>>>(p, rows_per_block);
    bn_bwd_apply_kernel<false><<<(int)blocks, kApplyThreads, smem_a, st>>>(p);
  }
  return check_launch("bn_bwd");
}

}  // namespace os2s
