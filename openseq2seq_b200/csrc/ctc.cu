// FullyConnectedCTCDecoder projection, CTC loss forward/backward and greedy CTC decoding.
//
// Reference ops being replaced:
//   tf.layers.dense            open_seq2seq/decoders/fc_decoders.py:135-140   (K5)
//   tf.nn.ctc_loss             open_seq2seq/losses/ctc_loss.py:77-89           (K6, a CPU kernel in TF1)
//   tf.nn.ctc_greedy_decoder   open_seq2seq/decoders/fc_decoders.py:247-250   (K7)
// Logits are kept batch-major [B, T, V] fp32 in HBM (the time-major [T,B,V] view the reference
// exposes is a stride permutation, never a copy); all kernels take explicit (t, b) strides.
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

namespace os2s {

constexpr float kNegBig = -1e30f;

__device__ __forceinline__ float log_add(float a, float b) {
  const float m = fmaxf(a, b);
  const float d = -fabsf(a - b);
  return m + log1pf(__expf(d));
}

// ------------------------------------------------------------------ FC forward (small-N GEMM)
// logits[m, v] = sum_h x[m, h] * w[h, v] + bias[v];  x bf16 [M, H], w fp32 [H, V] staged in smem.
// HBM-bound (49 MB of activations for 1.4 GFLOP): one 768-thread CTA per SM, a warp takes two rows at a
// time, stages them in shared memory with 16-byte coalesced loads (all loads of a row pair in flight
// before the first use), then lane l owns h = l, l+32, ... (row stride V = 29 floats: conflict-free).
// layer outputs are bf16, or fp16 with OS2S_ACT_F16
template <bool XF16>
__device__ __forceinline__ float act_to_float(uint16_t v) {
  if (XF16) return __half2float(__ushort_as_half(v));
  return __uint_as_float((uint32_t)v << 16);
}
constexpr int kFcThreads = 768;
constexpr int kFcWarps = kFcThreads / 32;
template <int VMAX, bool XF16>
__global__ void __launch_bounds__(kFcThreads)
fc_fwd_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              float* __restrict__ logits, int M, int H, int V) {
  extern __shared__ float wsh[];  // [H][V] fp32, then per warp 2 rows of x (bf16)
  uint16_t* xsh = reinterpret_cast<uint16_t*>(wsh + (((size_t)H * V + 3) & ~(size_t)3));
  for (int i = threadIdx.x; i < H * V; i += kFcThreads) wsh[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* xw = xsh + (size_t)warp * 2 * H;
  const int warp_global = blockIdx.x * kFcWarps + warp;
  const int n_warps = gridDim.x * kFcWarps;
  const int hv = H >> 3;  // 16-byte vectors per row
  for (int m0 = warp_global * 2; m0 < M; m0 += n_warps * 2) {
    const bool two = (m0 + 1) < M;
    const uint4* g0 = reinterpret_cast<const uint4*>(x + (size_t)m0 * H);
    const uint4* g1 = reinterpret_cast<const uint4*>(x + (size_t)(two ? m0 + 1 : m0) * H);
    for (int i = lane; i < hv; i += 32) {
      reinterpret_cast<uint4*>(xw)[i] = __ldg(g0 + i);
      reinterpret_cast<uint4*>(xw + H)[i] = __ldg(g1 + i);
    }
    __syncwarp();
    float acc0[VMAX], acc1[VMAX];
#pragma unroll
    for (int v = 0; v < VMAX; ++v) acc0[v] = acc1[v] = 0.f;
    for (int h = lane; h < H; h += 32) {
      const float a0 = act_to_float<XF16>(xw[h]);
      const float a1 = act_to_float<XF16>(xw[H + h]);
      const float* wr = &wsh[h * V];
#pragma unroll
      for (int v = 0; v < VMAX; ++v) {
        if (v < V) {
          const float wv = wr[v];
          acc0[v] += a0 * wv;
          acc1[v] += a1 * wv;
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc0[v] += __shfl_xor_sync(0xffffffffu, acc0[v], o);
        acc1[v] += __shfl_xor_sync(0xffffffffu, acc1[v], o);
      }
    }
    // lane v writes column v
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
      if (lane == v) {
        r0 = acc0[v];
        r1 = acc1[v];
      }
    }
    if (lane < V) {
      const float b = bias ? bias[lane] : 0.f;
      logits[(size_t)m0 * V + lane] = r0 + b;
      if (two) logits[(size_t)(m0 + 1) * V + lane] = r1 + b;
    }
  }
}

static size_t fc_smem_bytes(int H, int V) {
  return ((((size_t)H * V + 3) & ~(size_t)3) * sizeof(float)) + (size_t)kFcWarps * 2 * H * sizeof(__nv_bfloat16);
}

int fc_fwd(const void* x, const float* w, const float* bias, float* logits, int M, int H, int V,
           cudaStream_t st, int x_f16) {
  if (V > 32 || V < 1) return fail(ERR_UNSUPPORTED, "fc_fwd: vocabulary > 32 not supported by the small-N kernel");
  if (H % 8 != 0) return fail(ERR_UNSUPPORTED, "fc_fwd: H must be a multiple of 8");
  const size_t smem = fc_smem_bytes(H, V);
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "fc_fwd: H*V too large for shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(fc_fwd_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(fc_fwd_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  const int grid = device_sm_count();
  if (x_f16) fc_fwd_kernel<32, true><<<grid, kFcThreads, smem, st>>>((const uint16_t*)x, w, bias, logits, M, H, V);
  else fc_fwd_kernel<32, false><<<grid, kFcThreads, smem, st>>>((const uint16_t*)x, w, bias, logits, M, H, V);
  return check_launch("fc_fwd");
}

// ------------------------------------------------------------------ FC backward
// dx[m, h] = sum_v dl[m, v] * w[h, v]   (bf16 out);  one warp per row, lanes over h; the row is
// assembled in shared memory and written with 16-byte coalesced stores.
__global__ void __launch_bounds__(kFcThreads)
fc_dgrad_kernel(const float* __restrict__ dl, const float* __restrict__ w, uint16_t* __restrict__ dx,
                int M, int H, int V, int f16) {
  extern __shared__ float wsh[];  // [H][V], then per warp one output row (bf16 / fp16)
  uint16_t* xsh = reinterpret_cast<uint16_t*>(wsh + (((size_t)H * V + 3) & ~(size_t)3));
  for (int i = threadIdx.x; i < H * V; i += kFcThreads) wsh[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* xw = xsh + (size_t)warp * 2 * H;
  const int warp_global = blockIdx.x * kFcWarps + warp;
  const int n_warps = gridDim.x * kFcWarps;
  const int hv = H >> 3;
  for (int m = warp_global; m < M; m += n_warps) {
    const float mine = (lane < V) ? __ldg(dl + (size_t)m * V + lane) : 0.f;
    float d[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) d[v] = __shfl_sync(0xffffffffu, mine, v);
    for (int h = lane; h < H; h += 32) {
      const float* wr = &wsh[h * V];
      float acc = 0.f;
#pragma unroll
      for (int v = 0; v < 32; ++v)
        if (v < V) acc += d[v] * wr[v];
      xw[h] = f16 ? __half_as_ushort(__float2half_rn(acc)) : __bfloat16_as_ushort(__float2bfloat16(acc));
    }
    __syncwarp();
    uint4* g = reinterpret_cast<uint4*>(dx + (size_t)m * H);
    for (int i = lane; i < hv; i += 32) g[i] = reinterpret_cast<const uint4*>(xw)[i];
    __syncwarp();
  }
}

// dw[h, v] += sum_m x[m, h] * dl[m, v];  db[v] += sum_m dl[m, v].  Grid = (row chunks, H / 256): thread
// h of a CTA keeps V partial sums over the CTA's row chunk (dl rows broadcast from shared memory),
// then one atomicAdd per (h, v) -- chunks are sized for ~2 CTAs per SM so the atomics stay few.
constexpr int kFcWgTile = 64;   // dl rows staged per pass
template <bool XF16>
__global__ void __launch_bounds__(256)
fc_wgrad_kernel(const uint16_t* __restrict__ x, const float* __restrict__ dl, float* __restrict__ dw,
                float* __restrict__ db, int M, int H, int V, int rows_per_cta) {
  __shared__ float dsh[kFcWgTile][32];
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(M, row0 + rows_per_cta);
  const int h = blockIdx.y * 256 + threadIdx.x;
  float acc[32];
#pragma unroll
  for (int v = 0; v < 32; ++v) acc[v] = 0.f;
  float dbs = 0.f;
  for (int r0 = row0; r0 < row1; r0 += kFcWgTile) {
    const int rows = min(kFcWgTile, row1 - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * 32; i += 256) {
      const int r = i >> 5, v = i & 31;
      dsh[r][v] = (v < V) ? __ldg(dl + (size_t)(r0 + r) * V + v) : 0.f;
    }
    __syncthreads();
    if (h < H) {
      const uint16_t* xp = x + (size_t)r0 * H + h;
#pragma unroll 8
      for (int r = 0; r < rows; ++r) {
        const float a = act_to_float<XF16>(xp[(size_t)r * H]);
        const float4* drow = reinterpret_cast<const float4*>(dsh[r]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 dv = drow[q];
          acc[4 * q + 0] += a * dv.x;
          acc[4 * q + 1] += a * dv.y;
          acc[4 * q + 2] += a * dv.z;
          acc[4 * q + 3] += a * dv.w;
        }
      }
    }
    if (blockIdx.y == 0 && threadIdx.x < V)
      for (int r = 0; r < rows; ++r) dbs += dsh[r][threadIdx.x];
  }
  if (h < H) {
#pragma unroll
    for (int v = 0; v < 32; ++v)
      if (v < V) atomicAdd(&dw[(size_t)h * V + v], acc[v]);
  }
  if (blockIdx.y == 0 && threadIdx.x < V) atomicAdd(&db[threadIdx.x], dbs);
}

int fc_bwd(const void* x, const float* dl, const float* w, void* dx, float* dw, float* db, int M, int H, int V,
           cudaStream_t st, int x_f16) {
  if (V > 32 || V < 1) return fail(ERR_UNSUPPORTED, "fc_bwd: vocabulary > 32 not supported");
  if (H % 8 != 0) return fail(ERR_UNSUPPORTED, "fc_bwd: H must be a multiple of 8");
  const size_t smem = fc_smem_bytes(H, V);
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "fc_bwd: H*V too large for shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(fc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  if (dx) {
    fc_dgrad_kernel<<<device_sm_count(), kFcThreads, smem, st>>>(dl, w, (uint16_t*)dx, M, H, V, x_f16);
    int s = check_launch("fc_dgrad");
    if (s) return s;
  }
  if (dw) {
    OS2S_CUDA(cudaMemsetAsync(dw, 0, (size_t)H * V * sizeof(float), st));
    OS2S_CUDA(cudaMemsetAsync(db, 0, (size_t)V * sizeof(float), st));
    const int hy = (H + 255) / 256;
    int chunks = (2 * device_sm_count() + hy - 1) / hy;   // ~2 CTAs per SM in total
    int rows_per_cta = (M + chunks - 1) / chunks;
    rows_per_cta = ((rows_per_cta + kFcWgTile - 1) / kFcWgTile) * kFcWgTile;
    chunks = (M + rows_per_cta - 1) / rows_per_cta;
    if (x_f16) fc_wgrad_kernel<true><<<dim3(chunks, hy), 256, 0, st>>>((const uint16_t*)x, dl, dw, db, M, H, V, rows_per_cta);
    else fc_wgrad_kernel<false><<<dim3(chunks, hy), 256, 0, st>>>((const uint16_t*)x, dl, dw, db, M, H, V, rows_per_cta);
    return check_launch("fc_wgrad");
  }
  return OK;
}

// ------------------------------------------------------------------ CTC
// Step 1: lse[b, t] = logsumexp_v logits[b, t, :]   (one thread per (b, t); V <= 32 floats)
__global__ void ctc_lse_kernel(const float* __restrict__ logits, float* __restrict__ lse, int B, int T, int V,
                               long long stride_b, long long stride_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T, t = i - b * T;
  const float* row = logits + b * stride_b + t * stride_t;
  float m = -CUDART_INF_F;
  for (int v = 0; v < V; ++v) m = fmaxf(m, row[v]);
  float s = 0.f;
  for (int v = 0; v < V; ++v) s += __expf(row[v] - m);
  lse[i] = m + __logf(s);
}

// Step 2: alpha (blockIdx.y == 0) and beta (blockIdx.y == 1) lattices, one CTA per utterance and
// direction.  The T time steps are serial, so the step latency is everything: thread i owns states
// s = i, i + 512, ... (NS of them) and keeps their labels and skip-transition flags in registers for
// the whole utterance; a step is three independent shared-memory loads of the previous row, one
// max, three exp, one log, the emission (staged 32 steps at a time with coalesced loads), one shared
// and one HBM store, and one barrier.  alpha/beta rows go to HBM [B][T][S_max] for the gradient kernel.
constexpr int kCtcThreads = 512;
constexpr int kCtcChunk = 32;
template <int NS>
__global__ void __launch_bounds__(kCtcThreads)
ctc_alpha_beta_kernel(const float* __restrict__ logits, const float* __restrict__ lse,
                      const int* __restrict__ labels, const int* __restrict__ label_lens,
                      const int* __restrict__ input_lens, float* __restrict__ alpha, float* __restrict__ beta,
                      float* __restrict__ loglik, int T, int V, int L_max, int S_max, long long stride_b,
                      long long stride_t, int blank) {
  extern __shared__ float sh[];
  // lattice rows carry 2 guard cells on both sides (-inf), so neighbour loads need no bounds checks
  const int LP = S_max + 4;
  float* lat = sh;                                  // [2][LP]
  float* em = sh + 2 * LP;                          // [2][kCtcChunk][32] staged log-probs
  int* ext = reinterpret_cast<int*>(em + 2 * kCtcChunk * 32);  // [S_max]
  const int b = blockIdx.x;
  const bool backward = blockIdx.y == 1;
  const int Tb = min(input_lens[b], T);
  const int L = min(label_lens[b], L_max);
  const int S = 2 * L + 1;
  const int tid = threadIdx.x;

  for (int s = tid; s < S; s += kCtcThreads) ext[s] = (s & 1) ? labels[(size_t)b * L_max + (s >> 1)] : blank;
  for (int i = tid; i < 2 * LP; i += kCtcThreads) lat[i] = kNegBig;
  __syncthreads();
  // feasibility (ignore_longer_outputs_than_inputs=True): need L + repeats <= Tb
  __shared__ int repeats;
  if (tid == 0) {
    int r = 0;
    for (int i = 1; i < L; ++i) r += (ext[2 * i + 1] == ext[2 * i - 1]);
    repeats = r;
  }
  __syncthreads();
  if (Tb <= 0 || L + repeats > Tb) {
    if (tid == 0 && !backward) loglik[b] = CUDART_NAN_F;  // marks "skipped": loss 0, grad 0
    return;
  }

  float* out = (backward ? beta : alpha) + (size_t)b * T * S_max;
  const float* lg = logits + b * stride_b;
  const float* ls = lse + (size_t)b * T;

  // per-thread constants of the owned states
  const int d = backward ? 1 : -1;        // direction of the "previous" neighbours
  int cls[NS];
  bool own[NS], skip[NS], start[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = tid + j * kCtcThreads;
    own[j] = s < S;
    cls[j] = own[j] ? ext[s] : 0;
    const int s2 = s + 2 * d;
    skip[j] = own[j] && s2 >= 0 && s2 < S && cls[j] != blank && cls[j] != ext[s2];
    start[j] = own[j] && (backward ? (s >= S - 2) : (s <= 1));
  }

  auto stage = [&](int chunk, int buf) {
    // chunk covers steps [chunk*32, chunk*32+32) in processing order
    for (int i = tid; i < kCtcChunk * 32; i += kCtcThreads) {
      const int st = i >> 5, v = i & 31;
      const int step = chunk * kCtcChunk + st;
      float val = kNegBig;
      if (step < Tb && v < V) {
        const int t = backward ? (Tb - 1 - step) : step;
        val = lg[t * stride_t + v] - ls[t];
      }
      em[(buf * kCtcChunk + st) * 32 + v] = val;
    }
  };

  const int n_chunks = (Tb + kCtcChunk - 1) / kCtcChunk;
  stage(0, 0);
  __syncthreads();
  int cur = 0;
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    if (chunk + 1 < n_chunks) stage(chunk + 1, (chunk + 1) & 1);
    const float* emc = em + (chunk & 1) * kCtcChunk * 32;
    const int steps = min(kCtcChunk, Tb - chunk * kCtcChunk);
    for (int st = 0; st < steps; ++st) {
      const int step = chunk * kCtcChunk + st;
      const int t = backward ? (Tb - 1 - step) : step;
      const float* prev = lat + cur * LP + 2;
      float* next = lat + (cur ^ 1) * LP + 2;
      float* orow = out + (size_t)t * S_max;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (!own[j]) continue;
        const int s = tid + j * kCtcThreads;
        float a;
        if (step == 0) {
          a = start[j] ? 0.f : kNegBig;
        } else {
          const float a0 = prev[s];
          const float a1 = prev[s + d];
          const float a2 = skip[j] ? prev[s + 2 * d] : kNegBig;
          const float m = fmaxf(a0, fmaxf(a1, a2));
          a = m + __logf(__expf(a0 - m) + __expf(a1 - m) + __expf(a2 - m));
        }
        a = fmaxf(a + emc[st * 32 + cls[j]], kNegBig);
        next[s] = a;
        orow[s] = a;
      }
      cur ^= 1;
      __syncthreads();
    }
  }
  if (!backward && tid == 0) {
    const float* fin = lat + cur * LP + 2;
    float ll = fin[S - 1];
    if (S > 1) ll = log_add(ll, fin[S - 2]);
    loglik[b] = ll;
  }
}

// Step 3: gradient wrt logits and per-utterance loss. One warp per (b, t).
//   grad[b,t,v] = gscale * (softmax_v - sum_{s: ext[s]=v} exp(alpha+beta - logp_v - ll))   t < len
// loss[b] = -ll (0 when skipped / NaN, mask_nans); gscale = loss_scale / B (loss_scale read from device).
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ logits, const float* __restrict__ lse, const int* __restrict__ labels,
                const int* __restrict__ label_lens, const int* __restrict__ input_lens,
                const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ loglik,
                float* __restrict__ grad, float* __restrict__ loss, const float* __restrict__ loss_scale,
                int B, int T, int V, int L_max, int S_max, long long stride_b, long long stride_t,
                long long gstride_b, long long gstride_t, int blank) {
  __shared__ float acc[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= B * T) return;
  const int b = i / T, t = i - b * T;
  const float ll = loglik[b];
  const bool skipped = !(ll == ll) || ll <= kNegBig * 0.5f;  // NaN or no path
  const int Tb = min(input_lens[b], T);
  float* g = grad + b * gstride_b + t * gstride_t;
  if (t == 0 && lane == 0) loss[b] = skipped ? 0.f : -ll;
  if (skipped || t >= Tb) {
    if (lane < V) g[lane] = 0.f;
    return;
  }
  const int L = min(label_lens[b], L_max);
  const int S = 2 * L + 1;
  acc[warp][lane] = 0.f;
  __syncwarp();
  const float* row = logits + b * stride_b + t * stride_t;
  const float l = lse[(size_t)b * T + t];
  const float* al = alpha + ((size_t)b * T + t) * S_max;
  const float* be = beta + ((size_t)b * T + t) * S_max;
  for (int s = lane; s < S; s += 32) {
    const int c = (s & 1) ? labels[(size_t)b * L_max + (s >> 1)] : blank;
    const float lp = row[c] - l;
    const float e = __expf(al[s] + be[s] - lp - ll);
    atomicAdd(&acc[warp][c], e);
  }
  __syncwarp();
  if (lane < V) {
    const float gscale = (loss_scale ? *loss_scale : 1.f) / (float)B;
    const float sm = __expf(row[lane] - l);
    g[lane] = gscale * (sm - acc[warp][lane]);
  }
}

int ctc_loss_fwd_bwd(const float* logits, const int* labels, const int* label_lens, const int* input_lens,
                     float* grad, float* loss, float* workspace, size_t workspace_bytes,
                     const float* loss_scale, int B, int T, int V, int L_max, long long stride_b,
                     long long stride_t, cudaStream_t st) {
  if (V > 32) return fail(ERR_UNSUPPORTED, "ctc: vocabulary > 32 not supported");
  const int S_max = 2 * L_max + 1;
  // workspace: lse [B*T] | loglik [B] | alpha [B*T*S_max] | beta [B*T*S_max]
  const size_t need = ((size_t)B * T + B + 2 * (size_t)B * T * S_max) * sizeof(float);
  if (workspace_bytes < need) return fail(ERR_INVALID, "ctc: workspace too small, need " + std::to_string(need));
  float* lse = workspace;
  float* loglik = lse + (size_t)B * T;
  float* alpha = loglik + B;
  float* beta = alpha + (size_t)B * T * S_max;
  const int blank = V - 1;
  ctc_lse_kernel<<<(B * T + 255) / 256, 256, 0, st>>>(logits, lse, B, T, V, stride_b, stride_t);
  const size_t smem = (2 * (size_t)(S_max + 4) + 2 * kCtcChunk * 32) * sizeof(float) + (size_t)S_max * sizeof(int);
  if (smem > 200 * 1024 || S_max > 4 * kCtcThreads)
    return fail(ERR_UNSUPPORTED, "ctc: label sequence too long for the lattice kernel");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    OS2S_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  // states per thread: 1 (S <= 512), 2 (<= 1024) or 4 (<= 2048)
  if (S_max <= kCtcThreads)
    ctc_alpha_beta_kernel<1><<<dim3(B, 2), kCtcThreads, smem, st>>>(logits, lse, labels, label_lens, input_lens, alpha,
                                                                    beta, loglik, T, V, L_max, S_max, stride_b,
                                                                    stride_t, blank);
  else if (S_max <= 2 * kCtcThreads)
    ctc_alpha_beta_kernel<2><<<dim3(B, 2), kCtcThreads, smem, st>>>(logits, lse, labels, label_lens, input_lens, alpha,
                                                                    beta, loglik, T, V, L_max, S_max, stride_b,
                                                                    stride_t, blank);
  else
    ctc_alpha_beta_kernel<4><<<dim3(B, 2), kCtcThreads, smem, st>>>(logits, lse, labels, label_lens, input_lens, alpha,
                                                                    beta, loglik, T, V, L_max, S_max, stride_b,
                                                                    stride_t, blank);
  ctc_grad_kernel<<<(B * T + 7) / 8, 256, 0, st>>>(logits, lse, labels, label_lens, input_lens, alpha, beta, loglik,
                                                   grad, loss, loss_scale, B, T, V, L_max, S_max, stride_b, stride_t,
                                                   (long long)T * V, (long long)V, blank);
  return check_launch("ctc_loss_fwd_bwd");
}

size_t ctc_workspace_bytes(int B, int T, int L_max) {
  const int S_max = 2 * L_max + 1;
  return ((size_t)B * T + B + 2 * (size_t)B * T * S_max) * sizeof(float);
}

// ------------------------------------------------------------------ greedy decode
// One warp per utterance: argmax per frame (first max on ties), merge repeats, drop blank.
__global__ void ctc_greedy_kernel(const float* __restrict__ logits, const int* __restrict__ input_lens,
                                  int* __restrict__ tokens, int* __restrict__ out_lens, float* __restrict__ neg_sum,
                                  int B, int T, int V, long long stride_b, long long stride_t, int blank,
                                  int merge_repeated) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const int Tb = min(input_lens[b], T);
  int count = 0;
  int carry_prev = -1;
  float score = 0.f;
  for (int t0 = 0; t0 < Tb; t0 += 32) {
    const int t = t0 + lane;
    int c = -1;
    float best = 0.f;
    if (t < Tb) {
      const float* row = logits + b * stride_b + t * stride_t;
      best = row[0];
      c = 0;
      for (int v = 1; v < V; ++v) {
        const float x = row[v];
        if (x > best) {
          best = x;
          c = v;
        }
      }
    }
    int prev = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) prev = carry_prev;
    const bool emit = (t < Tb) && (c != blank) && !(merge_repeated && c == prev);
    const unsigned mask = __ballot_sync(0xffffffffu, emit);
    const int pos = count + __popc(mask & ((1u << lane) - 1u));
    if (emit) tokens[(size_t)b * T + pos] = c;
    count += __popc(mask);
    const int last = min(31, Tb - 1 - t0);
    carry_prev = __shfl_sync(0xffffffffu, c, last);
    float sc = (t < Tb) ? best : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
    score += sc;
  }
  if (lane == 0) {
    out_lens[b] = count;
    if (neg_sum) neg_sum[b] = -score;
  }
}

int ctc_greedy(const float* logits, const int* input_lens, int* tokens, int* out_lens, float* neg_sum, int B,
               int T, int V, long long stride_b, long long stride_t, int merge_repeated, cudaStream_t st) {
  ctc_greedy_kernel<<<B, 32, 0, st>>>(logits, input_lens, tokens, out_lens, neg_sum, B, T, V, stride_b, stride_t,
                                      V - 1, merge_repeated);
  return check_launch("ctc_greedy");
}

}  // namespace os2s
