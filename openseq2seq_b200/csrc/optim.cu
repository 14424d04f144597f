// Fused multi-tensor optimizer step: unscale -> (gradients already all-reduced) -> LARC -> NaN/Inf
// check -> Backoff loss-scaler update -> NovoGrad (on TF Momentum) / Momentum / Adam -> fp32 master ->
// bf16 copies.
//
// Reference chain being replaced (hundreds of tiny TF ops per step):
//   open_seq2seq/optimizers/mp_wrapper.py:44-122          loss scale, fp32 masters, skip on overflow
//   open_seq2seq/optimizers/optimizers.py:333-377         LARC
//   open_seq2seq/optimizers/automatic_loss_scaler.py:31-106  check_grads + BackoffScaler
//   open_seq2seq/optimizers/novograd.py:93-126            NovoGrad (+ tf.train.MomentumOptimizer)
//   open_seq2seq/optimizers/lr_policies.py:15-170         fixed_lr, exp_decay, poly_decay, cosine_decay
//   tf.train.AdamOptimizer (optimizers.py:36-44 "Adam")   m, v moments with the lr_t bias correction
// Everything (loss scale, step counters, learning rate, skip decision) lives in device memory, so a
// training step never synchronises with the host.
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>

namespace os2s {

constexpr int kOptChunk = 16384;  // elements per CTA
constexpr int kOptThreads = 256;

// Pass 1: per-tensor sum g^2, sum w^2 and a global non-finite flag.
__global__ void __launch_bounds__(kOptThreads)
opt_norms_kernel(const OptTable tab, const OptHParams hp, const float* __restrict__ fstate,
                 float* __restrict__ norms, int* __restrict__ nonfinite) {
  const int tid = tab.chunk_tensor[blockIdx.x];
  // L2 regulariser (mp_wrapper.py:81-89: its gradient reg*w is added to the UNSCALED fp32 gradient before
  // LARC): in units of the scaled, rank-summed gradient G that is G + (reg * loss_scale * world) * w
  const float kreg = tab.reg ? tab.reg[tid] * fstate[0] * (float)hp.world_size : 0.f;
  const long long off = tab.chunk_offset[blockIdx.x];
  const long long n = tab.sizes[tid];
  const float* g = reinterpret_cast<const float*>(tab.g[tid]) + off;
  const float* w = reinterpret_cast<const float*>(tab.w[tid]) + off;
  const int len = (int)min((long long)kOptChunk, n - off);
  float sg = 0.f, sw = 0.f;
  bool bad = false;
  const int len4 = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(w)) & 15) == 0 ? (len & ~3) : 0;
  for (int i = threadIdx.x * 4; i < len4; i += kOptThreads * 4) {
    float4 gv = *reinterpret_cast<const float4*>(g + i);
    const float4 wv = *reinterpret_cast<const float4*>(w + i);
    bad |= !(isfinite(gv.x) && isfinite(gv.y) && isfinite(gv.z) && isfinite(gv.w));
    gv.x += kreg * wv.x; gv.y += kreg * wv.y; gv.z += kreg * wv.z; gv.w += kreg * wv.w;
    sg += gv.x * gv.x + gv.y * gv.y + gv.z * gv.z + gv.w * gv.w;
    sw += wv.x * wv.x + wv.y * wv.y + wv.z * wv.z + wv.w * wv.w;
  }
  for (int i = len4 + threadIdx.x; i < len; i += kOptThreads) {
    const float wv = w[i];
    float gv = g[i];
    bad |= !isfinite(gv);
    gv += kreg * wv;
    sg += gv * gv;
    sw += wv * wv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  __shared__ float shg[kOptThreads / 32], shw[kOptThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    shg[warp] = sg;
    shw[warp] = sw;
  }
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < kOptThreads / 32; ++i) {
      a += shg[i];
      b += shw[i];
    }
    atomicAdd(&norms[2 * tid], a);
    atomicAdd(&norms[2 * tid + 1], b);
    if (anybad) atomicOr(nonfinite, 1);
  }
}

// Pass 2 (one CTA): scaler / step / lr bookkeeping and the per-tensor gradient coefficient.
//   fstate: [0] loss_scale (used by the NEXT backward)  [1] lr of this step  [2] global grad norm
//   istate: [0] scaler iteration  [1] last overflow iteration  [2] global_step  [3] skip flag (this step)
//           [4] number of skipped steps so far  [5] attempted steps
__global__ void opt_prepare_kernel(const OptTable tab, const OptHParams hp, float* __restrict__ norms,
                                   int* __restrict__ nonfinite, float* __restrict__ fstate,
                                   long long* __restrict__ istate, float* __restrict__ coef,
                                   float* __restrict__ ema) {
  __shared__ float s_scale_used, s_lr;
  __shared__ int s_skip;
  if (threadIdx.x == 0) {
    const float scale_used = fstate[0];
    const bool overflow = (*nonfinite) != 0;
    long long iteration = istate[0], last_of = istate[1], step = istate[2];
    // lr = lr_policy(global_step) evaluated before the step is applied (optimizers.py:172)
    float lr0 = hp.lr0;
    if (hp.warmup_steps > 0 && step < hp.warmup_steps) lr0 = lr0 * (float)step / (float)hp.warmup_steps;
    float lr = lr0;
    if (hp.lr_policy == 2) {
      // exp_decay (lr_policies.py:55-92): no warm-up, tf.train.exponential_decay, floor at min_lr
      lr = hp.lr0;
      if (step >= hp.begin_decay_at && hp.decay_steps > 0) {
        float e = (float)(step - hp.begin_decay_at) / (float)hp.decay_steps;
        if (hp.staircase) e = floorf(e);
        lr = hp.lr0 * powf(hp.decay_rate, e);
      }
      lr = fmaxf(lr, hp.min_lr);
    } else if (hp.lr_policy == 3) {
      lr = hp.lr0;  // fixed_lr
    } else if (step >= hp.begin_decay_at && hp.decay_steps > 0) {
      const long long s = min(step - hp.begin_decay_at, hp.decay_steps);
      if (hp.lr_policy == 1) {
        // tf.train.cosine_decay with alpha = min_lr (as the reference calls it, lr_policies.py:160-166)
        const float cosd = 0.5f * (1.f + cospif((float)s / (float)hp.decay_steps));
        lr = lr0 * ((1.f - hp.min_lr) * cosd + hp.min_lr);
      } else {
        lr = (lr0 - hp.min_lr) * powf(1.f - (float)s / (float)hp.decay_steps, hp.power) + hp.min_lr;
      }
    }
    // BackoffScaler.update_op (automatic_loss_scaler.py:78-106)
    float scale = scale_used;
    if (hp.use_loss_scaler) {
      if (overflow) {
        scale = fminf(fmaxf(scale / hp.step_factor, hp.scale_min), hp.scale_max);
        last_of = iteration;
      } else if (((iteration - last_of) % hp.step_window) == 0) {
        scale = fminf(fmaxf(scale * hp.step_factor, hp.scale_min), hp.scale_max);
      }
      iteration += 1;
    }
    const bool skip = overflow;  // mp_wrapper.py:115-120
    fstate[0] = scale;
    fstate[1] = lr;
    fstate[4] = scale_used;  // the update kernel needs the scale this step's gradients carry
    if (hp.algo == 2) {
      // tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), t = applied steps incl. this one
      const float t = (float)(step + 1);
      fstate[3] = lr * sqrtf(1.f - powf(hp.beta2, t)) / (1.f - powf(hp.beta1, t));
    }
    istate[0] = iteration;
    istate[1] = last_of;
    istate[2] = skip ? step : step + 1;
    istate[3] = skip ? 1 : 0;
    if (skip) istate[4] += 1;
    istate[5] += 1;  // attempted steps (drives the dropout stream)
    s_scale_used = scale_used;
    s_lr = lr;
    s_skip = skip ? 1 : 0;
  }
  __syncthreads();
  const float unscale = 1.f / (s_scale_used * (float)hp.world_size);  // mp_wrapper.py:94, hvd mean
  const float lr = s_lr;
  // global gradient norm over the trainable variables (summary, optimizers.py:292-296, and
  // tf.clip_by_global_norm when max_grad_norm is set, optimizers.py:408-433)
  float total = 0.f;
  for (int t = threadIdx.x; t < tab.n_tensors; t += blockDim.x)
    if (!(tab.frozen && tab.frozen[t])) total += norms[2 * t] * unscale * unscale;
  __shared__ float red[32];
  __shared__ float s_clip;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    const float gn = sqrtf(s);
    fstate[2] = gn;
    s_clip = (hp.max_grad_norm > 0.f) ? hp.max_grad_norm / fmaxf(gn, hp.max_grad_norm) : 1.f;
    *nonfinite = 0;
  }
  __syncthreads();
  const float clip = s_clip;
  for (int t = threadIdx.x; t < tab.n_tensors; t += blockDim.x) {
    const float g_norm = sqrtf(norms[2 * t]) * unscale * clip;
    const float w_norm = sqrtf(norms[2 * t + 1]);
    float r = 1.f;
    if (hp.larc_eta > 0.f) {  // optimizers.py:349-369
      if (hp.larc_mode == 0) {
        r = fmaxf(hp.larc_eta * w_norm / (lr * (g_norm + hp.larc_eps)), hp.larc_min_update);
        r = fminf(r, 1.f);
      } else {
        r = fmaxf(hp.larc_eta * w_norm / (g_norm + hp.larc_eps), hp.larc_min_update);
      }
    }
    float c = unscale * clip * r;
    if (tab.frozen && tab.frozen[t]) c = 0.f;
    if (hp.algo == 0) {  // NovoGrad (novograd.py:108-115)
      const float g2 = (r * g_norm) * (r * g_norm);
      const float prev = ema[t];
      const float v = (prev == 0.f) ? g2 : prev * hp.beta2 + g2 * (1.f - hp.beta2);
      if (hp.ema_persist && !s_skip && !(tab.frozen && tab.frozen[t])) ema[t] = v;
      c *= rsqrtf(v + hp.epsilon);
      if (hp.grad_averaging) c *= (1.f - hp.beta1);
    }
    coef[t] = c;
    norms[2 * t] = 0.f;
    norms[2 * t + 1] = 0.f;
  }
}

// Pass 3: the update.  g_hat = coef * G + wd * w (weight-decay term scaled by (1-beta1) when
// grad_averaging, novograd.py:117-121); m = beta1*m + g_hat; w -= lr*m; bf16 copy.
__global__ void __launch_bounds__(kOptThreads)
opt_update_kernel(const OptTable tab, const OptHParams hp, const float* __restrict__ fstate,
                  const long long* __restrict__ istate, const float* __restrict__ coef) {
  if (istate[3] != 0) return;  // overflow: skip the whole step
  const int tid = tab.chunk_tensor[blockIdx.x];
  if (tab.frozen && tab.frozen[tid]) return;  // freeze_variables_regex: not in var_list (model.py:502-507)
  const long long off = tab.chunk_offset[blockIdx.x];
  const long long n = tab.sizes[tid];
  const int len = (int)min((long long)kOptChunk, n - off);
  const float* g = reinterpret_cast<const float*>(tab.g[tid]) + off;
  float* w = reinterpret_cast<float*>(tab.w[tid]) + off;
  float* m = reinterpret_cast<float*>(tab.m[tid]) + off;
  uint16_t* wb = tab.wb[tid] ? reinterpret_cast<uint16_t*>(tab.wb[tid]) + off : nullptr;
  auto half16 = [&](float v) -> uint16_t {
    return hp.wb_f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16(v));
  };
  const float c = coef[tid];
  const float lr = fstate[1];
  const float kreg = tab.reg ? tab.reg[tid] * fstate[4] * (float)hp.world_size : 0.f;
  const float wd = hp.weight_decay * ((hp.algo == 0 && hp.grad_averaging) ? (1.f - hp.beta1) : 1.f);
  const float mom = (hp.algo == 0) ? hp.beta1 : hp.momentum;
  if (hp.algo == 2) {
    float* v = reinterpret_cast<float*>(tab.v[tid]) + off;
    const float lr_t = fstate[3];
    for (int i = threadIdx.x; i < len; i += kOptThreads) {
      const float wv = w[i];
      const float gh = c * (g[i] + kreg * wv) + wd * wv;
      const float mv = hp.beta1 * m[i] + (1.f - hp.beta1) * gh;
      const float vv = hp.beta2 * v[i] + (1.f - hp.beta2) * gh * gh;
      const float nw = wv - lr_t * mv / (sqrtf(vv) + hp.epsilon);
      m[i] = mv;
      v[i] = vv;
      w[i] = nw;
      if (wb) wb[i] = half16(nw);
    }
    return;
  }
  for (int i = threadIdx.x; i < len; i += kOptThreads) {
    const float wv = w[i];
    const float gh = c * (g[i] + kreg * wv) + wd * wv;
    const float mv = mom * m[i] + gh;
    const float nw = wv - lr * mv;
    m[i] = mv;
    w[i] = nw;
    if (wb) wb[i] = half16(nw);
  }
}

int opt_step(const OptTable& tab, const OptHParams& hp, float* norms, int* nonfinite, float* fstate,
             long long* istate, float* coef, float* ema, cudaStream_t st) {
  if (tab.n_tensors <= 0 || tab.n_chunks <= 0) return fail(ERR_INVALID, "opt_step: empty table");
  opt_norms_kernel<<<tab.n_chunks, kOptThreads, 0, st>>>(tab, hp, fstate, norms, nonfinite);
  opt_prepare_kernel<<<1, 1024, 0, st>>>(tab, hp, norms, nonfinite, fstate, istate, coef, ema);
  opt_update_kernel<<<tab.n_chunks, kOptThreads, 0, st>>>(tab, hp, fstate, istate, coef);
  return check_launch("opt_step");
}

int opt_chunk_elems() { return kOptChunk; }

// ------------------------------------------------------------- multi-tensor bf16 transpose
// For every conv kernel: wt[k][c][r] = w[k][r][c]  (bf16 -> bf16), one launch for all tensors.
__global__ void multi_transpose_kernel(const TransposeTable tab) {
  __shared__ __nv_bfloat16 tile[32][33];
  // locate tensor by binary search over the tile prefix sums
  int lo = 0, hi = tab.n_tensors - 1;
  const long long bid = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.tile_start[mid] <= bid) lo = mid; else hi = mid - 1;
  }
  const int t = lo;
  const int R = tab.R[t], C = tab.C[t];
  const int tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
  long long local = bid - tab.tile_start[t];
  const int k = (int)(local / ((long long)tiles_r * tiles_c));
  local -= (long long)k * tiles_r * tiles_c;
  const int tr = (int)(local / tiles_c), tc = (int)(local - (long long)tr * tiles_c);
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(tab.src[t]) + (size_t)k * R * C;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(tab.dst[t]) + (size_t)k * R * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = tr * 32 + i, c = tc * 32 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = src[(size_t)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = tc * 32 + i, r = tr * 32 + threadIdx.x;
    if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

int multi_transpose(const TransposeTable& tab, long long total_tiles, cudaStream_t st) {
  if (total_tiles <= 0) return OK;
  multi_transpose_kernel<<<(unsigned)total_tiles, dim3(32, 8), 0, st>>>(tab);
  return check_launch("multi_transpose");
}

}  // namespace os2s
