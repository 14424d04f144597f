// Separable 1-D convolution (QuartzNet / Jasper-Mini): tf.layers.separable_conv1d(use_bias=False,
// padding=SAME) as called from open_seq2seq/parts/cnns/conv_blocks.py:27-40,180-193 (main conv) and :79-85
// (the residual branch of a sep_conv1d block is ALSO a separable conv, with kernel_size 1).
//
//   z[b,t,c] = sum_k D[k,c] * x[b, t*s - pad + k*dil, c]        depthwise_kernel D [K, C_in, 1]
//   y[b,t,o] = sum_c z[b,t,c] * P[c,o]                          pointwise_kernel P [1, C_in, C_out]
//
// Two realisations:
//  * split (wide stride-1 layers, C = 256 ... 1024, K = 11 ... 87): the depthwise stage runs on the CUDA cores
//    (this file: 2 K FLOP per 4 bytes moved, shared-memory tiles with a time halo, fp32 accumulation), the
//    pointwise stage is the existing 1x1 tcgen05 GEMM (conv_tc.cu) incl. its fused BN statistics;
//  * composed (K = 1, i.e. a per-channel scale in front of a 1x1 conv -- every residual branch -- and the
//    stride-2 first layer on the 64 features): the dense kernel W[k,c,o] = D[k,c] * P[c,o] is formed on the fly
//    and the layer runs as an ordinary dense convolution; its weight gradient dW is folded back,
//      dD[k,c] = sum_o dW[k,c,o] * P[c,o],   dP[c,o] = sum_k dW[k,c,o] * D[k,c].
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace os2s {

__device__ __forceinline__ float h16_to_f(uint16_t v, int f16) {
  return f16 ? __half2float(__ushort_as_half(v)) : __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ uint16_t f_to_h16(float v, int f16) {
  return f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16(v));
}
__device__ __forceinline__ void h16x8_to_f(const uint4& v, float (&f)[8], int f16) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (f16) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    } else {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
}

// ---------------------------------------------------------------------------- composed kernels
// w[k,c,o] = D[k,c] * P[c,o]  (16-bit working copy of the equivalent dense kernel)
__global__ void sep_compose_kernel(const float* __restrict__ D, const float* __restrict__ P, uint16_t* __restrict__ w,
                                   int K, int C, int Co, int f16) {
  const long long n = (long long)K * C * Co;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % Co);
    const long long kc = i / Co;
    const int c = (int)(kc % C);
    w[i] = f_to_h16(D[kc] * P[(size_t)c * Co + o], f16);
  }
}
int sepconv_compose(const float* D, const float* P, void* w, int K, int C, int Co, int f16, cudaStream_t st) {
  if (K <= 0 || C <= 0 || Co <= 0) return fail(ERR_INVALID, "sepconv_compose: bad shape");
  const long long n = (long long)K * C * Co;
  int grid = (int)((n + 255) / 256);
  if (grid > 4 * device_sm_count()) grid = 4 * device_sm_count();
  sep_compose_kernel<<<grid, 256, 0, st>>>(D, P, (uint16_t*)w, K, C, Co, f16);
  return check_launch("sep_compose");
}

// dD[k,c] = sum_o dW[k,c,o] * P[c,o] : one warp per (k, c)
__global__ void sep_grad_depthwise_kernel(const float* __restrict__ dW, const float* __restrict__ P,
                                          float* __restrict__ dD, int K, int C, int Co) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= K * C) return;
  const int c = warp % C;
  const float* row = dW + (size_t)warp * Co;
  const float* prow = P + (size_t)c * Co;
  float s = 0.f;
  for (int o = lane; o < Co; o += 32) s += row[o] * prow[o];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dD[warp] = s;
}
// dP[c,o] = sum_k dW[k,c,o] * D[k,c] : one thread per (c, o)
__global__ void sep_grad_pointwise_kernel(const float* __restrict__ dW, const float* __restrict__ D,
                                          float* __restrict__ dP, int K, int C, int Co) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)C * Co) return;
  const int c = (int)(i / Co);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += dW[(size_t)k * C * Co + i] * D[(size_t)k * C + c];
  dP[i] = s;
}
int sepconv_decompose_grad(const float* dW, const float* D, const float* P, float* dD, float* dP, int K, int C, int Co,
                           cudaStream_t st) {
  if (K <= 0 || C <= 0 || Co <= 0) return fail(ERR_INVALID, "sepconv_decompose_grad: bad shape");
  sep_grad_depthwise_kernel<<<(K * C * 32 + 255) / 256, 256, 0, st>>>(dW, P, dD, K, C, Co);
  sep_grad_pointwise_kernel<<<(int)(((long long)C * Co + 255) / 256), 256, 0, st>>>(dW, D, dP, K, C, Co);
  return check_launch("sep_decompose_grad");
}

// ---------------------------------------------------------------------------- depthwise conv (fwd / dgrad)
// out[b,t,c] = sum_k taps[k,c] * x[b, t + off0 + k*step, c]   (zero outside [0,T))
//   forward : off0 = -pad_left, step = +dilation;   data gradient: off0 = +pad_left, step = -dilation
// CTA = 128 output rows x 64 channels of one utterance; thread = 8 channels (one 16-byte vector) x 4 rows.
constexpr int kDwRows = 128, kDwCh = 64, kDwThreads = 256;
enum DwOut : int { DW_OUT_HALF = 0, DW_OUT_F32 = 1, DW_OUT_F32_ACC = 2 };

__global__ void __launch_bounds__(kDwThreads)
depthwise_kernel(const uint16_t* __restrict__ x, const float* __restrict__ taps, void* __restrict__ out, int T, int C,
                 int K, int off0, int step, int out_mode, int f16) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int span = (K - 1) * (step < 0 ? -step : step);
  const int n_rows = kDwRows + span;
  uint4* xs = reinterpret_cast<uint4*>(smem);                         // [n_rows][8] 16-byte vectors
  float* ws = reinterpret_cast<float*>(smem + (size_t)n_rows * kDwCh * 2);   // [K][64]
  const int t0 = blockIdx.x * kDwRows, c0 = blockIdx.y * kDwCh, b = blockIdx.z;
  const int lo = t0 + off0 + (step < 0 ? (K - 1) * step : 0);         // first input row of the tile
  const uint16_t* xb = x + ((size_t)b * T) * C + c0;
  for (int i = threadIdx.x; i < n_rows * 8; i += kDwThreads) {
    const int r = i >> 3, v = i & 7;
    const int row = lo + r;
    xs[i] = (row >= 0 && row < T) ? __ldg(reinterpret_cast<const uint4*>(xb + (size_t)row * C) + v) : make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < K * kDwCh; i += kDwThreads) ws[i] = taps[(size_t)(i / kDwCh) * C + c0 + (i % kDwCh)];
  __syncthreads();
  const int v = threadIdx.x & 7, rg = threadIdx.x >> 3;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  const int base = off0 - (lo - t0);                                  // smem row of (output row 0, tap 0)
  for (int k = 0; k < K; ++k) {
    float w8[8];
    *reinterpret_cast<float4*>(&w8[0]) = *reinterpret_cast<const float4*>(ws + k * kDwCh + v * 8);
    *reinterpret_cast<float4*>(&w8[4]) = *reinterpret_cast<const float4*>(ws + k * kDwCh + v * 8 + 4);
    const int r0 = rg * 4 + base + k * step;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
      h16x8_to_f(xs[(r0 + j) * 8 + v], f, f16);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] += w8[i] * f[i];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int t = t0 + rg * 4 + j;
    if (t >= T) break;
    const size_t o = ((size_t)b * T + t) * C + c0 + v * 8;
    if (out_mode == DW_OUT_HALF) {
      uint4 pk;
      uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pw[i] = (uint32_t)f_to_h16(acc[j][2 * i], f16) | ((uint32_t)f_to_h16(acc[j][2 * i + 1], f16) << 16);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + o) = pk;
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o);
      float4 a0 = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      float4 a1 = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
      if (out_mode == DW_OUT_F32_ACC) {
        const float4 o0 = dst[0], o1 = dst[1];
        a0.x += o0.x; a0.y += o0.y; a0.z += o0.z; a0.w += o0.w;
        a1.x += o1.x; a1.y += o1.y; a1.z += o1.z; a1.w += o1.w;
      }
      dst[0] = a0;
      dst[1] = a1;
    }
  }
}

static size_t dw_smem(int K, int step) {
  const int span = (K - 1) * (step < 0 ? -step : step);
  return (size_t)(kDwRows + span) * kDwCh * 2 + (size_t)K * kDwCh * 4;
}

int depthwise_conv1d(const void* x, const float* taps, void* out, int B, int T, int C, int K, int off0, int step,
                     int out_mode, int f16, cudaStream_t st) {
  if (C % kDwCh != 0) return fail(ERR_UNSUPPORTED, "depthwise_conv1d: channels must be a multiple of 64");
  if (B <= 0 || T <= 0 || K <= 0 || step == 0) return fail(ERR_INVALID, "depthwise_conv1d: bad shape");
  if (out_mode < 0 || out_mode > 2) return fail(ERR_INVALID, "depthwise_conv1d: bad output mode");
  const size_t smem = dw_smem(K, step);
  if (smem > 200 * 1024) return fail(ERR_UNSUPPORTED, "depthwise_conv1d: kernel span too large for one shared-memory tile");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(depthwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  dim3 grid((T + kDwRows - 1) / kDwRows, C / kDwCh, B);
  depthwise_kernel<<<grid, kDwThreads, smem, st>>>((const uint16_t*)x, taps, out, T, C, K, off0, step, out_mode, f16);
  return check_launch("depthwise_conv1d");
}

// ---------------------------------------------------------------------------- depthwise weight gradient
// dtaps[k,c] += sum_{b,t} x[b, t - pad + k*dil, c] * dz[b,t,c]
// CTA = one utterance x 64 channels, looping over 128-row tiles; thread = 8 channels x up to 3 taps
// (k = g, g + 32, g + 64); the per-CTA sums go out with one atomicAdd per (k, c).
constexpr int kDwMaxTaps = 96;
__global__ void __launch_bounds__(kDwThreads)
depthwise_wgrad_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dz, float* __restrict__ dtaps,
                       int T, int C, int K, int dil, int pad, int f16) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int span = (K - 1) * dil;
  const int n_rows = kDwRows + span;
  uint4* xs = reinterpret_cast<uint4*>(smem);
  uint4* zs = xs + (size_t)n_rows * 8;
  const int c0 = blockIdx.x * kDwCh, b = blockIdx.y;
  const uint16_t* xb = x + ((size_t)b * T) * C + c0;
  const uint16_t* zb = dz + ((size_t)b * T) * C + c0;
  const int v = threadIdx.x & 7, g = threadIdx.x >> 3;
  float acc[3][8];
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
  for (int t0 = 0; t0 < T; t0 += kDwRows) {
    __syncthreads();
    const int lo = t0 - pad;
    for (int i = threadIdx.x; i < n_rows * 8; i += kDwThreads) {
      const int r = i >> 3, vv = i & 7;
      const int row = lo + r;
      xs[i] = (row >= 0 && row < T) ? __ldg(reinterpret_cast<const uint4*>(xb + (size_t)row * C) + vv) : make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < kDwRows * 8; i += kDwThreads) {
      const int r = i >> 3, vv = i & 7;
      const int row = t0 + r;
      zs[i] = (row < T) ? __ldg(reinterpret_cast<const uint4*>(zb + (size_t)row * C) + vv) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    for (int r = 0; r < kDwRows; ++r) {
      float d8[8];
      h16x8_to_f(zs[r * 8 + v], d8, f16);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int k = g + 32 * q;
        if (k < K) {
          float f[8];
          h16x8_to_f(xs[(r + k * dil) * 8 + v], f, f16);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[q][i] += f[i] * d8[i];
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int k = g + 32 * q;
    if (k < K) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&dtaps[(size_t)k * C + c0 + v * 8 + i], acc[q][i]);
    }
  }
}

int depthwise_conv1d_wgrad(const void* x, const void* dz, float* dtaps, int B, int T, int C, int K, int dil, int pad,
                           int f16, cudaStream_t st) {
  if (C % kDwCh != 0) return fail(ERR_UNSUPPORTED, "depthwise_conv1d_wgrad: channels must be a multiple of 64");
  if (K > kDwMaxTaps || K <= 0) return fail(ERR_UNSUPPORTED, "depthwise_conv1d_wgrad: 1..96 taps");
  if (B <= 0 || T <= 0 || dil <= 0) return fail(ERR_INVALID, "depthwise_conv1d_wgrad: bad shape");
  const size_t smem = (size_t)(kDwRows + (K - 1) * dil) * kDwCh * 2 + (size_t)kDwRows * kDwCh * 2;
  if (smem > 200 * 1024) return fail(ERR_UNSUPPORTED, "depthwise_conv1d_wgrad: kernel span too large");
  static bool attr_done = false;
  if (!attr_done) {
    OS2S_CUDA(cudaFuncSetAttribute(depthwise_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  OS2S_CUDA(cudaMemsetAsync(dtaps, 0, (size_t)K * C * sizeof(float), st));
  depthwise_wgrad_kernel<<<dim3(C / kDwCh, B), kDwThreads, smem, st>>>((const uint16_t*)x, (const uint16_t*)dz, dtaps, T, C, K,
                                                                      dil, pad, f16);
  return check_launch("depthwise_conv1d_wgrad");
}

}  // namespace os2s
