// HBM-bound elementwise / reduction kernels of the Jasper path.
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>

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

}  // namespace os2s
