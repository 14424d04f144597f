// Internal (C++) declarations of the kernel launchers behind include/os2s.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace os2s {

const char* last_error_cstr();

// conv_tc.cu
int conv_kmajor(const void* act, const void* wmat, void* out, int B, int T, int C_red, int N_total,
                int K, int t_off0, int t_step, int out_mode, cudaStream_t st);
int conv_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out, int K,
               int dil, int pad_left, int* splits_used, cudaStream_t st);

// elementwise.cu
int weight_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int K, int C_in, int C_out,
                          cudaStream_t st);

}  // namespace os2s
