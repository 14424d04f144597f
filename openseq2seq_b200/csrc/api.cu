// extern "C" boundary of libos2s_b200 (declared in include/os2s.h).
#include "../../include/os2s.h"

#include "common.h"
#include "kernels.h"

using namespace os2s;

extern "C" {

const char* os2s_last_error(void) { return last_error_cstr(); }
int os2s_version(void) { return 100; }

int os2s_conv1d_fwd(const void* x, const void* wt, void* y, int B, int T, int C_in, int C_out,
                    int K, int dil, int pad_left, int out_mode, void* stream) {
  if (!x || !wt || !y) return fail(ERR_INVALID, "os2s_conv1d_fwd: null pointer");
  return conv_kmajor(x, wt, y, B, T, C_in, C_out, K, -pad_left, dil, out_mode, (cudaStream_t)stream);
}

int os2s_conv1d_dgrad(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, int out_mode, void* stream) {
  if (!dy || !w || !dx) return fail(ERR_INVALID, "os2s_conv1d_dgrad: null pointer");
  return conv_kmajor(dy, w, dx, B, T, C_out, C_in, K, pad_left, -dil, out_mode, (cudaStream_t)stream);
}

int os2s_conv1d_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, void* stream) {
  if (!x || !dy || !dw) return fail(ERR_INVALID, "os2s_conv1d_wgrad: null pointer");
  return conv_wgrad(x, dy, dw, B, T, C_in, C_out, K, dil, pad_left, nullptr, (cudaStream_t)stream);
}

int os2s_weight_cast_transpose(const float* w_master, void* w_bf16, void* wt_bf16, int K, int C_in,
                               int C_out, void* stream) {
  if (!w_master) return fail(ERR_INVALID, "os2s_weight_cast_transpose: null pointer");
  return weight_cast_transpose(w_master, w_bf16, wt_bf16, K, C_in, C_out, (cudaStream_t)stream);
}

}  // extern "C"
