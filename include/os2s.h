/* libos2s_b200 -- C ABI of the B200-native Jasper speech-to-text training path.
 *
 * The reference (NVIDIA/OpenSeq2Seq) has no C ABI: its plugin interface is a Python class protocol
 * (DataLayer / Encoder / Decoder / Loss / optimize_loss) that builds TensorFlow-1 graph ops.  This
 * header is the boundary one level below that protocol: every entry point replaces the TF op (or
 * chain of ops) named in its comment, and is what the Python plugin classes in
 * openseq2seq_b200/ bind through ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer
 *   - no hidden allocation, no hidden synchronisation; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream)
 *   - return value: 0 = OK, negative = error; os2s_last_error() gives the message (thread local)
 *   - activations are NWC bf16 [B, T, C] with C contiguous; parameters/gradients are fp32 masters
 *     with bf16 working copies written by the optimizer step
 */
#ifndef OS2S_H_
#define OS2S_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OS2S_OK 0
#define OS2S_ERR_INVALID (-1)
#define OS2S_ERR_CUDA (-2)
#define OS2S_ERR_UNSUPPORTED (-3)
#define OS2S_ERR_NCCL (-4)

/* Output modes of the conv epilogue. */
#define OS2S_OUT_BF16 0
#define OS2S_OUT_F32 1
#define OS2S_OUT_F32_ACC 2 /* out(fp32) += result */

const char* os2s_last_error(void);
int os2s_version(void);

/* ---- K2: tf.layers.conv1d(use_bias=False, padding=SAME), stride 1 -----------------------------
 * reference: open_seq2seq/parts/cnns/conv_blocks.py:195-206 (main), :79-85 (1x1 residual).
 * y[b,t,o] = sum_k sum_c x[b, t - pad_left + k*dil, c] * W[k,c,o]      (zero outside [0,T))
 *   x  : bf16 [B,T,C_in]          wt : bf16 [K][C_out][C_in]  (transposed working copy)
 *   y  : bf16 / fp32 [B,T,C_out]  (out_mode)
 * The stride-2 first Jasper layer is expressed by the caller as a stride-1 conv over the input
 * viewed as [B, T/2, 2*C_in] with K' = ceil(K/2) taps (see openseq2seq_b200/runtime/layers.py).
 * Constraints: C_in % 64 == 0, C_out % 64 == 0. */
int os2s_conv1d_fwd(const void* x, const void* wt, void* y, int B, int T, int C_in, int C_out,
                    int K, int dil, int pad_left, int out_mode, void* stream);

/* dgrad of the above: dx[b,t,c] = sum_k sum_o dy[b, t + pad_left - k*dil, o] * W[k,c,o]
 *   dy : bf16 [B,T,C_out]   w : bf16 [K][C_in][C_out] (natural TF layout)   dx : out_mode */
int os2s_conv1d_dgrad(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, int out_mode, void* stream);

/* wgrad: dw[k,c,o] = sum_{b,t} x[b, t - pad_left + k*dil, c] * dy[b,t,o]   (fp32 [K][C_in][C_out])
 * Overwrites dw.  Constraints: C_in % 128 == 0, C_out % 64 == 0. */
int os2s_conv1d_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, void* stream);

/* W fp32 [K][C_in][C_out] -> w bf16 (same layout) and wt bf16 [K][C_out][C_in]. Either output may
 * be NULL.  Replaces the fp32->fp16 assign of mp_wrapper.py:104-109 when used standalone. */
int os2s_weight_cast_transpose(const float* w_master, void* w_bf16, void* wt_bf16, int K, int C_in,
                               int C_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OS2S_H_ */
