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
 *   - activations are NWC 16-bit [B, T, C] with C contiguous (bf16, or fp16 with OS2S_HALF_F16);
 *     parameters/gradients are fp32 masters with 16-bit working copies written by the optimizer step
 *   - rendezvous and the initial broadcast (utils/hooks.py:15-55) are the caller's business (torch.distributed,
 *     openseq2seq_b200/dist.py).  The gradient sum (hvd.allreduce, optimizers/optimizers.py:77-104) is either
 *     the caller's NCCL all-reduce or os2s_peer_* below (CUDA IPC + copy engines over NVLink, one node); the
 *     optimizer only requires that `g` holds the rank-summed gradients when os2s_opt_step* runs
 *     (os2s_opt_hparams.world_size folds the 1/N of the mean into the unscale factor)
 */
#ifndef OS2S_H_
#define OS2S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OS2S_OK 0
#define OS2S_ERR_INVALID (-1)
#define OS2S_ERR_CUDA (-2)
#define OS2S_ERR_UNSUPPORTED (-3)

/* Output modes of the conv epilogue. */
#define OS2S_OUT_BF16 0
#define OS2S_OUT_F32 1
#define OS2S_OUT_F32_ACC 2 /* out(fp32) += result */
#define OS2S_OUT_F16 3     /* fp16, saturating: conv outputs that only feed the BN kernels */
#define OS2S_OUT_F16_GRAD 4 /* fp16, overflow -> inf: gradients in OS2S_HALF_F16 mode */

/* Storage formats: bit flags of the `dtypes` argument of the *_p entry points (the entry points without the
 * suffix are the same calls with dtypes = 0).
 *   OS2S_HALF_F16 : every 16-bit tensor of the path -- layer inputs / outputs (x, a, the features), the weight
 *                   working copies and all activation gradients (dy, dA) -- is fp16 instead of bf16: the
 *                   reference's own "mixed" mode (fp16 storage + fp32 masters + loss scaling, mp_wrapper.py).
 *                   It is one switch, not per tensor: tcgen05.mma kind::f16 faults (illegal instruction) when
 *                   its two operands have different formats, and every tensor above meets every other one in
 *                   some MMA (x*w forward, dy*w data gradient, x*dy weight gradient).  fp16 gradient stores
 *                   do not saturate (OS2S_OUT_F16_GRAD): an overflow becomes inf and the Backoff scaler skips.
 *   OS2S_CONV_F32 : conv outputs y (the BN inputs, never a tensor-core operand) are fp32 instead of fp16.
 * With both flags the full 54-layer Jasper 10x5 reproduces the fp32 reference's logits to < 1e-2 (L2) at
 * random initialisation; DESIGN.md section 4 has the measured error of every combination. */
#define OS2S_HALF_F16 1
#define OS2S_CONV_F32 2

const char* os2s_last_error(void);
int os2s_version(void);
/* Kernel-variant selection for the conv entry points (all variants give bitwise-identical forward /
 * data-gradient outputs; the switch exists for A/B measurements and the parity tests).
 *   pair_mode: 0 = one CTA per 128-row tile, 1 = CTA pairs (cta_group::2) where the shape allows
 *              (default), 2 = pairs for every shape that can be paired, -1 = leave unchanged
 *   halo_mode: 0 = one activation tile per tap, 1 = one halo tile shared by all taps of a pair
 *              tile (default), -1 = leave unchanged
 * Initial values come from the environment (OS2S_CONV_PAIR, OS2S_CONV_HALO). */
int os2s_conv_tuning(int pair_mode, int halo_mode);
/* Grid size of the conv kernels in units of one CTA (or CTA pair) per SM: 1 = one persistent wave with a
 * static round-robin over the tiles (default), m > 1 = m x as many CTAs with 1/m of the tiles each, handed
 * out by the hardware block scheduler as CTAs retire (bounds the tail when the all-reduce's CTAs hold SMs).
 * Initial value: environment OS2S_CONV_WAVES. */
int os2s_conv_grid_waves(int waves);

/* ---- K2: tf.layers.conv1d(use_bias=False, padding=SAME), stride 1 -----------------------------
 * reference: open_seq2seq/parts/cnns/conv_blocks.py:195-206 (main), :79-85 (1x1 residual).
 * y[b,t,o] = sum_k sum_c x[b, t - pad_left + k*dil, c] * W[k,c,o]      (zero outside [0,T))
 *   x  : bf16 [B,T,C_in]          w  : bf16 [K][C_in][C_out]  (natural TF kernel layout)
 *   y  : bf16 / fp32 [B,T,C_out]  (out_mode)
 * The stride-2 first Jasper layer is expressed by the caller as a stride-1 conv over the input
 * viewed as [B, T/2, 2*C_in] with K' = ceil(K/2) taps (see JasperEngine._fix_folded_layout in openseq2seq_b200/engine.py).
 * bn_stats (may be NULL): fp32 [2][C_out], pre-zeroed; the epilogue adds the per-channel sum and sum of
 * squares of the ROUNDED outputs (what os2s_bn_stats would compute), fusing the BN statistics pass.
 * Constraints: C_in % 64 == 0, C_out % 64 == 0. */
int os2s_conv1d_fwd(const void* x, const void* w, void* y, int B, int T, int C_in, int C_out,
                    int K, int dil, int pad_left, int out_mode, float* bn_stats, void* stream);
/* dtypes & OS2S_HALF_F16: x and w are fp16.  bn_stats may be combined with OS2S_OUT_F32 (statistics of the fp32
 * outputs) as well as with the 2-byte modes.
 * row_lens (int32 [B] on the device, may be NULL) -- length-aware tile skipping, all four *_p conv entry points:
 * the caller guarantees that rows t >= row_lens[b] of the ACTIVATION the call reads (x here and in wgrad; for the
 * data gradients: of the layer input whose gradient is produced) are zero / masked, as the conv mask of
 * tdnn_encoder.py:185-186,204-205 and the zero padding of the batch make them.  Output tiles that then are exactly
 * zero (forward: t0 >= row_lens[b] + pad_left), are never read ungated (data gradient: t0 >= row_lens[b]) or add
 * nothing (weight gradient: 64-row chunks past row_lens[b] + pad_left) are not computed; a skipped forward tile
 * leaves y untouched (the BN kernels never read those rows unmasked, the fused statistics miss only zeros). */
int os2s_conv1d_fwd_p(const void* x, const void* w, void* y, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, int out_mode, float* bn_stats, const int* row_lens, int dtypes,
                      void* stream);

/* Same convolution with the weights given transposed, wt : bf16 [K][C_out][C_in] (K-major B
 * operand).  Kept for A/B measurements of the two operand layouts (tools/gpu_conv_check.py). */
int os2s_conv1d_fwd_wt(const void* x, const void* wt, void* y, int B, int T, int C_in, int C_out,
                       int K, int dil, int pad_left, int out_mode, void* stream);

/* dgrad of the above: dx[b,t,c] = sum_k sum_o dy[b, t + pad_left - k*dil, o] * W[k,c,o]
 *   dy : bf16 [B,T,C_out]   w : bf16 [K][C_in][C_out] (natural TF layout)   dx : out_mode */
int os2s_conv1d_dgrad(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, int out_mode, void* stream);
/* dtypes & OS2S_HALF_F16: dy and w are fp16; pass out_mode = OS2S_OUT_F16_GRAD for an fp16 dx. */
int os2s_conv1d_dgrad_p(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                        int K, int dil, int pad_left, int out_mode, const int* row_lens, int dtypes, void* stream);

/* Data gradient whose output dx is the gradient dA of a single-branch BN + ReLU + dropout layer
 * (conv_blocks.py:208-227 followed by tdnn_encoder.py:255): dx is written as bf16 and the epilogue
 * also accumulates that layer's batch-norm backward reductions,
 *   red[0][c] += sum_rows dz,  red[1][c] += sum_rows dz * y,   dz = dx * [a != 0] / keep,
 * (a = the layer's forward output, y = its conv output, both [B,T,C_in]; red fp32 [2][C_in], zeroed by
 * the caller), so that os2s_bn_bwd_apply can skip the separate reduction pass over dA, a and y. */
int os2s_conv1d_dgrad_bnred(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                            int K, int dil, int pad_left, const void* a, const void* y, float keep, float* red,
                            void* stream);
/* dtypes & OS2S_CONV_F32: y is fp32; dtypes & OS2S_HALF_F16: dy, w and dx are fp16 (a is only tested for
 * zero bits, so either format is accepted). */
int os2s_conv1d_dgrad_bnred_p(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                              int K, int dil, int pad_left, const void* a, const void* y, float keep, float* red,
                              const int* row_lens, int dtypes, void* stream);
/* Second half of os2s_bn_bwd for one branch when `red` already holds the two sums (see above). */
int os2s_bn_bwd_apply(const void* y, const float* mean_invstd, const float* gamma, float* dgamma, float* dbeta,
                      void* dy, const void* dA, const void* a, const float* red, int M, int C, float keep,
                      void* stream);
int os2s_bn_bwd_apply_p(const void* y, const float* mean_invstd, const float* gamma, float* dgamma, float* dbeta,
                        void* dy, const void* dA, const void* a, const float* red, int M, int C, float keep,
                        int dtypes, void* stream);

/* wgrad: dw[k,c,o] = sum_{b,t} x[b, t - pad_left + k*dil, c] * dy[b,t,o]   (fp32 [K][C_in][C_out])
 * Overwrites dw.  Constraints: C_in % 128 == 0, C_out % 64 == 0.  (_p: dtypes & OS2S_HALF_F16: x and dy are fp16.) */
int os2s_conv1d_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, void* stream);
int os2s_conv1d_wgrad_p(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out,
                        int K, int dil, int pad_left, const int* row_lens, int dtypes, void* stream);

/* W fp32 [K][C_in][C_out] -> w bf16 (same layout) and wt bf16 [K][C_out][C_in]. Either output may
 * be NULL.  Replaces the fp32->fp16 assign of mp_wrapper.py:104-109 when used standalone. */
int os2s_weight_cast_transpose(const float* w_master, void* w_bf16, void* wt_bf16, int K, int C_in,
                               int C_out, void* stream);
int os2s_weight_cast_transpose_p(const float* w_master, void* w_half, void* wt_half, int K, int C_in,
                                 int C_out, int dtypes, void* stream);

/* ---- K3/K4: batch norm (training) + residual sum + ReLU + dropout + sequence mask ------------
 * reference: tf.layers.batch_normalization(training=True) conv_blocks.py:208-227 / :91-101,
 * residual sum :154, activation :166, tf.nn.dropout tdnn_encoder.py:255, mask :185-186,204-205. */

/* stats[0..C) += sum_rows y ; stats[C..2C) += sum_rows y^2   (y FP16 [M,C]; stats pre-zeroed).
 * All BN entry points take the conv output y as fp16 (OS2S_OUT_F16): it is never a tensor-core
 * operand, and fp16 keeps 3 more mantissa bits than bf16 at the same HBM traffic. */
int os2s_bn_stats(const void* y, float* stats, int M, int C, void* stream);

/* out = rowmask(dropout(act(sum_j gamma_j*(y_j-mean_j)*invstd_j + beta_j))), bf16 [B,T,C].
 * The *_host arrays hold n_branch device pointers (branch 0 = main conv, 1.. = residual convs).
 * mean_invstd[j] ([2][C]) is written for backward; moving[j] ([2][C] moving_mean, moving_variance,
 * may be NULL) is updated in place.  lens (int32 [B], may be NULL) gives valid rows per utterance;
 * rows t >= lens[b] are written as zeros.  keep = dropout keep probability (1 disables dropout);
 * apply_relu: 0 = identity, 1 = relu, with relu_clip > 0 -> min(relu(x), relu_clip).
 * use_moving = 1 is inference mode (training=False): normalise with moving[j], touch no statistics.
 * step_counter_dev (int64 on the device, may be NULL) is mixed into the dropout seed so that a
 * CUDA-graph replay of the same launch draws a fresh mask every step. */
int os2s_bn_apply_fwd(int n_branch, const void* const* y_host, const float* const* stats_host,
                      const float* const* gamma_host, const float* const* beta_host,
                      float* const* mean_invstd_host, float* const* moving_host, void* out,
                      const int* lens, int B, int T, int C, float eps, float momentum, float keep,
                      uint64_t seed, int apply_relu, float relu_clip, int use_moving,
                      const long long* step_counter_dev, void* stream);

/* Backward of the above.  dA: gradient wrt `out` (bf16, or fp32 when dA_is_f32), a: the forward
 * output (its zeros encode relu / dropout / mask).  Writes dy[j] (bf16 [M,C]) for every branch and
 * dgamma[j], dbeta[j] (fp32 [C]).  red: fp32 scratch [(1+n_branch)*C], zeroed by this call. */
int os2s_bn_bwd(int n_branch, const void* const* y_host, const float* const* mean_invstd_host,
                const float* const* gamma_host, float* const* dgamma_host, float* const* dbeta_host,
                void* const* dy_host, const void* dA, int dA_is_f32, const void* a, float* red, int M,
                int C, float keep, int apply_relu, void* stream);

/* Same two calls for branches whose conv outputs / gradients are COLUMN SLICES of wider matrices
 * (the dense-residual 1x1 convolutions of one source are computed as ONE GEMM whose output holds the
 * slices of all consumer blocks side by side, see os2s_multi_copy_2d).  ld_host[j]: row stride of
 * y[j] (and dy[j]) in elements; stats_ld_host[j]: distance between the sum row and the
 * sum-of-squares row of stats[j].  Either array may be NULL (= C). */
int os2s_bn_apply_fwd_ld(int n_branch, const void* const* y_host, const int* ld_host,
                         const float* const* stats_host, const int* stats_ld_host,
                         const float* const* gamma_host, const float* const* beta_host,
                         float* const* mean_invstd_host, float* const* moving_host, void* out,
                         const int* lens, int B, int T, int C, float eps, float momentum, float keep,
                         uint64_t seed, int apply_relu, float relu_clip, int use_moving,
                         const long long* step_counter_dev, void* stream);
int os2s_bn_bwd_ld(int n_branch, const void* const* y_host, const int* ld_host,
                   const float* const* mean_invstd_host, const float* const* gamma_host,
                   float* const* dgamma_host, float* const* dbeta_host, void* const* dy_host, const void* dA,
                   int dA_is_f32, const void* a, float* red, int M, int C, float keep, int apply_relu,
                   void* stream);
/* The _ld calls with selectable storage formats: OS2S_CONV_F32 -> every y[j] is fp32 (ld still in
 * elements); OS2S_HALF_F16 -> `out` is written as fp16 (forward); a 16-bit dA is read and every dy[j] is
 * written as fp16 (backward; `a` is only tested for zero bits). */
int os2s_bn_apply_fwd_p(int n_branch, const void* const* y_host, const int* ld_host,
                        const float* const* stats_host, const int* stats_ld_host,
                        const float* const* gamma_host, const float* const* beta_host,
                        float* const* mean_invstd_host, float* const* moving_host, void* out,
                        const int* lens, int B, int T, int C, float eps, float momentum, float keep,
                        uint64_t seed, int apply_relu, float relu_clip, int use_moving,
                        const long long* step_counter_dev, int dtypes, void* stream);
int os2s_bn_bwd_p(int n_branch, const void* const* y_host, const int* ld_host,
                  const float* const* mean_invstd_host, const float* const* gamma_host,
                  float* const* dgamma_host, float* const* dbeta_host, void* const* dy_host, const void* dA,
                  int dA_is_f32, const void* a, float* red, int M, int C, float keep, int apply_relu,
                  int dtypes, void* stream);

/* n (<= 64) strided 2-D copies in one launch: dst[i][r][0:row_bytes) = src[i][r][0:row_bytes) for
 * r < rows[i], row r at base + r * pitch.  row_bytes, pitches and base addresses are multiples of 16.
 * Used to lay the 1x1 residual kernels W_{b,j} [C_j][C_b] of all consumer blocks b of a source j
 * side by side ([C_j][sum_b C_b], one GEMM operand; conv_blocks.py:79-85 issues one conv1d per
 * (b, j)) and to scatter the weight gradient of that GEMM back to the per-variable buffers. */
int os2s_multi_copy_2d(int n, const void* const* src_host, void* const* dst_host, const int* rows_host,
                       const int* row_bytes_host, const long long* src_pitch_host,
                       const long long* dst_pitch_host, void* stream);

/* ---- K5: tf.layers.dense of FullyConnectedTimeDecoder (fc_decoders.py:135-140) --------------
 * logits fp32 [M,V] = x bf16 [M,H] * w fp32 [H,V] + bias;  V <= 32. */
int os2s_fc_fwd(const void* x, const float* w, const float* bias, float* logits, int M, int H, int V,
                void* stream);
/* dx bf16 [M,H] (may be NULL), dw fp32 [H,V], db fp32 [V] (dw/db overwritten; may be NULL). */
int os2s_fc_bwd(const void* x, const float* dlogits, const float* w, void* dx, float* dw, float* db,
                int M, int H, int V, void* stream);
/* dtypes & OS2S_HALF_F16: x is fp16 and dx is written as fp16. */
int os2s_fc_fwd_p(const void* x, const float* w, const float* bias, float* logits, int M, int H, int V,
                  int dtypes, void* stream);
int os2s_fc_bwd_p(const void* x, const float* dlogits, const float* w, void* dx, float* dw, float* db,
                  int M, int H, int V, int dtypes, void* stream);

/* ---- K6: tf.nn.ctc_loss(ignore_longer_outputs_than_inputs=True) + mask_nans (ctc_loss.py:77-89)
 * logits fp32, element (b,t,v) at logits[b*stride_b + t*stride_t + v]; blank = V-1.
 * labels int32 [B,L_max], label_lens/input_lens int32 [B].
 * loss[b] = -log p(l|x) (0 for skipped/NaN utterances); grad [B,T,V] fp32 contiguous =
 * d(mean_b loss)/dlogits * (*loss_scale_dev) (loss_scale_dev may be NULL = 1).
 * workspace: os2s_ctc_workspace_bytes(B,T,L_max) bytes of device memory. */
size_t os2s_ctc_workspace_bytes(int B, int T, int L_max);
int os2s_ctc_loss_fwd_bwd(const float* logits, const int* labels, const int* label_lens,
                          const int* input_lens, float* grad, float* loss, void* workspace,
                          size_t workspace_bytes, const float* loss_scale_dev, int B, int T, int V,
                          int L_max, long long stride_b, long long stride_t, void* stream);

/* ---- K7: tf.nn.ctc_greedy_decoder(merge_repeated) (fc_decoders.py:247-250) --------------------
 * tokens int32 [B,T] (first out_lens[b] entries valid), neg_sum_logits fp32 [B] (may be NULL). */
int os2s_ctc_greedy(const float* logits, const int* input_lens, int* tokens, int* out_lens,
                    float* neg_sum_logits, int B, int T, int V, long long stride_b,
                    long long stride_t, int merge_repeated, void* stream);

/* ---- K8: optimizer chain (mp_wrapper.py, optimizers.py LARC, automatic_loss_scaler.py,
 *          novograd.py, lr_policies.py poly_decay) as three launches ---------------------------- */
typedef struct os2s_opt_hparams {
  int algo;              /* 0 = NovoGrad, 1 = Momentum SGD, 2 = Adam (tf.train.AdamOptimizer; os2s_opt_step2) */
  float beta1, beta2, epsilon, weight_decay, momentum;
  int grad_averaging;
  int ema_persist;       /* 0 = reference as written (v_t = |g_t|^2), 1 = corrected NovoGrad EMA */
  float larc_eta;        /* <= 0 disables LARC */
  float larc_eps, larc_min_update;
  int larc_mode;         /* 0 = clip, 1 = scale */
  float lr0, min_lr, power;
  long long decay_steps, begin_decay_at, warmup_steps;
  int use_loss_scaler;   /* Backoff scaler on/off */
  float scale_min, scale_max, step_factor;
  long long step_window;
  int world_size;        /* gradients are SUMS over this many ranks */
  /* learning-rate policy (lr_policies.py): 0 = poly_decay (:95-131), 1 = cosine_decay (:134-170; the
   * reference passes min_lr as tf.train.cosine_decay's alpha, i.e. a FRACTION of lr0 -- kept),
   * 2 = exp_decay (:55-92, no warm-up), 3 = fixed_lr (:15-27) */
  int lr_policy;
  float decay_rate;      /* exp_decay */
  int staircase;         /* exp_decay: use_staircase_decay */
  float max_grad_norm;   /* > 0: tf.clip_by_global_norm on the unscaled, rank-averaged gradients
                          * (optimizers.py:408-433); exclusive with LARC as in the reference (:161-164) */
  int wb_f16;            /* working copies are written as fp16 (OS2S_HALF_F16) instead of bf16 */
} os2s_opt_hparams;

/* Tensor table (all device memory, owned by the caller):
 *   w,g,m,wb : arrays [n_tensors] of device pointers (fp32 master, fp32 gradient of loss*scale,
 *              fp32 momentum, bf16 working copy or 0)
 *   sizes    : int64 [n_tensors];  chunk_tensor int32 / chunk_offset int64 [n_chunks] with chunks of
 *              os2s_opt_chunk_elems() elements
 *   norms fp32 [2*n_tensors] and nonfinite int32 [1] zeroed once at creation; coef, ema fp32 [n_tensors]
 *   fstate fp32 [8]: [0] loss scale, [1] lr of the last step, [2] global grad norm, [3] Adam's lr_t,
 *                    [4] loss scale the current gradients carry
 *   istate int64 [8]: [0] scaler iteration, [1] last overflow iteration (init -1), [2] global_step,
 *                     [3] last step skipped?, [4] skipped-step count, [5] attempted-step count */
int os2s_opt_chunk_elems(void);
int os2s_opt_step(void* const* w, void* const* g, void* const* m, void* const* wb,
                  const long long* sizes, const int* chunk_tensor, const long long* chunk_offset,
                  int n_tensors, int n_chunks, const os2s_opt_hparams* hp, float* norms,
                  int* nonfinite, float* fstate, long long* istate, float* coef, float* ema,
                  void* stream);

/* Same step with (a) a second per-parameter state array v (fp32, like m): required for algo = 2 (Adam:
 * m <- b1*m + (1-b1)*g, v <- b2*v + (1-b2)*g^2, w <- w - lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps),
 * t = number of applied steps), ignored (may be NULL) otherwise; (b) reg: device fp32 [n_tensors] of
 * per-variable L2-regulariser scales (tf.contrib.layers.l2_regularizer(scale): gradient scale*w added
 * to the unscaled fp32 gradient before LARC, mp_wrapper.py:81-89; 0 for variables built without a
 * regularizer), may be NULL. */
int os2s_opt_step2(void* const* w, void* const* g, void* const* m, void* const* v, void* const* wb,
                   const float* reg, const long long* sizes, const int* chunk_tensor,
                   const long long* chunk_offset, int n_tensors, int n_chunks, const os2s_opt_hparams* hp,
                   float* norms,
                   int* nonfinite, float* fstate, long long* istate, float* coef, float* ema,
                   void* stream);

/* Same step with `frozen`: device int32 [n_tensors], non-zero = the variable is not in var_list
 * (freeze_variables_regex, models/model.py:502-507): no update, no weight decay, no momentum, not part of the
 * global gradient norm.  May be NULL. */
int os2s_opt_step3(void* const* w, void* const* g, void* const* m, void* const* v, void* const* wb,
                   const float* reg, const int* frozen, const long long* sizes, const int* chunk_tensor,
                   const long long* chunk_offset, int n_tensors, int n_chunks, const os2s_opt_hparams* hp,
                   float* norms,
                   int* nonfinite, float* fstate, long long* istate, float* coef, float* ema,
                   void* stream);

/* wt[k][c][r] = w[k][r][c] for n tensors in one launch (bf16). src/dst: device arrays of device
 * pointers; K,R,C: host int arrays. tile_start: device int64 [n+1] prefix of K*ceil(R/32)*ceil(C/32);
 * Rdev/Cdev device copies of R/C. */
int os2s_multi_transpose(void* const* src, void* const* dst, const int* Rdev, const int* Cdev,
                         const long long* tile_start, int n_tensors, long long total_tiles,
                         void* stream);

/* ---- K1: Speech2TextDataLayer featurizer (speech_utils.py:322-441, librosa backend, logfbank) -
 * wave int16 (all utterances concatenated), offsets int64 [B], n_samples int32 [B];
 * mel fp32 [F][n_fft/2+1], mel_band int32 [F][2] = [lo,hi) support of each filter (may be NULL =
 * dense), window fp32 [win]; out [B,T_pad,F] bf16 and/or fp32 (either may be
 * NULL), out_lens int32 [B] = frames per utterance. absmax_ws: uint32 [B], raw_ws: fp32 [B*T_pad*F]. */
int os2s_logmel_forward(const int16_t* wave, const long long* offsets, const int* n_samples, int B,
                        const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop, int F,
                        int T_pad, int max_samples, float dither, uint64_t seed, float preemph,
                        void* absmax_ws, float* raw_ws, void* out_bf16, float* out_f32, int* out_lens,
                        void* stream);

/* Same kernels with the backend conventions selectable (speech_utils.py:444-535, the default
 * "psf" backend of the Wave2Letter(+) configs, input_type = "logfbank"):
 *   psf_backend = 1: signal re-quantised to int16 and zero-padded so that the frame count is a multiple
 *     of pad_to (:473-488), psf.logfbank = rectangular frames [f*hop, f*hop+win), pre-emphasis, |rfft|^2/n_fft,
 *     the caller's filterbank `mel` (python_speech_features.get_filterbanks), zeros -> eps, log;
 *     pass window = ones(win);
 *   norm_per_feature = 0: ONE mean / population std per utterance over frames x features (:531-533, and
 *     norm_per_feature=False of the librosa backend :411-417).
 * out_lens[b] = frames of utterance b under the selected convention. */
int os2s_features_forward(const int16_t* wave, const long long* offsets, const int* n_samples, int B,
                          const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop,
                          int F, int T_pad, int max_samples, float dither, uint64_t seed, float preemph,
                          int psf_backend, int pad_to, int norm_per_feature, void* absmax_ws, float* raw_ws,
                          void* out_bf16, float* out_f32, int* out_lens, void* stream);

/* ---- K2b: tf.layers.separable_conv1d(use_bias=False, padding=SAME) of the sep_conv1d layers -----------
 * reference: open_seq2seq/parts/cnns/conv_blocks.py:27-40,180-193 (main conv), :79-85 (the residual branch of a
 * sep_conv1d block is a separable conv with kernel_size 1).  depthwise_kernel D fp32 [K][C_in] (TF shape
 * [K, C_in, 1]), pointwise_kernel P fp32 [C_in][C_out] (TF shape [1, C_in, C_out]).
 *
 * Depthwise stage of a wide stride-1 layer (the pointwise stage is os2s_conv1d_* with K = 1):
 *   out[b,t,c] = sum_k taps[k,c] * x[b, t + t_off0 + k*t_step, c]      (zero outside [0,T))
 *   forward: t_off0 = -pad_left, t_step = dilation;  data gradient: t_off0 = +pad_left, t_step = -dilation.
 *   x: 16-bit [B,T,C]; out_mode: OS2S_OUT_BF16 / OS2S_OUT_F16_GRAD (the 16-bit format that `dtypes` selects),
 *   OS2S_OUT_F32 or OS2S_OUT_F32_ACC (gradient of a residual source).  C % 64 == 0. */
int os2s_depthwise_conv1d(const void* x, const float* taps, void* out, int B, int T, int C, int K, int t_off0,
                          int t_step, int out_mode, int dtypes, void* stream);
/* d_taps[k,c] = sum_{b,t} x[b, t - pad_left + k*dil, c] * dz[b,t,c]   (fp32 [K][C], overwritten; K <= 96) */
int os2s_depthwise_conv1d_wgrad(const void* x, const void* dz, float* d_taps, int B, int T, int C, int K, int dil,
                                int pad_left, int dtypes, void* stream);
/* Composed form (K = 1 residual branches, the stride-2 first layer): w_half[k,c,o] = D[k,c] * P[c,o], the
 * 16-bit working copy of the equivalent dense kernel, and the fold-back of its dense weight gradient:
 *   d_depthwise[k,c] = sum_o dw_dense[k,c,o] * P[c,o],   d_pointwise[c,o] = sum_k dw_dense[k,c,o] * D[k,c]. */
int os2s_sepconv_compose(const float* depthwise, const float* pointwise, void* w_half, int K, int C_in, int C_out,
                         int dtypes, void* stream);
int os2s_sepconv_decompose_grad(const float* dw_dense, const float* depthwise, const float* pointwise,
                                float* d_depthwise, float* d_pointwise, int K, int C_in, int C_out, void* stream);

/* ---- K1b: audio augmentation (speech_utils.py:225-268) and the remaining data-layer options ------------
 * absmax[b] = max |wave_b| (uint32 [B]): the gain of normalize_signal (:216-222) is 1 / (absmax + 1e-5). */
int os2s_wave_absmax(const int16_t* wave, const long long* offsets, const int* n_samples, int B, void* absmax,
                     void* stream);
/* out_b = resample(wave_b * gain_b, sr_orig -> sr_new[b]) + noise_amp[b] * N(0,1)   (fp32, normalised signal)
 *   speed perturbation (:245-259): resampy.resample(filter='kaiser_best'), i.e. band-limited sinc interpolation
 *   with the right half of the Kaiser-windowed sinc in interp_win (n_win = num_zeros * num_table + 1 entries,
 *   num_table samples per zero crossing, linear interpolation between entries; the caller builds the table);
 *   sr_new[b] = int(sr_orig * stretch) or 0 for "not resampled"; n_out[b] = int(n_samples[b] * sr_new / sr_orig).
 *   noise (:262-266): noise_amp[b] = 10^(dB / 20) or 0.  sr_new / noise_amp / interp_win may be NULL.
 *   gain > 0 replaces 1 / (absmax + 1e-5) (params['gain']); absmax may then be NULL. */
int os2s_augment_signal(const int16_t* wave, const long long* offsets, const int* n_samples, int B,
                        const void* absmax, float gain, const int* sr_new, int sr_orig, const float* interp_win,
                        int n_win, int num_table, const float* noise_amp, uint64_t seed, float* out,
                        const long long* out_offsets, const int* n_out, int max_out, void* stream);
/* os2s_features_forward with the remaining options of get_speech_features_librosa (:322-441):
 *   sig / sig_offsets : fp32 signal from os2s_augment_signal (replaces wave * gain; offsets / n_samples then
 *                       describe sig); librosa backend only
 *   gain              : > 0 = params['gain'] (fixed normalisation gain), 0 = 1 / (max|x| + 1e-5)
 *   features_mean / features_std : fp32 [F] (params['features_mean'], ['features_std_dev']) or NULL = computed
 *   masks             : int32 [B][n_masks][3] = (kind 0 = frequency / 1 = time, base, width): spec-augment
 *                       (:419-433) zeros written into the normalised features; width 0 = no-op
 *   feature_type      : 0 = logfbank; the psf backend's other input types (:490-512): 1 = spectrogram (Hann
 *                       frames, |DFT|^2 / win with NFFT = win, 10 log10, lowest F bins; pass window = hanning(win)),
 *                       2 = mfcc (psf.mfcc: `mel` has n_filt = 2 F rows, mfcc_matrix fp32 [F][n_filt] = sinusoidal
 *                       lifter x orthonormal DCT-II rows, applied to the log filterbank energies)
 *   offsets           : always a valid int64 [B] array (the wave offsets; unused values when sig is given)
 *   dtypes & OS2S_HALF_F16 : out16 is fp16 instead of bf16 */
int os2s_features_forward_p(const int16_t* wave, const float* sig, const long long* sig_offsets,
                            const long long* offsets, const int* n_samples, int B,
                            const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop,
                            int F, int T_pad, int max_samples, float dither, uint64_t seed, float preemph,
                            int psf_backend, int pad_to, int norm_per_feature, float gain,
                            const float* features_mean, const float* features_std, const int* masks, int n_masks,
                            int feature_type, const float* mfcc_matrix, int n_filt,
                            void* absmax_ws, float* raw_ws, void* out16, float* out_f32, int* out_lens,
                            int dtypes, void* stream);

/* ---- C1: gradient sum across the ranks of one node over NVLink peer memory ----------------------------
 * Replaces hvd.allreduce(grad) per variable (optimizers/optimizers.py:77-104, reduce_gradients).  Every rank
 * exports its flat fp32 gradient buffer and one zero-filled staging buffer (os2s_peer_stage_bytes) with
 * os2s_ipc_export, the 64-byte handles + offsets travel through the caller's rendezvous, and every rank maps
 * the others' buffers with os2s_ipc_open (base pointer of the exporter's allocation; add the offset).
 * os2s_peer_exchange_bucket(ctx, b, stream) sums bucket b = [bucket_start[b], bucket_end[b]) (floats, starts
 * 16-byte aligned, the same list on every rank) over all ranks IN PLACE in every rank's gradient buffer:
 *   contribution to slice p -> staging slot on rank p (copy engine), flag; rank r adds the N-1 staged slices
 *   onto its slice r (one narrow kernel); summed slice -> every other rank's gradient buffer (copy engine), flag.
 * os2s_peer_finish enqueues the wait for the last phase of ALL buckets: after it, `g` holds the sum on this
 * rank.  Each bucket must be exchanged exactly once between two os2s_peer_finish calls, in the same order on
 * every rank.  All work is stream-ordered (graph-capturable); waits that exceed timeout_s raise a device flag
 * (os2s_peer_timed_out, synchronising) instead of spinning forever.  Every rank ends with identical bits. */
long long os2s_peer_stage_bytes(int world, int n_buckets, const long long* bucket_start_host,
                                const long long* bucket_end_host);
int os2s_ipc_export(const void* ptr, unsigned char* handle64_host, long long* offset_host);
int os2s_ipc_open(const unsigned char* handle64_host, void** base_host);
int os2s_ipc_close(void* base);
/* grad_host / stage_host: `world` device pointers each (entry [rank] = the local buffers) */
int os2s_peer_create(int rank, int world, void* const* grad_host, void* const* stage_host, int n_buckets,
                     const long long* bucket_start_host, const long long* bucket_end_host, double timeout_s,
                     void** ctx_host);
int os2s_peer_destroy(void* ctx);
int os2s_peer_set_timeout(void* ctx, double timeout_s); /* applies to waits enqueued afterwards */
int os2s_peer_exchange_bucket(void* ctx, int bucket, void* stream);
int os2s_peer_finish(void* ctx, void* stream);
int os2s_peer_timed_out(void* ctx, int* flag_host);

#ifdef __cplusplus
}
#endif
#endif /* OS2S_H_ */
