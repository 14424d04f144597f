// extern "C" boundary of libos2s_b200 (declared in include/os2s.h).
#include "../../include/os2s.h"

#include "common.h"
#include "kernels.h"

using namespace os2s;

extern "C" {

const char* os2s_last_error(void) { return last_error_cstr(); }
int os2s_version(void) { return 101; }
int os2s_conv_tuning(int pair_mode, int halo_mode) { return conv_tuning(pair_mode, halo_mode); }
int os2s_conv_grid_waves(int waves) { return conv_grid_waves_set(waves); }

int os2s_conv1d_fwd_p(const void* x, const void* w, void* y, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, int out_mode, float* bn_stats, const int* row_lens, int dtypes,
                      void* stream) {
  if (!x || !w || !y) return fail(ERR_INVALID, "os2s_conv1d_fwd: null pointer");
  if (bn_stats && out_mode == OS2S_OUT_F32_ACC)
    return fail(ERR_INVALID, "os2s_conv1d_fwd: fused BN statistics need an overwriting output mode");
  if (dtypes & ~(OS2S_HALF_F16 | OS2S_CONV_F32)) return fail(ERR_INVALID, "os2s_conv1d_fwd: unknown dtype flag");
  // B operand MN-major straight from the natural [K][C_in][C_out] layout
  return conv_kmajor(x, w, y, B, T, C_in, C_out, K, -pad_left, dil, out_mode, 1, bn_stats, (cudaStream_t)stream,
                     nullptr, nullptr, 1.f, (dtypes & OS2S_HALF_F16) ? 1 : 0, 0, row_lens, pad_left);
}
int os2s_conv1d_fwd(const void* x, const void* w, void* y, int B, int T, int C_in, int C_out,
                    int K, int dil, int pad_left, int out_mode, float* bn_stats, void* stream) {
  return os2s_conv1d_fwd_p(x, w, y, B, T, C_in, C_out, K, dil, pad_left, out_mode, bn_stats, nullptr, 0, stream);
}

int os2s_conv1d_fwd_wt(const void* x, const void* wt, void* y, int B, int T, int C_in, int C_out,
                       int K, int dil, int pad_left, int out_mode, void* stream) {
  if (!x || !wt || !y) return fail(ERR_INVALID, "os2s_conv1d_fwd_wt: null pointer");
  return conv_kmajor(x, wt, y, B, T, C_in, C_out, K, -pad_left, dil, out_mode, 0, nullptr, (cudaStream_t)stream);
}

int os2s_conv1d_dgrad_p(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                        int K, int dil, int pad_left, int out_mode, const int* row_lens, int dtypes, void* stream) {
  if (!dy || !w || !dx) return fail(ERR_INVALID, "os2s_conv1d_dgrad: null pointer");
  return conv_kmajor(dy, w, dx, B, T, C_out, C_in, K, pad_left, -dil, out_mode, 0, nullptr, (cudaStream_t)stream,
                     nullptr, nullptr, 1.f, (dtypes & OS2S_HALF_F16) ? 1 : 0, 0, row_lens, 0);
}
int os2s_conv1d_dgrad(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, int out_mode, void* stream) {
  return os2s_conv1d_dgrad_p(dy, w, dx, B, T, C_in, C_out, K, dil, pad_left, out_mode, nullptr, 0, stream);
}

int os2s_conv1d_dgrad_bnred_p(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                              int K, int dil, int pad_left, const void* a, const void* y, float keep, float* red,
                              const int* row_lens, int dtypes, void* stream) {
  if (!dy || !w || !dx || !a || !y || !red) return fail(ERR_INVALID, "os2s_conv1d_dgrad_bnred: null pointer");
  if (!(keep > 0.f && keep <= 1.f)) return fail(ERR_INVALID, "os2s_conv1d_dgrad_bnred: keep must be in (0,1]");
  // dx is written in the 16-bit format of the mode (fp16 gradients are NOT saturated: overflow -> inf -> the
  // loss scaler backs off); a is only tested for zero bits
  const int f16 = (dtypes & OS2S_HALF_F16) ? 1 : 0;
  return conv_kmajor(dy, w, dx, B, T, C_out, C_in, K, pad_left, -dil, f16 ? OS2S_OUT_F16_GRAD : OS2S_OUT_BF16, 0, red,
                     (cudaStream_t)stream, a, y, 1.f / keep, f16, (dtypes & OS2S_CONV_F32) ? 1 : 0, row_lens, 0);
}
int os2s_conv1d_dgrad_bnred(const void* dy, const void* w, void* dx, int B, int T, int C_in, int C_out,
                            int K, int dil, int pad_left, const void* a, const void* y, float keep, float* red,
                            void* stream) {
  return os2s_conv1d_dgrad_bnred_p(dy, w, dx, B, T, C_in, C_out, K, dil, pad_left, a, y, keep, red, nullptr, 0, stream);
}

int os2s_conv1d_wgrad_p(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out,
                        int K, int dil, int pad_left, const int* row_lens, int dtypes, void* stream) {
  if (!x || !dy || !dw) return fail(ERR_INVALID, "os2s_conv1d_wgrad: null pointer");
  return conv_wgrad(x, dy, dw, B, T, C_in, C_out, K, dil, pad_left, nullptr, (cudaStream_t)stream,
                    (dtypes & OS2S_HALF_F16) ? 1 : 0, row_lens);
}
int os2s_conv1d_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out,
                      int K, int dil, int pad_left, void* stream) {
  return os2s_conv1d_wgrad_p(x, dy, dw, B, T, C_in, C_out, K, dil, pad_left, nullptr, 0, stream);
}

int os2s_weight_cast_transpose_p(const float* w_master, void* w_half, void* wt_half, int K, int C_in,
                                 int C_out, int dtypes, void* stream) {
  if (!w_master) return fail(ERR_INVALID, "os2s_weight_cast_transpose: null pointer");
  return weight_cast_transpose(w_master, w_half, wt_half, K, C_in, C_out, (cudaStream_t)stream,
                               (dtypes & OS2S_HALF_F16) ? 1 : 0);
}
int os2s_weight_cast_transpose(const float* w_master, void* w_bf16, void* wt_bf16, int K, int C_in,
                               int C_out, void* stream) {
  return os2s_weight_cast_transpose_p(w_master, w_bf16, wt_bf16, K, C_in, C_out, 0, stream);
}


int os2s_bn_stats(const void* y, float* stats, int M, int C, void* stream) {
  if (!y || !stats) return fail(ERR_INVALID, "os2s_bn_stats: null pointer");
  return bn_stats(y, stats, M, C, (cudaStream_t)stream);
}

int os2s_bn_apply_fwd_p(int n_branch, const void* const* y_host, const int* ld_host,
                        const float* const* stats_host, const int* stats_ld_host,
                        const float* const* gamma_host, const float* const* beta_host,
                        float* const* mean_invstd_host, float* const* moving_host, void* out,
                        const int* lens, int B, int T, int C, float eps, float momentum, float keep,
                        uint64_t seed, int apply_relu, float relu_clip, int use_moving,
                        const long long* step_counter_dev, int dtypes, void* stream) {
  if (n_branch < 1 || n_branch > kMaxBranches) return fail(ERR_INVALID, "os2s_bn_apply_fwd: 1..12 branches");
  if (!y_host || !stats_host || !gamma_host || !beta_host || !mean_invstd_host || !out)
    return fail(ERR_INVALID, "os2s_bn_apply_fwd: null pointer");
  if (use_moving && !moving_host) return fail(ERR_INVALID, "os2s_bn_apply_fwd: use_moving needs moving statistics");
  if (!(keep > 0.f && keep <= 1.f)) return fail(ERR_INVALID, "os2s_bn_apply_fwd: keep must be in (0,1]");
  BnFwdParams p;
  for (int j = 0; j < n_branch; ++j) {
    p.br[j].y = y_host[j];
    p.br[j].stats = stats_host[j];
    p.br[j].gamma = gamma_host[j];
    p.br[j].beta = beta_host[j];
    p.br[j].mean_invstd = mean_invstd_host[j];
    p.br[j].moving = moving_host ? moving_host[j] : nullptr;
    p.br[j].ld = ld_host ? ld_host[j] : C;
    p.br[j].stats_ld = stats_ld_host ? stats_ld_host[j] : C;
  }
  p.n_branch = n_branch;
  p.out = out;
  p.y_f32 = (dtypes & OS2S_CONV_F32) ? 1 : 0;
  p.out_f16 = (dtypes & OS2S_HALF_F16) ? 1 : 0;
  p.lens = lens;
  p.B = B; p.T = T; p.C = C;
  p.eps = eps; p.momentum = momentum; p.keep = keep; p.seed = seed;
  p.relu_clip = relu_clip; p.apply_relu = apply_relu; p.use_moving = use_moving;
  p.step_ctr = step_counter_dev;
  return bn_apply_fwd(p, (cudaStream_t)stream);
}
int os2s_bn_apply_fwd_ld(int n_branch, const void* const* y_host, const int* ld_host,
                         const float* const* stats_host, const int* stats_ld_host,
                         const float* const* gamma_host, const float* const* beta_host,
                         float* const* mean_invstd_host, float* const* moving_host, void* out,
                         const int* lens, int B, int T, int C, float eps, float momentum, float keep,
                         uint64_t seed, int apply_relu, float relu_clip, int use_moving,
                         const long long* step_counter_dev, void* stream) {
  return os2s_bn_apply_fwd_p(n_branch, y_host, ld_host, stats_host, stats_ld_host, gamma_host, beta_host,
                             mean_invstd_host, moving_host, out, lens, B, T, C, eps, momentum, keep, seed, apply_relu,
                             relu_clip, use_moving, step_counter_dev, 0, stream);
}

int os2s_bn_apply_fwd(int n_branch, const void* const* y_host, const float* const* stats_host,
                      const float* const* gamma_host, const float* const* beta_host,
                      float* const* mean_invstd_host, float* const* moving_host, void* out,
                      const int* lens, int B, int T, int C, float eps, float momentum, float keep,
                      uint64_t seed, int apply_relu, float relu_clip, int use_moving,
                      const long long* step_counter_dev, void* stream) {
  return os2s_bn_apply_fwd_ld(n_branch, y_host, nullptr, stats_host, nullptr, gamma_host, beta_host,
                              mean_invstd_host, moving_host, out, lens, B, T, C, eps, momentum, keep, seed,
                              apply_relu, relu_clip, use_moving, step_counter_dev, stream);
}

int os2s_bn_bwd_p(int n_branch, const void* const* y_host, const int* ld_host,
                  const float* const* mean_invstd_host, const float* const* gamma_host,
                  float* const* dgamma_host, float* const* dbeta_host, void* const* dy_host, const void* dA,
                  int dA_is_f32, const void* a, float* red, int M, int C, float keep, int apply_relu,
                  int dtypes, void* stream) {
  if (n_branch < 1 || n_branch > kMaxBranches) return fail(ERR_INVALID, "os2s_bn_bwd: 1..12 branches");
  if (!y_host || !mean_invstd_host || !gamma_host || !dgamma_host || !dbeta_host || !dy_host || !dA || !red)
    return fail(ERR_INVALID, "os2s_bn_bwd: null pointer");
  if (apply_relu && !a) return fail(ERR_INVALID, "os2s_bn_bwd: forward output required for relu backward");
  BnBwdParams p;
  for (int j = 0; j < n_branch; ++j) {
    p.br[j].y = y_host[j];
    p.br[j].mean_invstd = mean_invstd_host[j];
    p.br[j].gamma = gamma_host[j];
    p.br[j].dgamma = dgamma_host[j];
    p.br[j].dbeta = dbeta_host[j];
    p.br[j].dy = dy_host[j];
    p.br[j].ld = ld_host ? ld_host[j] : C;
  }
  p.n_branch = n_branch;
  p.dA = dA; p.dA_is_f32 = dA_is_f32; p.a = a;
  p.y_f32 = (dtypes & OS2S_CONV_F32) ? 1 : 0;
  p.h_f16 = (dtypes & OS2S_HALF_F16) ? 1 : 0;
  p.red = red; p.M = M; p.C = C; p.keep = keep; p.apply_relu = apply_relu;
  OS2S_CUDA(cudaMemsetAsync(red, 0, (size_t)(1 + n_branch) * C * sizeof(float), (cudaStream_t)stream));
  return bn_bwd(p, (cudaStream_t)stream);
}
int os2s_bn_bwd_ld(int n_branch, const void* const* y_host, const int* ld_host,
                   const float* const* mean_invstd_host, const float* const* gamma_host,
                   float* const* dgamma_host, float* const* dbeta_host, void* const* dy_host, const void* dA,
                   int dA_is_f32, const void* a, float* red, int M, int C, float keep, int apply_relu,
                   void* stream) {
  return os2s_bn_bwd_p(n_branch, y_host, ld_host, mean_invstd_host, gamma_host, dgamma_host, dbeta_host, dy_host, dA,
                       dA_is_f32, a, red, M, C, keep, apply_relu, 0, stream);
}

int os2s_bn_bwd_apply_p(const void* y, const float* mean_invstd, const float* gamma, float* dgamma, float* dbeta,
                        void* dy, const void* dA, const void* a, const float* red, int M, int C, float keep,
                        int dtypes, void* stream) {
  if (!y || !mean_invstd || !gamma || !dgamma || !dbeta || !dy || !dA || !a || !red)
    return fail(ERR_INVALID, "os2s_bn_bwd_apply: null pointer");
  BnBwdParams p;
  p.br[0].y = y;
  p.br[0].mean_invstd = mean_invstd;
  p.br[0].gamma = gamma;
  p.br[0].dgamma = dgamma;
  p.br[0].dbeta = dbeta;
  p.br[0].dy = dy;
  p.br[0].ld = C;
  p.n_branch = 1;
  p.dA = dA; p.dA_is_f32 = 0; p.a = a;
  p.y_f32 = (dtypes & OS2S_CONV_F32) ? 1 : 0;
  p.h_f16 = (dtypes & OS2S_HALF_F16) ? 1 : 0;
  p.red = const_cast<float*>(red); p.M = M; p.C = C; p.keep = keep; p.apply_relu = 1;
  return bn_bwd(p, (cudaStream_t)stream, /*reduce=*/false);
}
int os2s_bn_bwd_apply(const void* y, const float* mean_invstd, const float* gamma, float* dgamma, float* dbeta,
                      void* dy, const void* dA, const void* a, const float* red, int M, int C, float keep,
                      void* stream) {
  return os2s_bn_bwd_apply_p(y, mean_invstd, gamma, dgamma, dbeta, dy, dA, a, red, M, C, keep, 0, stream);
}

int os2s_bn_bwd(int n_branch, const void* const* y_host, const float* const* mean_invstd_host,
                const float* const* gamma_host, float* const* dgamma_host, float* const* dbeta_host,
                void* const* dy_host, const void* dA, int dA_is_f32, const void* a, float* red, int M,
                int C, float keep, int apply_relu, void* stream) {
  return os2s_bn_bwd_ld(n_branch, y_host, nullptr, mean_invstd_host, gamma_host, dgamma_host, dbeta_host, dy_host,
                        dA, dA_is_f32, a, red, M, C, keep, apply_relu, stream);
}

int os2s_multi_copy_2d(int n, const void* const* src_host, void* const* dst_host, const int* rows_host,
                       const int* row_bytes_host, const long long* src_pitch_host,
                       const long long* dst_pitch_host, void* stream) {
  if (n < 0 || n > kMaxCopies) return fail(ERR_INVALID, "os2s_multi_copy_2d: 0..64 copies per call");
  if (n == 0) return 0;
  if (!src_host || !dst_host || !rows_host || !row_bytes_host || !src_pitch_host || !dst_pitch_host)
    return fail(ERR_INVALID, "os2s_multi_copy_2d: null pointer");
  Copy2dTable tab;
  tab.n = n;
  for (int i = 0; i < n; ++i) {
    if (!src_host[i] || !dst_host[i]) return fail(ERR_INVALID, "os2s_multi_copy_2d: null pointer");
    if (((uintptr_t)src_host[i] | (uintptr_t)dst_host[i] | (uintptr_t)row_bytes_host[i] |
         (uintptr_t)src_pitch_host[i] | (uintptr_t)dst_pitch_host[i]) & 15)
      return fail(ERR_INVALID, "os2s_multi_copy_2d: addresses, row bytes and pitches must be multiples of 16");
    if (rows_host[i] < 0 || row_bytes_host[i] < 0) return fail(ERR_INVALID, "os2s_multi_copy_2d: negative extent");
    tab.src[i] = (const char*)src_host[i];
    tab.dst[i] = (char*)dst_host[i];
    tab.rows[i] = rows_host[i];
    tab.row_vecs[i] = row_bytes_host[i] / 16;
    tab.src_pitch[i] = src_pitch_host[i];
    tab.dst_pitch[i] = dst_pitch_host[i];
  }
  return multi_copy_2d(tab, (cudaStream_t)stream);
}

int os2s_fc_fwd_p(const void* x, const float* w, const float* bias, float* logits, int M, int H, int V,
                  int dtypes, void* stream) {
  if (!x || !w || !logits) return fail(ERR_INVALID, "os2s_fc_fwd: null pointer");
  return fc_fwd(x, w, bias, logits, M, H, V, (cudaStream_t)stream, (dtypes & OS2S_HALF_F16) ? 1 : 0);
}
int os2s_fc_fwd(const void* x, const float* w, const float* bias, float* logits, int M, int H, int V,
                void* stream) {
  return os2s_fc_fwd_p(x, w, bias, logits, M, H, V, 0, stream);
}

int os2s_fc_bwd_p(const void* x, const float* dlogits, const float* w, void* dx, float* dw, float* db,
                  int M, int H, int V, int dtypes, void* stream) {
  if (!x || !dlogits || !w) return fail(ERR_INVALID, "os2s_fc_bwd: null pointer");
  if ((dw == nullptr) != (db == nullptr)) return fail(ERR_INVALID, "os2s_fc_bwd: dw and db go together");
  return fc_bwd(x, dlogits, w, dx, dw, db, M, H, V, (cudaStream_t)stream, (dtypes & OS2S_HALF_F16) ? 1 : 0);
}
int os2s_fc_bwd(const void* x, const float* dlogits, const float* w, void* dx, float* dw, float* db,
                int M, int H, int V, void* stream) {
  return os2s_fc_bwd_p(x, dlogits, w, dx, dw, db, M, H, V, 0, stream);
}

size_t os2s_ctc_workspace_bytes(int B, int T, int L_max) { return ctc_workspace_bytes(B, T, L_max); }

int os2s_ctc_loss_fwd_bwd(const float* logits, const int* labels, const int* label_lens,
                          const int* input_lens, float* grad, float* loss, void* workspace,
                          size_t workspace_bytes, const float* loss_scale_dev, int B, int T, int V,
                          int L_max, long long stride_b, long long stride_t, void* stream) {
  if (!logits || !labels || !label_lens || !input_lens || !grad || !loss || !workspace)
    return fail(ERR_INVALID, "os2s_ctc_loss_fwd_bwd: null pointer");
  if (B <= 0 || T <= 0 || V < 2 || L_max < 0) return fail(ERR_INVALID, "os2s_ctc_loss_fwd_bwd: bad shape");
  return ctc_loss_fwd_bwd(logits, labels, label_lens, input_lens, grad, loss, (float*)workspace, workspace_bytes,
                          loss_scale_dev, B, T, V, L_max, stride_b, stride_t, (cudaStream_t)stream);
}

int os2s_ctc_greedy(const float* logits, const int* input_lens, int* tokens, int* out_lens,
                    float* neg_sum_logits, int B, int T, int V, long long stride_b,
                    long long stride_t, int merge_repeated, void* stream) {
  if (!logits || !input_lens || !tokens || !out_lens) return fail(ERR_INVALID, "os2s_ctc_greedy: null pointer");
  return ctc_greedy(logits, input_lens, tokens, out_lens, neg_sum_logits, B, T, V, stride_b, stride_t,
                    merge_repeated, (cudaStream_t)stream);
}

int os2s_opt_chunk_elems(void) { return opt_chunk_elems(); }

int os2s_opt_step3(void* const* w, void* const* g, void* const* m, void* const* v, void* const* wb,
                   const float* reg, const int* frozen, const long long* sizes, const int* chunk_tensor,
                   const long long* chunk_offset, int n_tensors, int n_chunks, const os2s_opt_hparams* hp,
                   float* norms,
                   int* nonfinite, float* fstate, long long* istate, float* coef, float* ema,
                   void* stream) {
  if (!w || !g || !m || !wb || !sizes || !chunk_tensor || !chunk_offset || !hp || !norms || !nonfinite ||
      !fstate || !istate || !coef || !ema)
    return fail(ERR_INVALID, "os2s_opt_step: null pointer");
  if (hp->world_size < 1) return fail(ERR_INVALID, "os2s_opt_step: world_size < 1");
  if (hp->algo < 0 || hp->algo > 2) return fail(ERR_INVALID, "os2s_opt_step: unknown algorithm");
  if (hp->algo == 2 && !v) return fail(ERR_INVALID, "os2s_opt_step: Adam needs the second-moment array (os2s_opt_step2)");
  if (hp->lr_policy < 0 || hp->lr_policy > 3) return fail(ERR_INVALID, "os2s_opt_step: unknown lr policy");
  if (hp->max_grad_norm > 0.f && hp->larc_eta > 0.f)
    return fail(ERR_INVALID, "os2s_opt_step: LARC and gradient norm clipping should not be used together");
  OptTable tab{w, g, m, wb, v, reg, frozen, sizes, chunk_tensor, chunk_offset, n_tensors, n_chunks};
  return opt_step(tab, *hp, norms, nonfinite, fstate, istate, coef, ema, (cudaStream_t)stream);
}
int os2s_opt_step2(void* const* w, void* const* g, void* const* m, void* const* v, void* const* wb,
                   const float* reg, const long long* sizes, const int* chunk_tensor,
                   const long long* chunk_offset, int n_tensors, int n_chunks, const os2s_opt_hparams* hp,
                   float* norms,
                   int* nonfinite, float* fstate, long long* istate, float* coef, float* ema,
                   void* stream) {
  return os2s_opt_step3(w, g, m, v, wb, reg, nullptr, sizes, chunk_tensor, chunk_offset, n_tensors, n_chunks, hp, norms,
                        nonfinite, fstate, istate, coef, ema, stream);
}

int os2s_opt_step(void* const* w, void* const* g, void* const* m, void* const* wb,
                  const long long* sizes, const int* chunk_tensor, const long long* chunk_offset,
                  int n_tensors, int n_chunks, const os2s_opt_hparams* hp, float* norms,
                  int* nonfinite, float* fstate, long long* istate, float* coef, float* ema,
                  void* stream) {
  return os2s_opt_step2(w, g, m, nullptr, wb, nullptr, sizes, chunk_tensor, chunk_offset, n_tensors, n_chunks, hp, norms,
                        nonfinite, fstate, istate, coef, ema, stream);
}

int os2s_multi_transpose(void* const* src, void* const* dst, const int* Rdev, const int* Cdev,
                         const long long* tile_start, int n_tensors, long long total_tiles,
                         void* stream) {
  if (!src || !dst || !Rdev || !Cdev || !tile_start) return fail(ERR_INVALID, "os2s_multi_transpose: null pointer");
  TransposeTable tab{src, dst, Rdev, Cdev, tile_start, n_tensors};
  return multi_transpose(tab, total_tiles, (cudaStream_t)stream);
}

int os2s_logmel_forward(const int16_t* wave, const long long* offsets, const int* n_samples, int B,
                        const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop, int F,
                        int T_pad, int max_samples, float dither, uint64_t seed, float preemph,
                        void* absmax_ws, float* raw_ws, void* out_bf16, float* out_f32, int* out_lens,
                        void* stream) {
  if (!wave || !offsets || !n_samples || !mel || !window || !absmax_ws || !raw_ws)
    return fail(ERR_INVALID, "os2s_logmel_forward: null pointer");
  if (!out_bf16 && !out_f32) return fail(ERR_INVALID, "os2s_logmel_forward: no output buffer");
  return logmel_forward(wave, offsets, n_samples, B, mel, mel_band, window, n_fft, win, hop, F, T_pad, max_samples, dither,
                        seed, preemph, (unsigned int*)absmax_ws, raw_ws, out_bf16, out_f32, out_lens,
                        (cudaStream_t)stream);
}

int os2s_features_forward(const int16_t* wave, const long long* offsets, const int* n_samples, int B,
                          const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop,
                          int F, int T_pad, int max_samples, float dither, uint64_t seed, float preemph,
                          int psf_backend, int pad_to, int norm_per_feature, void* absmax_ws, float* raw_ws,
                          void* out_bf16, float* out_f32, int* out_lens, void* stream) {
  if (!wave || !offsets || !n_samples || !mel || !window || !absmax_ws || !raw_ws)
    return fail(ERR_INVALID, "os2s_features_forward: null pointer");
  if (!out_bf16 && !out_f32) return fail(ERR_INVALID, "os2s_features_forward: no output buffer");
  if (psf_backend && dither != 0.f) return fail(ERR_INVALID, "os2s_features_forward: the psf backend has no dither");
  if (pad_to < 0) return fail(ERR_INVALID, "os2s_features_forward: pad_to < 0");
  return logmel_forward(wave, offsets, n_samples, B, mel, mel_band, window, n_fft, win, hop, F, T_pad, max_samples, dither,
                        seed, preemph, (unsigned int*)absmax_ws, raw_ws, out_bf16, out_f32, out_lens,
                        (cudaStream_t)stream, psf_backend, pad_to, norm_per_feature);
}

int os2s_sepconv_compose(const float* depthwise, const float* pointwise, void* w_half, int K, int C_in, int C_out,
                         int dtypes, void* stream) {
  if (!depthwise || !pointwise || !w_half) return fail(ERR_INVALID, "os2s_sepconv_compose: null pointer");
  return sepconv_compose(depthwise, pointwise, w_half, K, C_in, C_out, (dtypes & OS2S_HALF_F16) ? 1 : 0,
                         (cudaStream_t)stream);
}

int os2s_sepconv_decompose_grad(const float* dw_dense, const float* depthwise, const float* pointwise,
                                float* d_depthwise, float* d_pointwise, int K, int C_in, int C_out, void* stream) {
  if (!dw_dense || !depthwise || !pointwise || !d_depthwise || !d_pointwise)
    return fail(ERR_INVALID, "os2s_sepconv_decompose_grad: null pointer");
  return sepconv_decompose_grad(dw_dense, depthwise, pointwise, d_depthwise, d_pointwise, K, C_in, C_out,
                                (cudaStream_t)stream);
}

int os2s_depthwise_conv1d(const void* x, const float* taps, void* out, int B, int T, int C, int K, int t_off0,
                          int t_step, int out_mode, int dtypes, void* stream) {
  if (!x || !taps || !out) return fail(ERR_INVALID, "os2s_depthwise_conv1d: null pointer");
  int mode;
  if (out_mode == OS2S_OUT_F32) mode = 1;
  else if (out_mode == OS2S_OUT_F32_ACC) mode = 2;
  else if (out_mode == OS2S_OUT_BF16 || out_mode == OS2S_OUT_F16_GRAD) mode = 0;
  else return fail(ERR_INVALID, "os2s_depthwise_conv1d: out_mode is the 16-bit format of the mode, F32 or F32_ACC");
  return depthwise_conv1d(x, taps, out, B, T, C, K, t_off0, t_step, mode, (dtypes & OS2S_HALF_F16) ? 1 : 0,
                          (cudaStream_t)stream);
}

int os2s_depthwise_conv1d_wgrad(const void* x, const void* dz, float* d_taps, int B, int T, int C, int K, int dil,
                                int pad_left, int dtypes, void* stream) {
  if (!x || !dz || !d_taps) return fail(ERR_INVALID, "os2s_depthwise_conv1d_wgrad: null pointer");
  return depthwise_conv1d_wgrad(x, dz, d_taps, B, T, C, K, dil, pad_left, (dtypes & OS2S_HALF_F16) ? 1 : 0,
                                (cudaStream_t)stream);
}

int os2s_wave_absmax(const int16_t* wave, const long long* offsets, const int* n_samples, int B, void* absmax,
                     void* stream) {
  if (!wave || !offsets || !n_samples || !absmax) return fail(ERR_INVALID, "os2s_wave_absmax: null pointer");
  return wave_absmax(wave, offsets, n_samples, B, (unsigned int*)absmax, (cudaStream_t)stream);
}

int os2s_augment_signal(const int16_t* wave, const long long* offsets, const int* n_samples, int B,
                        const void* absmax, float gain, const int* sr_new, int sr_orig, const float* interp_win,
                        int n_win, int num_table, const float* noise_amp, uint64_t seed, float* out,
                        const long long* out_offsets, const int* n_out, int max_out, void* stream) {
  if (!wave || !offsets || !n_samples || !out || !out_offsets || !n_out)
    return fail(ERR_INVALID, "os2s_augment_signal: null pointer");
  if (!absmax && !(gain > 0.f)) return fail(ERR_INVALID, "os2s_augment_signal: needs absmax or a fixed gain");
  return augment_signal(wave, offsets, n_samples, B, (const unsigned int*)absmax, gain, sr_new, sr_orig, interp_win,
                        n_win, num_table, noise_amp, seed, out, out_offsets, n_out, max_out, (cudaStream_t)stream);
}

int os2s_features_forward_p(const int16_t* wave, const float* sig, const long long* sig_offsets,
                            const long long* offsets, const int* n_samples, int B,
                            const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop,
                            int F, int T_pad, int max_samples, float dither, uint64_t seed, float preemph,
                            int psf_backend, int pad_to, int norm_per_feature, float gain,
                            const float* features_mean, const float* features_std, const int* masks, int n_masks,
                            int feature_type, const float* mfcc_matrix, int n_filt,
                            void* absmax_ws, float* raw_ws, void* out16, float* out_f32, int* out_lens,
                            int dtypes, void* stream) {
  if ((!wave && !sig) || !offsets || !n_samples || !mel || !window || !absmax_ws || !raw_ws)
    return fail(ERR_INVALID, "os2s_features_forward: null pointer");
  if (!out16 && !out_f32) return fail(ERR_INVALID, "os2s_features_forward: no output buffer");
  if (psf_backend && dither != 0.f) return fail(ERR_INVALID, "os2s_features_forward: the psf backend has no dither");
  if (pad_to < 0 || n_masks < 0 || (n_masks > 0 && !masks)) return fail(ERR_INVALID, "os2s_features_forward: bad argument");
  FeatExtras ex{gain, sig, sig_offsets, features_mean, features_std, masks, n_masks, feature_type, mfcc_matrix, n_filt};
  return logmel_forward(wave, offsets, n_samples, B, mel, mel_band, window, n_fft, win, hop, F, T_pad, max_samples, dither,
                        seed, preemph, (unsigned int*)absmax_ws, raw_ws, out16, out_f32, out_lens,
                        (cudaStream_t)stream, psf_backend, pad_to, norm_per_feature, (dtypes & OS2S_HALF_F16) ? 1 : 0,
                        &ex);
}

// ---- gradient exchange over NVLink peer memory (peer.cu) ----
long long os2s_peer_stage_bytes(int world, int n_buckets, const long long* bucket_start_host,
                                const long long* bucket_end_host) {
  if (world < 2 || n_buckets < 1 || !bucket_start_host || !bucket_end_host) return -1;
  return peer_stage_bytes(world, n_buckets, bucket_start_host, bucket_end_host);
}

int os2s_ipc_export(const void* ptr, unsigned char* handle64_host, long long* offset_host) {
  if (!ptr || !handle64_host || !offset_host) return fail(ERR_INVALID, "os2s_ipc_export: null pointer");
  return ipc_export(ptr, handle64_host, offset_host);
}

int os2s_ipc_open(const unsigned char* handle64_host, void** base_host) {
  if (!handle64_host || !base_host) return fail(ERR_INVALID, "os2s_ipc_open: null pointer");
  return ipc_open(handle64_host, base_host);
}

int os2s_ipc_close(void* base) {
  if (!base) return fail(ERR_INVALID, "os2s_ipc_close: null pointer");
  return ipc_close(base);
}

int os2s_peer_create(int rank, int world, void* const* grad_host, void* const* stage_host, int n_buckets,
                     const long long* bucket_start_host, const long long* bucket_end_host, double timeout_s,
                     void** ctx_host) {
  if (!grad_host || !stage_host || !bucket_start_host || !bucket_end_host || !ctx_host)
    return fail(ERR_INVALID, "os2s_peer_create: null pointer");
  PeerExchange* px = nullptr;
  int s = peer_create(rank, world, grad_host, stage_host, n_buckets, bucket_start_host, bucket_end_host, timeout_s, &px);
  *ctx_host = px;
  return s;
}

int os2s_peer_destroy(void* ctx) {
  peer_destroy((PeerExchange*)ctx);
  return OK;
}

int os2s_peer_set_timeout(void* ctx, double timeout_s) { return peer_set_timeout((PeerExchange*)ctx, timeout_s); }

int os2s_peer_exchange_bucket(void* ctx, int bucket, void* stream) {
  return peer_exchange_bucket((PeerExchange*)ctx, bucket, (cudaStream_t)stream);
}

int os2s_peer_finish(void* ctx, void* stream) { return peer_finish((PeerExchange*)ctx, (cudaStream_t)stream); }

int os2s_peer_timed_out(void* ctx, int* flag_host) { return peer_timed_out((PeerExchange*)ctx, flag_host); }

}  // extern "C"
