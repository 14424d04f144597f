// Internal (C++) declarations of the kernel launchers behind include/os2s.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/os2s.h"

namespace os2s {

constexpr int kMaxBranches = 12;

const char* last_error_cstr();

// conv_tc.cu
int conv_tuning(int pair_mode, int halo_mode);
int conv_grid_waves_set(int waves);
int conv_kmajor(const void* act, const void* wmat, void* out, int B, int T, int C_red, int N_total,
                int K, int t_off0, int t_step, int out_mode, int b_mn_major, float* stats, cudaStream_t st,
                const void* bwd_a = nullptr, const void* bwd_y = nullptr, float bwd_inv_keep = 1.f,
                int act_f16 = 0, int bwd_y_f32 = 0, const int* row_lens = nullptr, int skip_margin = 0);
int conv_wgrad(const void* x, const void* dy, float* dw, int B, int T, int C_in, int C_out, int K,
               int dil, int pad_left, int* splits_used, cudaStream_t st, int x_f16 = 0, const int* row_lens = nullptr);

// elementwise.cu
int weight_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int K, int C_in, int C_out,
                          cudaStream_t st, int f16 = 0);
struct BnBranchFwd {
  const void* y;        // conv output, fp16 (fp32 when BnFwdParams::y_f32)
  const float* stats;   // [2][C] sums
  const float* gamma;
  const float* beta;
  float* mean_invstd;   // [2][C] saved for backward
  float* moving;        // [2][C] moving_mean, moving_variance (may be null)
  int ld;               // row stride of y in elements (C, or wider when y is a column slice)
  int stats_ld;         // offset of the sum-of-squares row from the sum row in `stats` (C by default)
};
struct BnFwdParams {
  BnBranchFwd br[kMaxBranches];
  int n_branch;
  void* out;            // bf16 (fp16 when out_f16)
  int y_f32, out_f16;
  const int* lens;      // [B] valid rows per utterance (nullptr = no mask)
  int B, T, C;
  float eps, momentum;  // momentum as in TF: moving = moving*momentum + batch*(1-momentum)
  float keep;           // dropout keep probability (1 = off)
  unsigned long long seed;
  float relu_clip;      // <= 0: plain relu, > 0: min(relu(x), clip)
  int apply_relu;
  int use_moving;       // 1 = inference mode: normalise with the moving statistics
  const long long* step_ctr;  // optional device counter mixed into the dropout seed (CUDA-graph replays)
};
struct BnBranchBwd {
  const void* y;            // conv output, fp16 (fp32 when BnBwdParams::y_f32)
  const float* mean_invstd;  // [2][C]
  const float* gamma;
  float* dgamma;             // [C] gradient outputs (scaled by loss scale like dA)
  float* dbeta;              // [C]
  void* dy;                  // [M, C] bf16 (fp16 when BnBwdParams::h_f16)
  int ld;                    // row stride of y AND dy in elements
};
struct BnBwdParams {
  BnBranchBwd br[kMaxBranches];
  int n_branch;
  const void* dA;            // bf16 or fp32 [M, C]
  int dA_is_f32;
  int y_f32;
  int h_f16;                 // dA (when 16-bit) and every dy are fp16 instead of bf16
  const void* a;             // forward output of this layer (post relu/dropout/mask), bf16 or fp16: only its zeros matter
  float* red;                // [1 + n_branch][C] fp32 scratch, pre-zeroed: dbeta, dgamma_j
  int M, C;
  float keep;
  int apply_relu;            // if 0: dz = dA (no activation), "a" unused
};

constexpr int kMaxCopies = 64;
struct Copy2dTable {
  const char* src[kMaxCopies];
  char* dst[kMaxCopies];
  long long src_pitch[kMaxCopies], dst_pitch[kMaxCopies];
  int rows[kMaxCopies], row_vecs[kMaxCopies];   // row length in 16-byte vectors
  int n;
};
int multi_copy_2d(const Copy2dTable& tab, cudaStream_t st);
int bn_stats(const void* y, float* stats, int M, int C, cudaStream_t st);
int bn_apply_fwd(const BnFwdParams& p, cudaStream_t st);
int bn_bwd(const BnBwdParams& p, cudaStream_t st, bool reduce = true);


// sepconv.cu
int sepconv_compose(const float* D, const float* P, void* w, int K, int C, int Co, int f16, cudaStream_t st);
int sepconv_decompose_grad(const float* dW, const float* D, const float* P, float* dD, float* dP, int K, int C, int Co,
                           cudaStream_t st);
int depthwise_conv1d(const void* x, const float* taps, void* out, int B, int T, int C, int K, int off0, int step,
                     int out_mode, int f16, cudaStream_t st);
int depthwise_conv1d_wgrad(const void* x, const void* dz, float* dtaps, int B, int T, int C, int K, int dil, int pad,
                           int f16, cudaStream_t st);

// ctc.cu
int fc_fwd(const void* x, const float* w, const float* bias, float* logits, int M, int H, int V, cudaStream_t st,
           int x_f16 = 0);
int fc_bwd(const void* x, const float* dl, const float* w, void* dx, float* dw, float* db, int M, int H, int V,
           cudaStream_t st, int x_f16 = 0);
size_t ctc_workspace_bytes(int B, int T, int L_max);
int ctc_loss_fwd_bwd(const float* logits, const int* labels, const int* label_lens, const int* input_lens,
                     float* grad, float* loss, float* workspace, size_t workspace_bytes, const float* loss_scale,
                     int B, int T, int V, int L_max, long long stride_b, long long stride_t, cudaStream_t st);
int ctc_greedy(const float* logits, const int* input_lens, int* tokens, int* out_lens, float* neg_sum, int B,
               int T, int V, long long stride_b, long long stride_t, int merge_repeated, cudaStream_t st);

// optim.cu
typedef os2s_opt_hparams OptHParams;
struct OptTable {
  void* const* w;
  void* const* g;
  void* const* m;
  void* const* wb;
  void* const* v;   // Adam second moments (nullptr otherwise)
  const float* reg; // per-tensor L2-regulariser scale (nullptr = none)
  const int* frozen; // per-tensor flag: excluded from the update and from the global norm (nullptr = none)
  const long long* sizes;
  const int* chunk_tensor;
  const long long* chunk_offset;
  int n_tensors, n_chunks;
};
struct TransposeTable {
  void* const* src;
  void* const* dst;
  const int* R;
  const int* C;
  const long long* tile_start;
  int n_tensors;
};
int opt_step(const OptTable& tab, const OptHParams& hp, float* norms, int* nonfinite, float* fstate,
             long long* istate, float* coef, float* ema, cudaStream_t st);
int opt_chunk_elems();
int multi_transpose(const TransposeTable& tab, long long total_tiles, cudaStream_t st);

// feat.cu / augment.cu
struct FeatExtras {
  float fixed_gain;
  const float* sig;
  const long long* sig_off;
  const float* fixed_mean;
  const float* fixed_std;
  const int* masks;
  int n_masks;
  int feature_type;          // 0 = logfbank, 1 = psf spectrogram, 2 = psf mfcc
  const float* post;         // mfcc: [F][n_filt]
  int n_filt;
};
int wave_absmax(const short* wave, const long long* offsets, const int* n_samples, int B, unsigned int* absmax,
                cudaStream_t st);
int augment_signal(const short* wave, const long long* offsets, const int* n_in, int B, const unsigned int* absmax,
                   float fixed_gain, const int* sr_new, int sr_orig, const float* win, int nwin, int num_table,
                   const float* noise_amp, unsigned long long seed, float* out, const long long* out_offsets,
                   const int* n_out, int max_out, cudaStream_t st);
int logmel_forward(const short* wave, const long long* offsets, const int* n_samples, int B,
                   const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop, int F, int T_pad,
                   int max_samples, float dither, unsigned long long seed, float preemph,
                   unsigned int* absmax_ws, float* raw_ws, void* out_bf16, float* out_f32, int* out_lens,
                   cudaStream_t st,
                   int psf_backend = 0, int pad_to = 0, int norm_per_feature = 1, int out_f16 = 0,
                   const FeatExtras* ex = nullptr);

// ---- peer.cu: gradient sum over NVLink peer memory (CUDA IPC + copy engines) ----
struct PeerExchange;
long long peer_stage_bytes(int world, int n_buckets, const long long* start, const long long* end);
int ipc_export(const void* ptr, unsigned char* handle, long long* offset);
int ipc_open(const unsigned char* handle, void** base);
int ipc_close(void* base);
int peer_create(int rank, int world, void* const* grad, void* const* stage, int n_buckets, const long long* start,
                const long long* end, double timeout_s, PeerExchange** out);
void peer_destroy(PeerExchange* px);
int peer_set_timeout(PeerExchange* px, double timeout_s);
int peer_exchange_bucket(PeerExchange* px, int b, cudaStream_t st);
int peer_finish(PeerExchange* px, cudaStream_t st);
int peer_timed_out(PeerExchange* px, int* flag);

}  // namespace os2s
