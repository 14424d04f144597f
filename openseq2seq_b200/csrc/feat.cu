// Speech2TextDataLayer log-mel featurizer on the GPU (librosa backend, input_type="logfbank").
//
// Reference being replaced (NumPy + librosa on 8 CPU py_func threads):
//   open_seq2seq/data/speech2text/speech_utils.py:216-222   normalize_signal
//   open_seq2seq/data/speech2text/speech_utils.py:364-365   dither
//   open_seq2seq/data/speech2text/speech_utils.py:271-272   preemphasis
//   open_seq2seq/data/speech2text/speech_utils.py:396-406   STFT(512, hop 160, hann 320) -> |.|^2 -> mel -> log
//   open_seq2seq/data/speech2text/speech_utils.py:411-417   per-feature mean / std over time
//   open_seq2seq/data/speech2text/speech2text.py:251-257,313-317  zero padding to [B, T_pad, F]
// Kernel 1: per-utterance max |x| (gain).  Kernel 2: one warp per frame: gather the reflect-padded,
// pre-emphasised, windowed frame into shared memory, radix-2 FFT, power spectrum, dense mel
// mat-vec, log.  Kernel 3: one warp per (utterance, feature): mean / population std over the
// utterance's frames with warp-shuffle reductions, normalise, cast to bf16, zero the padding.
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace os2s {

__global__ void feat_absmax_kernel(const short* __restrict__ wave, const long long* __restrict__ offsets,
                                   const int* __restrict__ n_samples, unsigned int* __restrict__ absmax) {
  const int b = blockIdx.y;
  const short* w = wave + offsets[b];
  const int n = n_samples[b];
  int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = w[i];
    m = max(m, v < 0 ? -v : v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&absmax[b], (unsigned int)m);
}

__device__ __forceinline__ float gauss_from_index(unsigned long long seed, unsigned long long idx) {
  // counter-based N(0,1): two splitmix64 uniforms -> Box-Muller
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u1 = ((float)((z >> 40) & 0xFFFFFF) + 1.f) * (1.f / 16777217.f);
  const float u2 = (float)((z >> 8) & 0xFFFFFF) * (1.f / 16777216.f);
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.28318530718f * u2);
}

constexpr int kNfftMax = 512;
constexpr int kFeatWarps = 4;

// Backend differences (speech_utils.py:322-441 librosa vs :444-535 python_speech_features):
//   librosa: frames = 1 + n/hop, centred with reflect padding, float signal (+ dither), log(x + 1e-20),
//            per-feature normalisation;
//   psf    : signal re-quantised to int16 and zero-padded so that frames % pad_to == 0 (the padding is
//            pre-emphasised with the signal), frame f = samples [f*hop, f*hop + win), power / n_fft,
//            zeros -> eps before the log, ONE mean / std per utterance.
struct FeatOpts {
  int psf;            // 0 = librosa conventions, 1 = psf conventions
  int pad_to;         // psf: frames per utterance are a multiple of this
  int per_feature;    // normalisation: 1 = per feature over time, 0 = one mean/std per utterance
  float power_scale;  // multiplies |X|^2
  int out_f16;        // the 16-bit output is fp16 instead of bf16 (OS2S_HALF_F16)
  float fixed_gain;   // > 0: params['gain'] instead of 1 / (max|x| + 1e-5) (speech_utils.py:216-222)
  // optional device arrays
  const float* sig;          // augmented float signal (already normalised): replaces wave * gain
  const long long* sig_off;  // [B] offsets into sig
  const float* fixed_mean;   // [F] params['features_mean'] (speech_utils.py:411-417), per-feature norm only
  const float* fixed_std;    // [F] params['features_std_dev']
  const int* masks;          // [B][n_masks][3] = (kind 0 freq / 1 time, base, width): spec-augment (:419-433)
  int n_masks;
  // psf feature types other than logfbank (speech_utils.py:490-512)
  int feature_type;          // 0 = logfbank, 1 = spectrogram (NFFT = window length DFT), 2 = mfcc
  const float* post;         // mfcc: [F][n_filt] = lifter * orthonormal DCT-II rows applied to the log energies
  int n_filt;                // mfcc: rows of `mel` (= 2 F)
};
// spec-augment: zeros written into the normalised features
__device__ __forceinline__ bool feat_masked(const FeatOpts& o, int b, int t, int f) {
  if (!o.masks) return false;
  const int* m = o.masks + (size_t)b * o.n_masks * 3;
  bool hit = false;
  for (int i = 0; i < o.n_masks; ++i) {
    const int kind = m[3 * i], base = m[3 * i + 1], width = m[3 * i + 2];
    const int x = kind == 0 ? f : t;
    hit |= (x >= base && x < base + width);
  }
  return hit;
}
__device__ __forceinline__ void store_act16(uint16_t* dst, size_t o, float v, int f16) {
  dst[o] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16(v));
}
__device__ __forceinline__ int frames_of(const FeatOpts& o, int n, int hop, int win) {
  if (!o.psf) return 1 + n / hop;
  int len = 1 + (n - win + hop - 1) / hop;      // 1 + ceil((n - win) / hop) for n >= win
  if (n <= win) len = 1;
  if (o.pad_to > 0 && len % o.pad_to) len += o.pad_to - len % o.pad_to;
  return len;
}

// signal value at (reflect-resolved) sample index i: normalised, dithered, then pre-emphasised.
__device__ __forceinline__ float sample_at(const short* w, const float* sig, int n, int i, float gain, float dither,
                                           unsigned long long seed, float preemph) {
  auto base = [&](int j) {
    float v = sig ? sig[j] : (float)w[j] * gain;
    if (dither > 0.f) v += dither * gauss_from_index(seed, (unsigned long long)j);
    return v;
  };
  const float cur = base(i);
  return (i == 0) ? cur : cur - preemph * base(i - 1);
}

// psf: q[i] = (int16)(float(x[i] * gain) * 32767) (truncation), zero for i >= n; the zero padding up to
// n_padded belongs to the signal when it is pre-emphasised, so sample n is -preemph * q[n-1].
__device__ __forceinline__ float sample_at_psf(const short* w, int n, int n_padded, int i, float gain, float preemph) {
  if (i < 0 || i >= n_padded) return 0.f;
  const float cur = (i < n) ? truncf(((float)w[i] * gain) * 32767.0f) : 0.f;
  if (i == 0) return cur;
  const float prev = (i - 1 < n) ? truncf(((float)w[i - 1] * gain) * 32767.0f) : 0.f;
  return cur - preemph * prev;
}

template <int NFFT>
__global__ void __launch_bounds__(kFeatWarps * 32)
feat_logmel_kernel(const short* __restrict__ wave, const long long* __restrict__ offsets,
                   const int* __restrict__ n_samples, const unsigned int* __restrict__ absmax,
                   const float* __restrict__ mel, const int* __restrict__ mel_band,
                   const float* __restrict__ window, float* __restrict__ raw,
                   int T_pad, int F, int hop, int win, float dither, unsigned long long seed, float preemph,
                   const FeatOpts opts) {
  constexpr int NB = NFFT / 2 + 1;
  __shared__ float re[kFeatWarps][NFFT];
  __shared__ float im[kFeatWarps][NFFT];
  __shared__ float2 tw[NFFT / 2];  // exp(-2 pi i k / NFFT), computed once per CTA
  for (int k = threadIdx.x; k < NFFT / 2; k += kFeatWarps * 32) {
    float sn, cs;
    sincospif(-2.f * (float)k / (float)NFFT, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int frame = blockIdx.x * kFeatWarps + warp;
  const int n = n_samples[b];
  const int n_frames = frames_of(opts, n, hop, win);
  if (frame >= n_frames) return;  // whole warp exits together
  // psf: the zero padding that makes the frame count a multiple of pad_to is part of the signal
  // (speech_utils.py:481-488 pads the int16 signal before psf pre-emphasises it; framesig's own zero
  // padding of the last frame comes after the pre-emphasis)
  int n_padded = n;
  if (opts.psf) {
    const int len0 = (n <= win) ? 1 : 1 + (n - win + hop - 1) / hop;
    n_padded = n + (n_frames - len0) * hop;
  }
  const short* w = wave + offsets[b];
  const float* sg = opts.sig ? opts.sig + opts.sig_off[b] : nullptr;
  const float gain = opts.fixed_gain > 0.f ? opts.fixed_gain : 1.f / ((float)absmax[b] + 1e-5f);
  const unsigned long long useed = seed + (unsigned long long)b * 0x632BE59BD9B4E019ull;
  const int lpad = (NFFT - win) / 2;
  // bit-reversed load of the windowed frame (center=True: frame starts at frame*hop - NFFT/2)
  constexpr int LOG2N = (NFFT == 512) ? 9 : (NFFT == 256) ? 8 : 10;
  for (int i = lane; i < NFFT; i += 32) {
    float v = 0.f;
    const int wi = i - lpad;
    if (wi >= 0 && wi < win) {
      if (opts.psf) {
        // frame f = samples [f*hop, f*hop + win); the position inside the FFT buffer does not change |X|
        v = sample_at_psf(w, n, n_padded, frame * hop + wi, gain, preemph) * window[wi];
      } else {
        int j = frame * hop - NFFT / 2 + i;
        if (j < 0) j = -j;                      // np.pad(mode="reflect")
        if (j >= n) j = 2 * (n - 1) - j;
        j = min(max(j, 0), n - 1);
        v = sample_at(w, sg, n, j, gain, dither, useed, preemph) * window[wi];
      }
    }
    const int r = __brev((unsigned)i) >> (32 - LOG2N);
    re[warp][r] = v;
    im[warp][r] = 0.f;
  }
  __syncwarp();
  // iterative radix-2 DIT FFT, NFFT/2 butterflies per stage spread over the warp
  for (int s = 1; s <= LOG2N; ++s) {
    const int half = 1 << (s - 1);
    for (int i = lane; i < NFFT / 2; i += 32) {
      const int grp = i / half, k = i - grp * half;
      const int i0 = grp * 2 * half + k, i1 = i0 + half;
      const float2 w2 = tw[k * (NFFT / 2 / half)];
      const float cs = w2.x, sn = w2.y;
      const float xr = re[warp][i1], xi = im[warp][i1];
      const float tr = xr * cs - xi * sn, ti = xr * sn + xi * cs;
      const float ur = re[warp][i0], ui = im[warp][i0];
      re[warp][i0] = ur + tr;
      im[warp][i0] = ui + ti;
      re[warp][i1] = ur - tr;
      im[warp][i1] = ui - ti;
    }
    __syncwarp();
  }
  // power spectrum into re[0..NB)
  for (int i = lane; i < NB; i += 32) {
    const float a = re[warp][i], c = im[warp][i];
    re[warp][i] = a * a + c * c;
  }
  __syncwarp();
  // mel mat-vec: filter f = lane, lane+32; mel is [n_filt][NB] row-major (dense)
  const int n_filt = opts.feature_type == 2 ? opts.n_filt : F;
  for (int f = lane; f < n_filt; f += 32) {
    const float* mrow = mel + (size_t)f * NB;
    // triangular filters are zero outside [lo, hi): only the support is visited when bands are given
    const int lo = mel_band ? mel_band[2 * f] : 0;
    const int hi = mel_band ? mel_band[2 * f + 1] : NB;
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc += mrow[k] * re[warp][k];
    acc *= opts.power_scale;
    // librosa path: log(S + 1e-20) (speech_utils.py:406); psf: zeros -> float eps, then log (base.fbank)
    const float lg = opts.psf ? logf(acc == 0.f ? 2.220446049250313e-16f : acc) : logf(acc + 1e-20f);
    if (opts.feature_type == 2) im[warp][f] = lg;       // mfcc: keep the log energies for the DCT
    else raw[((size_t)b * T_pad + frame) * F + f] = lg;
  }
  if (opts.feature_type == 2) {
    // psf.mfcc (speech_utils.py:504-512): cepstrum c = lifter * DCT-II_ortho(log energies), first F coefficients
    __syncwarp();
    for (int f = lane; f < F; f += 32) {
      const float* prow = opts.post + (size_t)f * n_filt;
      float acc = 0.f;
      for (int j = 0; j < n_filt; ++j) acc += prow[j] * im[warp][j];
      raw[((size_t)b * T_pad + frame) * F + f] = acc;
    }
  }
}

// psf spectrogram (speech_utils.py:490-502): frame f = samples [f*hop, f*hop + win) of the re-quantised, zero-padded
// signal times the window, |DFT_win|^2 / win with NFFT = the window length (320: not a power of two, so a direct
// DFT from a twiddle table), 10 log10(max(., 1e-30)), the lowest F bins.  (logpowspec's "minus the maximum" is a
// constant that the global mean / std normalisation removes.)  One warp per frame, lane = bins j, j + 32, ...
__global__ void __launch_bounds__(kFeatWarps * 32)
feat_spectrogram_kernel(const short* __restrict__ wave, const long long* __restrict__ offsets,
                        const int* __restrict__ n_samples, const unsigned int* __restrict__ absmax,
                        const float* __restrict__ window, float* __restrict__ raw, int T_pad, int F, int hop, int win,
                        const FeatOpts opts) {
  extern __shared__ float sh[];              // [2*win] twiddles, then per warp [win] samples
  float* tc = sh;
  float* ts = sh + win;
  float* xs = sh + 2 * win + (threadIdx.x >> 5) * win;
  for (int m = threadIdx.x; m < win; m += kFeatWarps * 32) {
    float sn, cs;
    sincospif(-2.f * (float)m / (float)win, &sn, &cs);
    tc[m] = cs;
    ts[m] = sn;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, b = blockIdx.y;
  const int frame = blockIdx.x * kFeatWarps + (threadIdx.x >> 5);
  const int n = n_samples[b];
  const int n_frames = frames_of(opts, n, hop, win);
  if (frame >= n_frames) return;
  const int len0 = (n <= win) ? 1 : 1 + (n - win + hop - 1) / hop;
  const int n_padded = n + (n_frames - len0) * hop;
  const short* w = wave + offsets[b];
  const float gain = 1.f / ((float)absmax[b] + 1e-5f);
  for (int i = lane; i < win; i += 32)   // no pre-emphasis for the spectrogram (sigproc.framesig on the raw signal)
    xs[i] = sample_at_psf(w, n, n_padded, frame * hop + i, gain, 0.f) * window[i];
  __syncwarp();
  for (int j = lane; j < F; j += 32) {
    float ar = 0.f, ai = 0.f;
    int idx = 0;
    for (int i = 0; i < win; ++i) {
      const float x = xs[i];
      ar += x * tc[idx];
      ai += x * ts[idx];
      idx += j;
      if (idx >= win) idx -= win;
    }
    float p = (ar * ar + ai * ai) / (float)win;
    p = fmaxf(p, 1e-30f);
    raw[((size_t)b * T_pad + frame) * F + j] = 10.f * log10f(p);
  }
}

// per (b, f): mean and population std over the utterance's frames, then normalise + pad.
__global__ void feat_norm_kernel(const float* __restrict__ raw, const int* __restrict__ n_samples,
                                 uint16_t* __restrict__ out_bf16, float* __restrict__ out_f32,
                                 int* __restrict__ out_lens, int T_pad, int F, int hop, int win, const FeatOpts opts) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  const int n_frames = min(frames_of(opts, n_samples[b], hop, win), T_pad);
  const float* src = raw + (size_t)b * T_pad * F + f;
  float s = 0.f;
  for (int t = lane; t < n_frames; t += 32) s += src[(size_t)t * F];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = opts.fixed_mean ? opts.fixed_mean[f] : s / (float)n_frames;
  float q = 0.f;
  for (int t = lane; t < n_frames; t += 32) {
    const float d = src[(size_t)t * F] - mean;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv_std = opts.fixed_std ? 1.f / opts.fixed_std[f] : rsqrtf(q / (float)n_frames);
  for (int t = lane; t < T_pad; t += 32) {
    float v = (t < n_frames) ? (src[(size_t)t * F] - mean) * inv_std : 0.f;
    if (t < n_frames && feat_masked(opts, b, t, f)) v = 0.f;
    const size_t o = ((size_t)b * T_pad + t) * F + f;
    if (out_bf16) store_act16(out_bf16, o, v, opts.out_f16);
    if (out_f32) out_f32[o] = v;
  }
  if (f == 0 && lane == 0 && out_lens) out_lens[b] = n_frames;
}

// one mean / population std per utterance over all frames x features (psf backend, or
// norm_per_feature = False): one CTA per utterance, three passes over its L2-resident [frames, F] slab.
__global__ void __launch_bounds__(256)
feat_norm_global_kernel(const float* __restrict__ raw, const int* __restrict__ n_samples,
                        uint16_t* __restrict__ out_bf16, float* __restrict__ out_f32,
                        int* __restrict__ out_lens, int T_pad, int F, int hop, int win, const FeatOpts opts) {
  __shared__ float red[8];
  __shared__ float bc;
  const int b = blockIdx.x;
  const int n_frames = min(frames_of(opts, n_samples[b], hop, win), T_pad);
  const int n = n_frames * F;
  const float* src = raw + (size_t)b * T_pad * F;
  auto block_sum = [&](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      bc = s;
    }
    __syncthreads();
    return bc;
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += src[i];
  const float mean = block_sum(s) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = src[i] - mean;
    q += d * d;
  }
  const float inv_std = rsqrtf(block_sum(q) / (float)n);
  for (int i = threadIdx.x; i < T_pad * F; i += 256) {
    float v = (i < n) ? (src[i] - mean) * inv_std : 0.f;
    if (i < n && feat_masked(opts, b, i / F, i % F)) v = 0.f;
    const size_t o = (size_t)b * T_pad * F + i;
    if (out_bf16) store_act16(out_bf16, o, v, opts.out_f16);
    if (out_f32) out_f32[o] = v;
  }
  if (threadIdx.x == 0 && out_lens) out_lens[b] = n_frames;
}

int logmel_forward(const short* wave, const long long* offsets, const int* n_samples, int B,
                   const float* mel, const int* mel_band, const float* window, int n_fft, int win, int hop, int F, int T_pad,
                   int max_samples, float dither, unsigned long long seed, float preemph,
                   unsigned int* absmax_ws, float* raw_ws, void* out_bf16, float* out_f32, int* out_lens,
                   cudaStream_t st, int psf_backend, int pad_to, int norm_per_feature, int out_f16,
                   const FeatExtras* ex) {
  if (n_fft != 512) return fail(ERR_UNSUPPORTED, "logmel_forward: only n_fft = 512 is built");
  if (win > n_fft || F > 256) return fail(ERR_INVALID, "logmel_forward: bad window / feature count");
  FeatOpts opts;
  opts.psf = psf_backend ? 1 : 0;
  opts.pad_to = pad_to;
  opts.per_feature = norm_per_feature ? 1 : 0;
  opts.power_scale = psf_backend ? 1.f / (float)n_fft : 1.f;
  opts.out_f16 = out_f16 ? 1 : 0;
  opts.fixed_gain = ex ? ex->fixed_gain : 0.f;
  opts.sig = ex ? ex->sig : nullptr;
  opts.sig_off = ex ? ex->sig_off : nullptr;
  opts.fixed_mean = ex ? ex->fixed_mean : nullptr;
  opts.fixed_std = ex ? ex->fixed_std : nullptr;
  opts.masks = ex ? ex->masks : nullptr;
  opts.n_masks = ex ? ex->n_masks : 0;
  opts.feature_type = ex ? ex->feature_type : 0;
  opts.post = ex ? ex->post : nullptr;
  opts.n_filt = ex ? ex->n_filt : 0;
  if (opts.feature_type != 0) {
    if (!psf_backend) return fail(ERR_UNSUPPORTED, "logmel_forward: spectrogram / mfcc are built for the psf backend");
    if (opts.feature_type == 2 && (!opts.post || opts.n_filt < F || opts.n_filt > kNfftMax))
      return fail(ERR_INVALID, "logmel_forward: mfcc needs the [F][n_filt] DCT matrix");
    if (opts.feature_type == 1 && F > win / 2 + 1)
      return fail(ERR_INVALID, "logmel_forward: num_features for spectrogram should be <= window / 2 + 1");
    if (opts.feature_type < 0 || opts.feature_type > 2) return fail(ERR_INVALID, "logmel_forward: unknown feature type");
  }
  if (opts.sig && psf_backend) return fail(ERR_UNSUPPORTED, "logmel_forward: augmented signals need the librosa backend");
  if ((opts.fixed_mean || opts.fixed_std) && !norm_per_feature)
    return fail(ERR_UNSUPPORTED, "logmel_forward: features_mean / features_std_dev need norm_per_feature");
  if (opts.sig && !opts.sig_off) return fail(ERR_INVALID, "logmel_forward: sig needs sig_off");
  if (!opts.sig && !(opts.fixed_gain > 0.f)) {
    // (with an augmented signal the caller has run os2s_wave_absmax + os2s_augment_signal already)
    OS2S_CUDA(cudaMemsetAsync(absmax_ws, 0, (size_t)B * sizeof(unsigned int), st));
    feat_absmax_kernel<<<dim3(32, B), 256, 0, st>>>(wave, offsets, n_samples, absmax_ws);
  }
  int max_frames = 1 + max_samples / hop;
  if (psf_backend) {
    max_frames = max_samples <= win ? 1 : 1 + (max_samples - win + hop - 1) / hop;
    if (pad_to > 0 && max_frames % pad_to) max_frames += pad_to - max_frames % pad_to;
  }
  if (max_frames > T_pad) return fail(ERR_INVALID, "logmel_forward: T_pad smaller than the frame count");
  dim3 grid((max_frames + kFeatWarps - 1) / kFeatWarps, B);
  if (opts.feature_type == 1) {
    const size_t smem = (size_t)(2 + kFeatWarps) * win * sizeof(float);
    feat_spectrogram_kernel<<<grid, kFeatWarps * 32, smem, st>>>(wave, offsets, n_samples, absmax_ws, window, raw_ws, T_pad, F,
                                                               hop, win, opts);
  } else {
    feat_logmel_kernel<512><<<grid, kFeatWarps * 32, 0, st>>>(wave, offsets, n_samples, absmax_ws, mel, mel_band, window,
                                                             raw_ws, T_pad, F, hop, win, dither, seed, preemph, opts);
  }
  if (norm_per_feature) {
    dim3 grid2((F + 7) / 8, B);
    feat_norm_kernel<<<grid2, 256, 0, st>>>(raw_ws, n_samples, (uint16_t*)out_bf16, out_f32, out_lens, T_pad, F,
                                            hop, win, opts);
  } else {
    feat_norm_global_kernel<<<B, 256, 0, st>>>(raw_ws, n_samples, (uint16_t*)out_bf16, out_f32, out_lens, T_pad,
                                               F, hop, win, opts);
  }
  return check_launch("logmel_forward");
}

int wave_absmax(const short* wave, const long long* offsets, const int* n_samples, int B, unsigned int* absmax,
                cudaStream_t st) {
  OS2S_CUDA(cudaMemsetAsync(absmax, 0, (size_t)B * sizeof(unsigned int), st));
  feat_absmax_kernel<<<dim3(32, B), 256, 0, st>>>(wave, offsets, n_samples, absmax);
  return check_launch("wave_absmax");
}

}  // namespace os2s
