// Audio-signal augmentation of the Speech2TextDataLayer on the GPU.
//
// Reference being replaced (NumPy + resampy on the py_func featurizer threads):
//   open_seq2seq/data/speech2text/speech_utils.py:216-222   normalize_signal (gain = 1 / (max|x| + 1e-5))
//   open_seq2seq/data/speech2text/speech_utils.py:245-259   speed perturbation: resampy.resample(signal, sr,
//                                                           int(sr * stretch), filter='kaiser_best')
//   open_seq2seq/data/speech2text/speech_utils.py:262-266   additive Gaussian noise at a drawn level (dB)
// resampy is band-limited sinc interpolation (J. O. Smith): every output sample is a dot product of the
// input around t / ratio with a Kaiser-windowed sinc that is tabulated 512 times per zero crossing and
// linearly interpolated between table entries.  One thread per output sample; the int16 waveform is read
// directly (normalised on the fly), the 128 KB table stays in L1 / L2.  The random draws (stretch factor,
// noise level) are made on the host in the reference's order and arrive as per-utterance arrays.
#include "common.h"
#include "kernels.h"

namespace os2s {

__device__ __forceinline__ float aug_gauss(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u1 = ((float)((z >> 40) & 0xFFFFFF) + 1.f) * (1.f / 16777217.f);
  const float u2 = (float)((z >> 8) & 0xFFFFFF) * (1.f / 16777216.f);
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.28318530718f * u2);
}

// one output sample of resampy's interpolation (general ratio): see augment_signal_kernel
__device__ __forceinline__ float resample_one(const short* __restrict__ x, int ni, int t, double tinc, double scale,
                                              int index_step, const float* __restrict__ win, int nwin, int num_table) {
  float acc;
  {
    const double treg = (double)t * tinc;
      const int n = (int)treg;
      acc = 0.f;
      // left wing: x[n], x[n-1], ...
      double frac = scale * (treg - (double)n);
      {
        const double index_frac = frac * (double)num_table;
        const int offset = (int)index_frac;
        const float eta = (float)(index_frac - (double)offset);
        const int i_max = min(n + 1, (nwin - offset) / index_step);
        for (int i = 0; i < i_max; ++i) {
          const int idx = offset + i * index_step;
          const float w0 = __ldg(win + idx);
          const float w1 = (idx + 1 < nwin) ? __ldg(win + idx + 1) : w0;
          acc += (w0 + eta * (w1 - w0)) * (float)x[n - i];
        }
      }
      // right wing: x[n+1], x[n+2], ...
      frac = scale - frac;
      {
        const double index_frac = frac * (double)num_table;
        const int offset = (int)index_frac;
        const float eta = (float)(index_frac - (double)offset);
        const int k_max = min(ni - n - 1, (nwin - offset) / index_step);
        for (int k = 0; k < k_max; ++k) {
          const int idx = offset + k * index_step;
          const float w0 = __ldg(win + idx);
          const float w1 = (idx + 1 < nwin) ? __ldg(win + idx + 1) : w0;
          acc += (w0 + eta * (w1 - w0)) * (float)x[n + k + 1];
        }
      }
  }
  return acc;
}

// out[b][t] = resample(wave[b] * gain_b)[t] + noise_amp[b] * N(0,1)
//   sr_new[b] == 0: no resampling (out = wave * gain + noise, n_out = n_in)
//   win: right half of the interpolation filter, nwin = num_zeros * num_table + 1 entries
// Fast path for rational ratios with a small numerator (the recipe's 0.9 / 1.0 / 1.1 = 9/10, 1/1, 11/10): output
// sample t sits at input time t*q/p, so the fractional position -- and with it the whole interpolated filter --
// repeats every p outputs.  The p x (left + right) tap weights are evaluated once per CTA into shared memory with
// exactly the arithmetic of the general kernel (same table entries, same linear interpolation), and every output is
// then a plain dot product: 2 loads per tap instead of 3 loads + index arithmetic.
constexpr int kAugMaxPhases = 16;
constexpr int kAugMaxTaps = 160;       // per wing: ceil(64 / 0.9) + 1 fits
__global__ void __launch_bounds__(256)
augment_signal_phase_kernel(const short* __restrict__ wave, const long long* __restrict__ offsets,
                            const int* __restrict__ n_in, const unsigned int* __restrict__ absmax, float fixed_gain,
                            const int* __restrict__ sr_new, int sr_orig, const float* __restrict__ win, int nwin,
                            int num_table, const float* __restrict__ noise_amp, unsigned long long seed,
                            float* __restrict__ out, const long long* __restrict__ out_offsets,
                            const int* __restrict__ n_out) {
  // odd row stride: the threads of a warp read the SAME tap of DIFFERENT phases (160 floats apart would put all
  // of them on one bank -- a p-way conflict on every load, measured 0.97 ms per batch in r02_final_kernel_evidence)
  __shared__ float wl[kAugMaxPhases][kAugMaxTaps + 1];
  __shared__ float wr[kAugMaxPhases][kAugMaxTaps + 1];
  __shared__ int nl[kAugMaxPhases], nr[kAugMaxPhases];
  const int b = blockIdx.y;
  const int no = n_out[b];
  const int ni = n_in[b];
  const short* x = wave + offsets[b];
  float* y = out + out_offsets[b];
  const float gain = fixed_gain > 0.f ? fixed_gain : 1.f / ((float)absmax[b] + 1e-5f);
  const float namp = noise_amp ? noise_amp[b] : 0.f;
  const unsigned long long useed = seed + (unsigned long long)b * 0xD1B54A32D192ED03ull;
  const int srn = sr_new ? sr_new[b] : 0;
  // reduced ratio p / q
  int pp = srn > 0 ? srn : sr_orig, qq = sr_orig;
  {
    int a = pp, c = qq;
    while (c) { const int t = a % c; a = c; c = t; }
    pp /= a;
    qq /= a;
  }
  const double ratio = (double)pp / (double)qq;
  const double scale = ratio < 1.0 ? ratio : 1.0;
  const int index_step = (int)(scale * (double)num_table);
  const float wscale = (ratio < 1.0 ? (float)ratio : 1.f) * gain;
  const bool phased = pp <= kAugMaxPhases;   // uniform over the CTA; larger numerators take the general path
  if (srn > 0 && phased) {
    // phase ph = (t * q) mod p  <=>  treg - floor(treg) = ph / p
    for (int e = threadIdx.x; e < pp * 2 * kAugMaxTaps; e += blockDim.x) {
      const int ph = e / (2 * kAugMaxTaps), r = e - ph * 2 * kAugMaxTaps;
      const bool right = r >= kAugMaxTaps;
      const int i = right ? r - kAugMaxTaps : r;
      double frac = scale * ((double)ph / (double)pp);
      if (right) frac = scale - frac;
      const double index_frac = frac * (double)num_table;
      const int offset = (int)index_frac;
      const float eta = (float)(index_frac - (double)offset);
      const int cnt = (nwin - offset) / index_step;
      float w = 0.f;
      if (i < cnt) {
        const int idx = offset + i * index_step;
        const float w0 = __ldg(win + idx);
        const float w1 = (idx + 1 < nwin) ? __ldg(win + idx + 1) : w0;
        w = w0 + eta * (w1 - w0);
      }
      (right ? wr : wl)[ph][i] = w;
      if (i == 0) (right ? nr : nl)[ph] = cnt < kAugMaxTaps ? cnt : kAugMaxTaps;
    }
  }
  __syncthreads();
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < no; t += gridDim.x * blockDim.x) {
    float acc;
    if (srn <= 0) {
      acc = (float)x[t] * gain;
    } else if (!phased) {
      acc = resample_one(x, ni, t, 1.0 / ratio, scale, index_step, win, nwin, num_table) * wscale;
    } else {
      const long long tq = (long long)t * qq;
      const int n = (int)(tq / pp);
      const int ph = (int)(tq - (long long)n * pp);
      acc = 0.f;
      const int i_max = min(n + 1, nl[ph]);
      const float* a = wl[ph];
      for (int i = 0; i < i_max; ++i) acc += a[i] * (float)x[n - i];
      const int k_max = min(ni - n - 1, nr[ph]);
      const float* c = wr[ph];
      for (int k = 0; k < k_max; ++k) acc += c[k] * (float)x[n + k + 1];
      acc *= wscale;
    }
    if (namp > 0.f) acc += namp * aug_gauss(useed, (unsigned long long)t);
    y[t] = acc;
  }
}

int augment_signal(const short* wave, const long long* offsets, const int* n_in, int B, const unsigned int* absmax,
                   float fixed_gain, const int* sr_new, int sr_orig, const float* win, int nwin, int num_table,
                   const float* noise_amp, unsigned long long seed, float* out, const long long* out_offsets,
                   const int* n_out, int max_out, cudaStream_t st) {
  if (B <= 0 || max_out <= 0) return fail(ERR_INVALID, "augment_signal: bad shape");
  if (sr_new && (!win || nwin < 2 || num_table < 1)) return fail(ERR_INVALID, "augment_signal: resampling needs the filter table");
  int bx = (max_out + 255) / 256;
  if (bx > 1024) bx = 1024;
  // (utterances whose ratio does not reduce to <= 16 phases take the general per-sample path inside the kernel)
  if (bx > 64) bx = 64;       // the per-CTA weight table is amortised over more outputs
  augment_signal_phase_kernel<<<dim3(bx, B), 256, 0, st>>>(wave, offsets, n_in, absmax, fixed_gain, sr_new, sr_orig, win,
                                                          nwin, num_table, noise_amp, seed, out, out_offsets, n_out);
  return check_launch("augment_signal");
}

}  // namespace os2s
