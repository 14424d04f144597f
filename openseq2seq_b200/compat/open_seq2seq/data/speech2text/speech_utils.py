"""Host-side helpers of the speech data layer (open_seq2seq/data/speech2text/speech_utils.py).

The feature arithmetic of the reference (NumPy + librosa on py_func threads) runs on the GPU in
os2s_logmel_forward; what stays on the host is only what has to: reading wav files and building the
constant mel filterbank / window tables once.  The Slaney mel filterbank follows librosa 0.6.3's
published definition (htk=False, area normalisation), see INTEGRATION.md."""
import math

import numpy as np


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax): [n_mels, 1+n_fft//2] float64."""
    fmax = sr / 2.0 if fmax is None else fmax
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz2mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def mel2hz(m):
        return min_log_hz * math.exp(logstep * (m - min_log_mel)) if m >= min_log_mel else f_sp * m

    pts = np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2)
    edges = np.array([mel2hz(m) for m in pts])
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    w = np.zeros((n_mels, freqs.size))
    for i in range(n_mels):
        up = (freqs - edges[i]) / (edges[i + 1] - edges[i])
        down = (edges[i + 2] - freqs) / (edges[i + 2] - edges[i + 1])
        w[i] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (edges[i + 2] - edges[i]))
    return w


def psf_filterbank(nfilt, nfft, samplerate, lowfreq=0.0, highfreq=None):
    """python_speech_features.base.get_filterbanks (the psf backend's logfbank, speech_utils.py:514-522):
    HTK mel scale, nfilt + 2 points evenly spaced in mel, bins floor((nfft+1)*hz/sr), plain triangles."""
    highfreq = highfreq or samplerate / 2.0
    hz2mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    mel2hz = lambda mel: 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    bins = np.floor((nfft + 1) * mel2hz(np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)) / samplerate)
    fb = np.zeros((nfilt, nfft // 2 + 1))
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fb


def psf_num_frames(n, n_win, n_hop, pad_to):
    """frames of the psf backend for an n-sample utterance (speech_utils.py:478-488)."""
    length = 1 if n <= n_win else 1 + -(-(n - n_win) // n_hop)
    if pad_to > 0 and length % pad_to:
        length += pad_to - length % pad_to
    return length


def num_fft_for(window_size, sample_freq):
    return 2 ** math.ceil(math.log2(window_size * sample_freq))


def read_wav(filename, expected_sr):
    """scipy.io.wavfile read + sample-rate check (speech_utils.py:188-195) -> int16 mono array."""
    import scipy.io.wavfile as wave
    sr, signal = wave.read(filename)
    if sr != expected_sr:
        raise ValueError("The sampling frequency set in params {} does not match the frequency {} read from "
                         "file {}".format(expected_sr, sr, filename))
    if signal.ndim > 1:
        signal = signal[:, 0]
    if signal.dtype != np.int16:
        signal = (np.clip(signal, -1.0, 1.0) * 32767).astype(np.int16) if signal.dtype.kind == "f" \
            else signal.astype(np.int16)
    return signal


def kaiser_best_table():
    """Interpolation table of resampy's 'kaiser_best' filter (the speed perturbation of
    augment_audio_signal, speech_utils.py:245-259): the right half of a Kaiser-windowed sinc with 64 zero
    crossings sampled 2**9 times per crossing, rolloff 0.9475937167399596, beta 14.769656459379492
    (resampy.filters.sinc_window).  Returns (float64 [64 * 512 + 1], 512); os2s_augment_signal interpolates it."""
    num_zeros, precision, rolloff, beta = 64, 9, 0.9475937167399596, 14.769656459379492
    num_table = 2 ** precision
    n = num_table * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_table


def psf_mfcc_matrix(numcep, nfilt, ceplifter):
    """python_speech_features.base.mfcc after the log filterbank (get_speech_features_psf, speech_utils.py:504-512):
    scipy.fftpack.dct(type=2, norm='ortho') over the nfilt log energies, the first numcep coefficients, the
    sinusoidal lifter 1 + (L/2) sin(pi n / L) -> one [numcep, nfilt] matrix (appendEnergy=False)."""
    n = np.arange(nfilt)
    dct = np.cos(np.pi * (n[None, :] + 0.5) * np.arange(nfilt)[:, None] / nfilt)
    scale = np.full(nfilt, math.sqrt(2.0 / nfilt))
    scale[0] = math.sqrt(1.0 / nfilt)
    m = (dct * scale[:, None])[:numcep]
    lift = 1.0 + (ceplifter / 2.0) * np.sin(np.pi * np.arange(numcep) / ceplifter) if ceplifter > 0 else np.ones(numcep)
    return m * lift[:, None]
