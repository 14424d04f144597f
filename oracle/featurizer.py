"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement, in NumPy, of the reference's librosa-backend log-mel featurizer used by Jasper:

  open_seq2seq/data/speech2text/speech_utils.py:216-222  normalize_signal
  open_seq2seq/data/speech2text/speech_utils.py:271-272  preemphasis
  open_seq2seq/data/speech2text/speech_utils.py:322-441  get_speech_features_librosa (logfbank branch :396-406,
                                                         per-feature normalisation :411-417)
  open_seq2seq/data/speech2text/speech2text.py:167-183   precomputed mel basis
  open_seq2seq/data/speech2text/speech2text.py:251-257,313-317  zero padding of the batch / pad_to

The arithmetic itself lives in librosa==0.6.3 (pinned in the reference's requirements.txt:8), which
is NOT under /root/reference and is not installed here.  Its published algorithm is restated:
  librosa.core.stft(y, n_fft, hop, win_length, center=True, window=np.hanning):
      window = np.hanning(win_length) (symmetric), zero-padded centrally to n_fft; y reflect-padded
      by n_fft//2 on both sides; frames = 1 + len(y)//hop; complex64 output.
  librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax): Slaney mel scale (htk=False), triangular
      filters on np.linspace(0, sr/2, 1+n_fft//2), Slaney area normalisation 2/(f[i+2]-f[i]).
Parity status: the reference has no golden vectors for this backend ("parity unpinned" by its own
tests).  The librosa / psf PRIMITIVES restated here are triangulated against torchaudio's
independent Slaney filterbank and torch.stft (tests/test_oracle.py); the COMPOSITION of this module (order of
the steps and of the random draws, constants, padding, normalisation) is pinned by executing the reference's own
speech_utils.py on top of these primitives (tests/test_reference_executed_cpu.py, build container only).
"""
import math

import numpy as np


def normalize_signal(signal, gain=None):
    """speech_utils.py:216-222."""
    if gain is None:
        gain = 1.0 / (np.max(np.abs(signal)) + 1e-5)
    return signal * gain


def preemphasis(signal, coeff=0.97):
    """speech_utils.py:271-272."""
    return np.append(signal[0], signal[1:] - coeff * signal[:-1])


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_t = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_t, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr=16000, n_fft=512, n_mels=64, fmin=0.0, fmax=None):
    """librosa.filters.mel (0.6.3) restated: [n_mels, 1 + n_fft//2] float64."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights


def stft_power(signal, n_fft=512, hop=160, win_length=320):
    """|librosa.core.stft(...)|**2 -> [1 + n_fft//2, frames] (float32, from complex64)."""
    window = np.hanning(win_length)
    lpad = (n_fft - win_length) // 2
    fft_window = np.zeros(n_fft)
    fft_window[lpad:lpad + win_length] = window
    y = np.pad(np.asarray(signal), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = y[idx] * fft_window[None, :]
    spec = np.fft.rfft(frames, n=n_fft, axis=1).astype(np.complex64)  # librosa dtype
    return (np.abs(spec) ** 2.0).T


def num_fft_for(window_size=20e-3, sample_freq=16000):
    """speech_utils.py:358 / speech2text.py:170-175."""
    return 2 ** math.ceil(math.log2(window_size * sample_freq))


def logfbank_features(signal_int16, sample_freq=16000, num_features=64, window_size=20e-3,
                      window_stride=10e-3, dither=0.0, norm_per_feature=True, mel_basis=None,
                      rng=None):
    """get_speech_features_librosa(features_type='logfbank') without augmentation.

    Returns (features [frames, num_features] float64, duration seconds)."""
    signal = normalize_signal(np.asarray(signal_int16).astype(np.float32))
    duration = len(signal) * 1.0 / sample_freq
    n_win = int(sample_freq * window_size)
    n_hop = int(sample_freq * window_stride)
    n_fft = num_fft_for(window_size, sample_freq)
    if dither > 0:
        rng = rng or np.random
        signal = signal + dither * rng.randn(*signal.shape)
    signal = preemphasis(signal, coeff=0.97)
    S = stft_power(signal, n_fft=n_fft, hop=n_hop, win_length=n_win)
    if mel_basis is None:
        mel_basis = mel_filterbank(sample_freq, n_fft, n_mels=num_features, fmin=0,
                                   fmax=int(sample_freq / 2))
    features = np.log(np.dot(mel_basis, S) + 1e-20).T
    axis = 0 if norm_per_feature else None
    mean = np.mean(features, axis=axis)
    std = np.std(features, axis=axis)
    features = (features - mean) / std
    return features, duration


def batch_features(signals_int16, pad_to=16, **kw):
    """padded_batch + pad_to of speech2text.py:251-257,313-317: ([B,T,F] float64, lengths int32)."""
    feats = [logfbank_features(s, **kw)[0] for s in signals_int16]
    lens = np.array([f.shape[0] for f in feats], dtype=np.int32)
    T = int(lens.max())
    if pad_to > 0 and T % pad_to != 0:
        T += pad_to - T % pad_to
    out = np.zeros((len(feats), T, feats[0].shape[1]))
    for i, f in enumerate(feats):
        out[i, :f.shape[0]] = f
    return out, lens


# ---------------------------------------------------------------------------------------------
# psf backend (speech_utils.py:444-535, the default backend: used by the Wave2Letter(+) configs and
# the reference's toy tests).  The arithmetic lives in python_speech_features (unpinned in
# requirements.txt:9, not under /root/reference, not installed); its published algorithm (v0.6,
# base.py / sigproc.py) is restated:
#   sigproc.preemphasis(x, c)        = append(x[0], x[1:] - c*x[:-1])
#   sigproc.framesig(sig, L, S, win) : numframes = 1 + ceil((len - L)/S) (1 if len <= L), signal zero-padded
#                                      to (numframes-1)*S + L, frames * win(L)
#   sigproc.powspec(frames, NFFT)    = 1/NFFT * |rfft(frames, NFFT)|^2
#   base.get_filterbanks(nfilt, nfft, sr, lo, hi): HTK mel (2595*log10(1+f/700)), nfilt+2 points evenly
#       spaced in mel, bin = floor((nfft+1)*hz/sr), unnormalised triangles between consecutive bins
#   base.fbank(): preemphasis -> framesig (winfunc: rectangular by default, as the reference calls it) ->
#       powspec -> energies = pspec . fb^T, zeros replaced by float eps;  logfbank = log(fbank)
# Pins in the reference's own tests (speech_utils_test.py:45-85): output shape and global mean 0 / std 1.


def psf_mel_filterbank(nfilt=64, nfft=512, samplerate=16000, lowfreq=0.0, highfreq=None):
    """python_speech_features.base.get_filterbanks -> [nfilt, nfft//2 + 1] float64."""
    highfreq = highfreq or samplerate / 2.0
    hz2mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    mel2hz = lambda mel: 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fb = np.zeros((nfilt, nfft // 2 + 1))
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fb


def psf_quantize_and_pad(signal_int16, sample_freq=16000, window_size=20e-3, window_stride=10e-3, pad_to=8):
    """speech_utils.py:473-488: re-quantise the normalised signal to int16 and zero-pad it so that the
    number of frames is a multiple of pad_to."""
    signal = (normalize_signal(np.asarray(signal_int16).astype(np.float32)) * 32767.0).astype(np.int16)
    n_win, n_hop = int(sample_freq * window_size), int(sample_freq * window_stride)
    length = 1 + int(math.ceil((1.0 * signal.shape[0] - n_win) / n_hop))
    if pad_to > 0 and length % pad_to != 0:
        signal = np.pad(signal, (0, (pad_to - length % pad_to) * n_hop), mode="constant")
    return signal


def _psf_frames(signal, n_win, n_hop, winfunc):
    """python_speech_features.sigproc.framesig."""
    slen = len(signal)
    numframes = 1 if slen <= n_win else 1 + int(math.ceil((1.0 * slen - n_win) / n_hop))
    padded = np.concatenate([signal, np.zeros((numframes - 1) * n_hop + n_win - slen)])
    idx = np.arange(n_win)[None, :] + n_hop * np.arange(numframes)[:, None]
    return padded[idx] * winfunc(n_win)[None, :]


def psf_logpowspec(frames, nfft):
    """python_speech_features.sigproc.logpowspec(frames, NFFT, norm=1): 10 log10(max(|rfft|^2 / NFFT, 1e-30))
    minus its maximum."""
    ps = (1.0 / nfft) * np.abs(np.fft.rfft(frames, nfft, axis=1)) ** 2
    ps[ps <= 1e-30] = 1e-30
    lps = 10.0 * np.log10(ps)
    return lps - np.max(lps)


def psf_log_fbank(signal, samplerate, n_win, n_hop, nfilt, nfft, lowfreq, highfreq, preemph):
    """python_speech_features.base.logfbank = log(fbank(...)[0]): pre-emphasis, RECTANGULAR frames (the
    reference passes no winfunc), power spectrum / NFFT, HTK-mel triangles, zeros -> float eps, log."""
    frames = _psf_frames(preemphasis(signal, preemph), n_win, n_hop, np.ones)
    pspec = (1.0 / nfft) * np.abs(np.fft.rfft(frames, nfft, axis=1)) ** 2
    fb = psf_mel_filterbank(nfilt, nfft, samplerate, lowfreq, highfreq)
    feat = np.dot(pspec, fb.T)
    return np.log(np.where(feat == 0, np.finfo(float).eps, feat))


def psf_mfcc(signal, samplerate, n_win, n_hop, numcep, nfilt, nfft, lowfreq, highfreq, preemph, ceplifter):
    """python_speech_features.base.mfcc(appendEnergy=False): orthonormal DCT-II of the log filterbank energies,
    first numcep coefficients, sinusoidal lifter 1 + (L/2) sin(pi n / L)."""
    feat = psf_log_fbank(signal, samplerate, n_win, n_hop, nfilt, nfft, lowfreq, highfreq, preemph)
    n = np.arange(nfilt)
    dct = np.cos(np.pi * (n[None, :] + 0.5) * np.arange(nfilt)[:, None] / nfilt)      # DCT-II rows
    scale = np.full(nfilt, np.sqrt(2.0 / nfilt))
    scale[0] = np.sqrt(1.0 / nfilt)                                                   # norm='ortho'
    ceps = np.dot(feat, (dct * scale[:, None]).T)[:, :numcep]
    L = ceplifter
    return ceps * (1.0 + (L / 2.0) * np.sin(np.pi * np.arange(numcep) / L))[None, :]


def psf_logfbank_features(signal_int16, sample_freq=16000, num_features=64, window_size=20e-3,
                          window_stride=10e-3, pad_to=8, nfft=512, preemph=0.97):
    """get_speech_features_psf(features_type='logfbank') without augmentation -> ([frames, F] f64, seconds)."""
    duration = len(signal_int16) * 1.0 / sample_freq
    signal = psf_quantize_and_pad(signal_int16, sample_freq, window_size, window_stride, pad_to).astype(np.float64)
    n_win, n_hop = int(sample_freq * window_size), int(sample_freq * window_stride)
    features = psf_log_fbank(signal, sample_freq, n_win, n_hop, num_features, nfft, 0.0, sample_freq / 2.0, preemph)
    features = (features - np.mean(features)) / np.std(features)   # global normalisation (:531-533)
    return features, duration


def psf_spectrogram_features(signal_int16, sample_freq=16000, num_features=161, window_size=20e-3,
                             window_stride=10e-3, pad_to=8):
    """get_speech_features_psf(features_type='spectrogram') (speech_utils.py:490-502): Hann-windowed frames,
    sigproc.logpowspec(frames, NFFT = window length), the lowest num_features bins, global mean/std
    normalisation."""
    duration = len(signal_int16) * 1.0 / sample_freq
    signal = psf_quantize_and_pad(signal_int16, sample_freq, window_size, window_stride, pad_to).astype(np.float64)
    n_win, n_hop = int(sample_freq * window_size), int(sample_freq * window_stride)
    assert num_features <= n_win // 2 + 1, \
        "num_features for spectrogram should be <= (sample_freq * window_size // 2 + 1)"
    features = psf_logpowspec(_psf_frames(signal, n_win, n_hop, np.hanning), n_win)[:, :num_features]
    features = (features - np.mean(features)) / np.std(features)
    return features, duration


def psf_mfcc_features(signal_int16, sample_freq=16000, num_features=13, window_size=20e-3, window_stride=10e-3,
                      pad_to=8, nfft=512, preemph=0.97):
    """get_speech_features_psf(features_type='mfcc') (speech_utils.py:504-512): psf.mfcc(numcep = F,
    nfilt = 2F, nfft = 512, ceplifter = 2F, appendEnergy = False); global normalisation."""
    duration = len(signal_int16) * 1.0 / sample_freq
    signal = psf_quantize_and_pad(signal_int16, sample_freq, window_size, window_stride, pad_to).astype(np.float64)
    n_win, n_hop = int(sample_freq * window_size), int(sample_freq * window_stride)
    ceps = psf_mfcc(signal, sample_freq, n_win, n_hop, num_features, 2 * num_features, nfft, 0.0, sample_freq / 2.0,
                    preemph, 2 * num_features)
    features = (ceps - np.mean(ceps)) / np.std(ceps)
    return features, duration
