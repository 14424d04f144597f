"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the reference's audio / feature augmentation:

  open_seq2seq/data/speech2text/speech_utils.py:225-268   augment_audio_signal: speed perturbation
                                                          (resampy.resample, filter='kaiser_best') + noise
  open_seq2seq/data/speech2text/speech_utils.py:419-433   spec-augment frequency / time masks (zeros written
                                                          into the NORMALISED features)

The resampler lives in a third-party dependency that is not vendored under /root/reference and not
installed here: `resampy` (requirements.txt:3, unpinned; 0.2.x at the reference's date).  Its published
algorithm (resampy/core.py `resample`, resampy/interpn.py `resample_f`, resampy/filters.py `sinc_window`) is
band-limited sinc interpolation after J. O. Smith ("Digital Audio Resampling"):

  filter  'kaiser_best' = sinc_window(num_zeros=64, precision=9, window=kaiser(beta=14.769656459379492),
          rolloff=0.9475937167399596): the right half of a Kaiser-windowed sinc sampled 2**9 times per zero
          crossing, interp_win[i] = rolloff * sinc(rolloff * i / 512) * kaiser(2n+1, beta)[n + i];
  resample(x, sr_orig, sr_new): ratio = sr_new / sr_orig, n_out = int(len(x) * ratio); the table is scaled
          by ratio when ratio < 1; y[t] = sum over the left and the right wing of table values linearly
          interpolated between table samples (interp_delta), see `resample` below.

Parity status: the RESAMPLER is unpinned by the reference except for the length bounds of
speech_utils_test.py:20-43 (tests/test_oracle.py::test_augmentation_length_bounds_pin); its filter response is
additionally checked against scipy.signal.resample_poly on a band-limited signal (same test file).  The DRAWS
and the control flow around it (which value is drawn when, with which np.random call, what is added to what) are
pinned by executing the reference's own augment_audio_signal / get_speech_features_librosa on seeded streams
(tests/test_reference_executed_cpu.py; fixture tests/golden/reference_augmentation_draws.json)."""
import numpy as np


KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)


def sinc_window(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492):
    """resampy.filters.sinc_window with a Kaiser taper -> (interp_win [num_zeros * 2**precision + 1], 2**precision)."""
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


_TABLE = {}


def kaiser_best_table():
    if "kb" not in _TABLE:
        _TABLE["kb"] = sinc_window(**KAISER_BEST)
    return _TABLE["kb"]


def resample_out_len(n, sr_orig, sr_new):
    return int(n * (float(sr_new) / sr_orig))


def resample(x, sr_orig, sr_new):
    """resampy.resample(x, sr_orig, sr_new, filter='kaiser_best') for a 1-D signal (vectorised over the
    output samples; same arithmetic as resampy.interpn.resample_f)."""
    x = np.asarray(x, dtype=np.float64)
    ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[0] * ratio)
    win, num_table = kaiser_best_table()
    win = win.copy()
    if ratio < 1:
        win *= ratio
    delta = np.zeros_like(win)
    delta[:-1] = np.diff(win)
    scale = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale * num_table)
    nwin = win.shape[0]
    n_orig = x.shape[0]
    # resampy accumulates time_register by repeated addition; reproduce that (not t * increment)
    treg = np.zeros(n_out)
    if n_out > 1:
        treg[1:] = np.cumsum(np.full(n_out - 1, time_increment))
    y = np.zeros(n_out)
    n = treg.astype(np.int64)
    frac = scale * (treg - n)
    # left wing
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    i_max = np.minimum(n + 1, (nwin - offset) // index_step)
    for i in range(int(i_max.max()) if n_out else 0):
        ok = i < i_max
        idx = np.where(ok, offset + i * index_step, 0)
        w = win[idx] + eta * delta[idx]
        xi = x[np.where(ok, n - i, 0)]
        y += np.where(ok, w * xi, 0.0)
    # right wing
    frac = scale - frac
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    k_max = np.minimum(n_orig - n - 1, (nwin - offset) // index_step)
    for k in range(int(max(k_max.max(), 0)) if n_out else 0):
        ok = k < k_max
        idx = np.where(ok, offset + k * index_step, 0)
        w = win[idx] + eta * delta[idx]
        xi = x[np.where(ok, n + k + 1, 0)]
        y += np.where(ok, w * xi, 0.0)
    return y


def draw_augmentation(n_samples, sample_freq, augmentation, rng):
    """The random draws of augment_audio_signal in the reference's order (np.random.choice / rand for the
    speed, np.random.randint for the noise level) -> (sr_new or 0, noise amplitude or 0.0)."""
    sr_new, noise_amp = 0, 0.0
    if "speed_perturbation_ratio" in augmentation:
        r = augmentation["speed_perturbation_ratio"]
        stretch = -1
        if isinstance(r, list):
            stretch = rng.choice(r)
        elif r > 0:
            stretch = 1.0 + (2.0 * rng.rand() - 1.0) * r
        if stretch > 0:
            sr_new = int(sample_freq * stretch)
    if "noise_level_min" in augmentation and "noise_level_max" in augmentation:
        db = rng.randint(low=augmentation["noise_level_min"], high=augmentation["noise_level_max"])
        noise_amp = 10.0 ** (db / 20.0)
    return sr_new, noise_amp


def augment_audio_signal(signal_float, sample_freq, augmentation, rng=None):
    """speech_utils.py:225-268 (rng: a np.random.RandomState; the reference uses the global np.random)."""
    rng = rng or np.random
    sr_new, noise_amp = draw_augmentation(len(signal_float), sample_freq, augmentation, rng)
    out = np.asarray(signal_float, dtype=np.float64)
    if sr_new > 0:
        out = resample(out, sample_freq, sr_new)
    if noise_amp > 0:
        out = out + rng.randn(out.shape[0]) * noise_amp
    return out


def draw_spec_masks(n_frames, n_features, augmentation, rng):
    """speech_utils.py:419-433: list of (kind, base, width), kind 0 = frequency band, 1 = time band, in the
    reference's draw order.  A time mask that does not fit (frames - width <= 0) is dropped as there."""
    masks = []
    for _ in range(augmentation.get("n_freq_mask", 0)):
        band = rng.randint(augmentation.get("width_freq_mask", 10) + 1)
        base = rng.randint(0, n_features - band)
        masks.append((0, int(base), int(band)))
    for _ in range(augmentation.get("n_time_mask", 0)):
        band = rng.randint(augmentation.get("width_time_mask", 50) + 1)
        if n_frames - band > 0:
            base = rng.randint(n_frames - band)
            masks.append((1, int(base), int(band)))
    return masks


def apply_spec_masks(features, masks):
    """features [frames, F] (normalised) -> copy with the masked bands zeroed."""
    out = np.array(features, copy=True)
    for kind, base, width in masks:
        if kind == 0:
            out[:, base:base + width] = 0
        else:
            out[base:base + width, :] = 0
    return out
