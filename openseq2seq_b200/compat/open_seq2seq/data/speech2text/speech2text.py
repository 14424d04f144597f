"""Speech2TextDataLayer (open_seq2seq/data/speech2text/speech2text.py:25-485) over the GPU featurizer.

Same constructor signature and params schema; CSV (wav_filename, transcript) datasets, vocab file
-> char2idx with `tgt_vocab_size = len(vocab) + 1` (blank last), duration filters, per-worker
seeding / eval sharding (`split_data`), padded batches with `pad_to`.

The reference builds a tf.data graph: shuffle -> repeat -> map(py_func featurizer, 8 threads) -> filter ->
padded_batch -> prefetch (speech2text.py:227-257).  Here the same stages are:
  * a pool of 8 reader threads decodes wav files (the reference's num_parallel_calls=8),
  * a producer thread collates a batch into a ring of PINNED host buffers, copies it to the device and
    runs augmentation + featurizer (os2s_wave_absmax / os2s_augment_signal / os2s_features_forward_p) on a
    SIDE stream, `prefetch` batches ahead of the training step,
  * `iterator` hands out finished batches; the consumer's stream waits on the batch's CUDA event, so the
    host never blocks on the device and the featurizer overlaps the previous training step.
Augmentation (speech_utils.py:225-268 speed perturbation + noise, :419-433 spec-augment masks), `gain`,
`features_mean` / `features_std_dev` are applied on the GPU; the random draws are made on the host with
NumPy in the reference's order.

Extension used by bench.py and tests: dataset_files may contain entries of the form
"synthetic:<n_utts>:<seconds>[:<seed>]" which generate band-limited noise utterances and random
transcripts in memory -- there is no network / dataset in the build image."""
import collections
import ctypes
import queue
import threading

import numpy as np
import torch

from open_seq2seq.data.data_layer import DataLayer
from open_seq2seq.data.utils import load_pre_existing_vocabulary
from . import speech_utils


class _PinnedRing(object):
    """A few reusable pinned host buffers (grown on demand): no cudaHostAlloc per batch.  A slot is handed out
    again only after the asynchronous copies that read it have finished (mark() records an event for them)."""

    def __init__(self, slots):
        self.bufs = [None] * slots
        self.events = [None] * slots
        self.k = 0
        self.last = None

    def get(self, n_bytes):
        i = self.k
        self.k = (self.k + 1) % len(self.bufs)
        if self.events[i] is not None:
            self.events[i].synchronize()
            self.events[i] = None
        b = self.bufs[i]
        if b is None or b.numel() < n_bytes:
            b = torch.empty(int(n_bytes * 1.25) + 4096, dtype=torch.uint8).pin_memory()
            self.bufs[i] = b
        self.last = i
        return b

    def mark(self):
        """Call after enqueueing the copies out of the buffer returned by the last get()."""
        if self.last is not None:
            ev = torch.cuda.Event()
            ev.record()
            self.events[self.last] = ev


class Speech2TextDataLayer(DataLayer):
    @staticmethod
    def get_required_params():
        return dict(DataLayer.get_required_params(), **{
            "num_audio_features": int,
            "input_type": ["spectrogram", "mfcc", "logfbank"],
            "vocab_file": str,
            "dataset_files": list,
        })

    @staticmethod
    def get_optional_params():
        return dict(DataLayer.get_optional_params(), **{
            "backend": ["psf", "librosa"], "augmentation": dict, "pad_to": int,
            "max_duration": float, "min_duration": float, "bpe": bool, "autoregressive": bool,
            "syn_enable": bool, "syn_subdirs": list, "window_size": float, "window_stride": float,
            "dither": float, "norm_per_feature": bool, "window": ["hanning", "hamming", "none"],
            "num_fft": int, "precompute_mel_basis": bool, "sample_freq": int, "gain": float,
            "features_mean": np.ndarray, "features_std_dev": np.ndarray,
        })

    def __init__(self, params, model, num_workers, worker_id):
        super(Speech2TextDataLayer, self).__init__(params, model, num_workers, worker_id)
        p = self.params
        if p.get("bpe", False) or p.get("autoregressive", False):
            raise NotImplementedError("Speech2TextDataLayer: bpe / autoregressive targets are not built")
        if p.get("syn_enable", False):
            raise NotImplementedError("Speech2TextDataLayer: syn_enable (synthetic-data subdirectories) is not built")
        if p.get("cache_features", False):
            raise NotImplementedError("Speech2TextDataLayer: cache_features is not built (features are computed "
                                      "on the GPU every step)")
        p["char2idx"] = load_pre_existing_vocabulary(p["vocab_file"], read_chars=True)
        p["idx2char"] = {i: w for w, i in p["char2idx"].items()}
        p["tgt_vocab_size"] = len(p["char2idx"]) + 1  # + blank (speech2text.py:120-125)
        self.target_pad_value = 0
        p["min_duration"] = p.get("min_duration", -1.0)
        p["max_duration"] = p.get("max_duration", -1.0)
        p["window_size"] = p.get("window_size", 20e-3)
        p["window_stride"] = p.get("window_stride", 10e-3)
        p["sample_freq"] = p.get("sample_freq", 16000)
        self._psf = p.get("backend", "psf") == "psf"
        aug = p.get("augmentation")
        # augmentation is a training-time transform of the librosa backend in every reference config
        # (get_speech_features_psf re-quantises to int16 after it, a path that is not built)
        if aug and "n_freq_mask" in aug and aug.get("width_freq_mask", 10) > p["num_audio_features"]:
            raise ValueError("'width_freq_mask'={} should be smaller than 'num_audio_features'={}".format(
                aug.get("width_freq_mask", 10), p["num_audio_features"]))              # speech2text.py:184-192
        if aug and "time_stretch_ratio" in aug:                                        # speech2text.py:195-197
            print("WARNING: Please update time_stretch_ratio to speed_perturbation_ratio")
            aug = dict(aug)
            aug["speed_perturbation_ratio"] = aug.pop("time_stretch_ratio")
            p["augmentation"] = aug
        self._aug = dict(aug) if (aug and p["mode"] == "train") else None
        if self._aug:
            known = {"speed_perturbation_ratio", "noise_level_min", "noise_level_max", "n_freq_mask", "n_time_mask",
                     "width_freq_mask", "width_time_mask"}
            unknown = set(self._aug) - known
            if unknown:
                raise ValueError("Speech2TextDataLayer: unknown augmentation keys %r" % sorted(unknown))
            if self._psf:
                raise NotImplementedError("Speech2TextDataLayer: augmentation is built for backend='librosa' only")
        if self._psf and (p.get("gain") is not None or p.get("features_mean") is not None
                          or p.get("features_std_dev") is not None):
            raise NotImplementedError("Speech2TextDataLayer: gain / features_mean / features_std_dev belong to the "
                                      "librosa backend (speech_utils.py:306-313)")
        self.feature_dtype = "bf16"       # set_feature_dtype(): the 16-bit format of the engine that consumes us
        self.prefetch = 2                 # batches featurised ahead of the consumer (tf.data prefetch)
        self.reader_threads = 8           # speech2text.py:241 num_parallel_calls=8
        self._files = None
        self._load_file_list()
        self._dev = None
        self._iterator = None
        self._input_tensors = None
        self._syn_cache = {}
        self._producer = None

    def set_feature_dtype(self, name):
        if name not in ("bf16", "fp16"):
            raise ValueError("feature dtype is 'bf16' or 'fp16'")
        self.feature_dtype = name
        self._ws = collections.OrderedDict()

    # ------------------------------------------------------------------ dataset
    def _load_file_list(self):
        p = self.params
        rows = []
        for f in p["dataset_files"]:
            if isinstance(f, str) and f.startswith("synthetic:"):
                parts = f.split(":")
                n, secs = int(parts[1]), float(parts[2])
                seed = int(parts[3]) if len(parts) > 3 else 1234
                for i in range(n):
                    rows.append(("synthetic:%d:%g:%d" % (i, secs, seed), None))
                continue
            import pandas as pd
            csv = pd.read_csv(f, encoding="utf-8")
            col = "wav_filename"
            for fn, tr in zip(csv[col].values, csv["transcript"].values):
                rows.append((fn, tr))
        if p["mode"] != "infer":
            self._files = rows
        else:
            self._files = [(fn, None) for fn, _ in rows]
        self._all_size = len(self._files)
        self.split_data()

    def split_data(self):
        """speech2text.py:200-210: eval/infer shard contiguously per worker; train sees everything."""
        if self.params["mode"] != "train" and self._num_workers is not None and self._num_workers > 1:
            size = len(self._files)
            start = size // self._num_workers * self._worker_id
            end = size if self._worker_id == self._num_workers - 1 else size // self._num_workers * (self._worker_id + 1)
            self._files = self._files[start:end]

    def get_size_in_samples(self):
        return len(self._files)

    # ---------------------------------------------------------------- featurizer
    def _setup_device_tables(self):
        p = self.params
        self._ftype = {"logfbank": 0, "spectrogram": 1, "mfcc": 2}[p["input_type"]]
        if self._ftype != 0 and not self._psf:
            raise NotImplementedError("Speech2TextDataLayer: input_type %r is built for the python_speech_features "
                                      "backend (the default); the librosa backend has 'logfbank'" % p["input_type"])
        # psf normalises with one mean / std per utterance; librosa per feature unless norm_per_feature=False
        self._per_feature = (not self._psf) and bool(p.get("norm_per_feature", False))
        sr = p["sample_freq"]
        self.n_win = int(sr * p["window_size"])
        self.n_hop = int(sr * p["window_stride"])
        F = p["num_audio_features"]
        self._post, self._n_filt = None, 0
        if self._psf:
            # psf.logfbank(nfft=512, winfunc = rectangular) (speech_utils.py:514-522); psf.mfcc(nfilt = 2 F, ceplifter =
            # 2 F) (:504-512); spectrogram: Hann frames, NFFT = window length (:490-502)
            self.n_fft = 512
            n_filt = 2 * F if self._ftype == 2 else F
            mel = speech_utils.psf_filterbank(n_filt, self.n_fft, sr, 0.0, sr / 2.0)
            win = np.hanning(self.n_win) if self._ftype == 1 else np.ones(self.n_win)
            if self._ftype == 1 and F > self.n_win // 2 + 1:
                raise AssertionError("num_features for spectrogram should be <= (sample_freq * window_size // 2 + 1)")
            if self._ftype == 2:
                self._n_filt = n_filt
                self._post_np = speech_utils.psf_mfcc_matrix(F, n_filt, 2 * F)
        else:
            self.n_fft = p.get("num_fft") or speech_utils.num_fft_for(p["window_size"], sr)
            mel = speech_utils.mel_filterbank(sr, self.n_fft, F, 0.0, int(sr / 2))
            win_name = p.get("window", "hanning")
            win = {"hanning": np.hanning, "hamming": np.hamming, "none": np.ones}[win_name](self.n_win)
        self._dev = torch.device("cuda")
        dev = self._dev
        self._mel = torch.tensor(mel, dtype=torch.float32, device=dev)
        band = [[int(np.nonzero(r)[0].min()), int(np.nonzero(r)[0].max()) + 1] if np.any(r) else [0, 0] for r in mel]
        self._band = torch.tensor(band, dtype=torch.int32, device=dev)
        self._win = torch.tensor(win, dtype=torch.float32, device=dev)
        if self._ftype == 2:
            self._post = torch.tensor(self._post_np, dtype=torch.float32, device=dev)
        self._fixed_mean = self._fixed_std = None
        if p.get("features_mean") is not None or p.get("features_std_dev") is not None:
            if not self._per_feature:
                raise NotImplementedError("Speech2TextDataLayer: features_mean / features_std_dev need norm_per_feature")
            if p.get("features_mean") is not None:
                self._fixed_mean = torch.tensor(np.asarray(p["features_mean"], dtype=np.float32).reshape(F), device=dev)
            if p.get("features_std_dev") is not None:
                self._fixed_std = torch.tensor(np.asarray(p["features_std_dev"], dtype=np.float32).reshape(F), device=dev)
        self._gain = float(p["gain"]) if p.get("gain") is not None else 0.0
        self._resample_tab = None
        if self._aug and "speed_perturbation_ratio" in self._aug:
            tab, self._num_table = speech_utils.kaiser_best_table()
            self._resample_tab = torch.tensor(tab, dtype=torch.float32, device=dev)
        self._n_masks = 0
        if self._aug:
            self._n_masks = int(self._aug.get("n_freq_mask", 0)) + int(self._aug.get("n_time_mask", 0))
        self._ws = collections.OrderedDict()
        self._pinned = _PinnedRing(2 * (self.prefetch + 2))
        self._dev_cap = {"wave": 0, "sig": 0}
        self._dev_buf = {}

    def _device_buffer(self, name, n, dtype):
        """Capacity-grown device scratch (waveforms / augmented signals): never keyed on the exact size."""
        if self._dev_cap[name] < n:
            self._dev_cap[name] = int(n * 1.25) + 1024
            self._dev_buf[name] = torch.empty(self._dev_cap[name], dtype=dtype, device=self._dev)
        return self._dev_buf[name]

    def _out_buffers(self, B, T):
        """Output / scratch buffers for one (B, T): a small ring (the consumer copies a batch into the engine's
        static input buffer before the ring comes round), LRU-bounded over shapes."""
        key = (B, T)
        ws = self._ws.get(key)
        F = self.params["num_audio_features"]
        if ws is None:
            dev = self._dev
            odt = torch.float16 if self.feature_dtype == "fp16" else torch.bfloat16
            ring = self.prefetch + 2
            ws = {"raw": torch.empty(B * T * F, dtype=torch.float32, device=dev),
                  "out": [torch.empty(B, T, F, dtype=odt, device=dev) for _ in range(ring)],
                  "lens": [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(ring)],
                  "k": 0}
            self._ws[key] = ws
            while len(self._ws) > 16:
                self._ws.popitem(last=False)
        else:
            self._ws.move_to_end(key)
        k = ws["k"]
        ws["k"] = (k + 1) % len(ws["out"])
        return ws["raw"], ws["out"][k], ws["lens"][k]

    def _draw_augmentation(self, lens, rng):
        """The reference's draws, per utterance, in its order (speech_utils.py:245-266): np.random.choice over a
        list of ratios or a uniform stretch, then np.random.randint for the noise level."""
        a = self._aug
        sr = self.params["sample_freq"]
        B = len(lens)
        sr_new = np.zeros(B, dtype=np.int32)
        noise = np.zeros(B, dtype=np.float32)
        n_out = np.asarray(lens, dtype=np.int64).copy()
        for b in range(B):
            if "speed_perturbation_ratio" in a:
                r = a["speed_perturbation_ratio"]
                stretch = -1
                if isinstance(r, list):
                    stretch = rng.choice(r)
                elif r > 0:
                    stretch = 1.0 + (2.0 * rng.rand() - 1.0) * r
                if stretch > 0:
                    sr_new[b] = int(sr * stretch)
                    n_out[b] = int(lens[b] * (float(sr_new[b]) / sr))
            if "noise_level_min" in a and "noise_level_max" in a:
                db = rng.randint(low=a["noise_level_min"], high=a["noise_level_max"])
                noise[b] = 10.0 ** (db / 20.0)
        return sr_new, noise, n_out

    def _draw_one(self, n, rng):
        """(sr_new, noise amplitude, augmented length) of one utterance; (0, 0, n) without augmentation."""
        if self._aug is None:
            return (0, 0.0, n)
        sr_new, noise, n_out = self._draw_augmentation([n], rng)
        return (int(sr_new[0]), float(noise[0]), int(n_out[0]))

    def _draw_masks(self, n_frames, rng):
        """speech_utils.py:419-433 in the reference's draw order -> int32 [n_masks, 3] (kind, base, width)."""
        a = self._aug
        F = self.params["num_audio_features"]
        out = np.zeros((self._n_masks, 3), dtype=np.int32)
        k = 0
        for _ in range(int(a.get("n_freq_mask", 0))):
            band = rng.randint(a.get("width_freq_mask", 10) + 1)
            base = rng.randint(0, F - band)
            out[k] = (0, base, band)
            k += 1
        for _ in range(int(a.get("n_time_mask", 0))):
            band = rng.randint(a.get("width_time_mask", 50) + 1)
            if n_frames - band > 0:
                base = rng.randint(n_frames - band)
                out[k] = (1, base, band)
            k += 1
        return out

    def featurize(self, waves, seed=0, rng=None, draws=None):
        """waves: list of int16 numpy arrays, or (pinned int16 tensor holding them back to back, lengths).
        Returns (features [B,T_pad,F] in the 16-bit feature format, lengths int32 [B]), both on the device,
        produced on the CURRENT stream."""
        from openseq2seq_b200 import _lib as L
        if self._dev is None:
            self._setup_device_tables()
        lib = L.load()
        p = self.params
        staged = None
        if isinstance(waves, tuple):
            staged, lens = waves
            lens = np.asarray(lens, dtype=np.int64)
        else:
            lens = np.asarray([len(w) for w in waves], dtype=np.int64)
        B = len(lens)
        F = p["num_audio_features"]
        pad_to = p.get("pad_to", 8)
        total = int(lens.sum())
        aug = self._aug is not None
        if aug:
            rng = rng or np.random
            # draws: per-utterance (sr_new, noise amplitude, n_out) made when the utterance was read (the
            # duration filter of the reference sees the augmented length, speech_utils.py:356)
            sr_new, noise, n_out = draws if draws is not None else self._draw_augmentation(lens, rng)
        else:
            n_out = lens
        max_n = int(n_out.max())
        if self._psf:
            T = speech_utils.psf_num_frames(max_n, self.n_win, self.n_hop, pad_to)
        else:
            T = 1 + max_n // self.n_hop
        if pad_to > 0 and T % pad_to:
            T += pad_to - T % pad_to
        masks = None
        if aug and self._n_masks:
            masks = np.stack([self._draw_masks(1 + int(n) // self.n_hop, rng) for n in n_out])
        # ---- one pinned staging buffer: [waveforms | offsets i64 | out offsets i64 | n i32 | n_out i32 | sr_new i32 |
        #      noise f32 | masks i32]; two async copies (waveforms, metadata)
        meta_i64 = np.zeros((2, B), dtype=np.int64)
        meta_i64[0, 1:] = np.cumsum(lens[:-1])
        meta_i64[1, 1:] = np.cumsum(n_out[:-1])
        meta_i32 = np.zeros((3, B), dtype=np.int32)
        meta_i32[0] = lens
        meta_i32[1] = n_out
        if aug:
            meta_i32[2] = sr_new
        wave_bytes = (total * 2 + 15) // 16 * 16
        nm = self._n_masks if masks is not None else 0
        meta_bytes = meta_i64.nbytes + meta_i32.nbytes + B * 4 + B * nm * 12
        host = self._pinned.get(wave_bytes + meta_bytes)
        hv = host.numpy()
        if staged is None:
            np.concatenate(waves, out=hv[:total * 2].view(np.int16))
        o = wave_bytes
        hv[o:o + meta_i64.nbytes] = meta_i64.view(np.uint8).reshape(-1)
        o64 = o
        o += meta_i64.nbytes
        hv[o:o + meta_i32.nbytes] = meta_i32.view(np.uint8).reshape(-1)
        o32 = o
        o += meta_i32.nbytes
        on = o
        hv[o:o + B * 4] = (noise if aug else np.zeros(B, np.float32)).view(np.uint8)
        o += B * 4
        om = o
        if nm:
            hv[o:o + B * nm * 12] = masks.reshape(-1).view(np.uint8)
        dwave = self._device_buffer("wave", wave_bytes + meta_bytes + 64, torch.uint8)
        if staged is None:
            dwave[:wave_bytes + meta_bytes].copy_(host[:wave_bytes + meta_bytes], non_blocking=True)
        else:
            # the caller's own pinned waveform buffer: copied from where it lies
            dwave[:total * 2].view(torch.int16).copy_(staged[:total], non_blocking=True)
            dwave[wave_bytes:wave_bytes + meta_bytes].copy_(host[wave_bytes:wave_bytes + meta_bytes], non_blocking=True)
        self._pinned.mark()
        base = dwave.data_ptr()
        vp = ctypes.c_void_p
        p_wave, p_off, p_ooff = vp(base), vp(base + o64), vp(base + o64 + 8 * B)
        p_n, p_nout, p_srn = vp(base + o32), vp(base + o32 + 4 * B), vp(base + o32 + 8 * B)
        p_noise, p_masks = vp(base + on), vp(base + om if nm else 0)
        raw, out, out_lens = self._out_buffers(B, T)
        absmax = self._device_buffer_small(B)
        st = L.stream_ptr()
        dither = float(p.get("dither", 0.0)) if p["mode"] == "train" or p.get("dither", 0.0) else 0.0
        if self._psf:
            dither = 0.0   # get_speech_features_psf takes no dither (speech_utils.py:444-449): silently unused
        dt = 1 if self.feature_dtype == "fp16" else 0
        sig_ptr, sigoff_ptr, n_ptr = vp(0), vp(0), p_n
        if aug:
            sig = self._device_buffer("sig", int(n_out.sum()) + 16, torch.float32)
            if self._gain <= 0.0:
                L.check(lib.os2s_wave_absmax(p_wave, p_off, p_n, B, L.ptr(absmax), st), "os2s_wave_absmax")
            L.check(lib.os2s_augment_signal(
                p_wave, p_off, p_n, B, L.ptr(absmax), ctypes.c_float(self._gain),
                p_srn if self._resample_tab is not None else vp(0), int(p["sample_freq"]),
                L.ptr(self._resample_tab), int(self._resample_tab.numel()) if self._resample_tab is not None else 0,
                int(self._num_table) if self._resample_tab is not None else 0, p_noise,
                ctypes.c_uint64((seed * 2654435761 + 97) & 0xFFFFFFFFFFFFFFFF), L.ptr(sig), p_ooff, p_nout, max_n, st),
                "os2s_augment_signal")
            sig_ptr, sigoff_ptr, n_ptr = L.ptr(sig), p_ooff, p_nout
        L.check(lib.os2s_features_forward_p(
            p_wave, sig_ptr, sigoff_ptr, p_off, n_ptr, B, L.ptr(self._mel), L.ptr(self._band), L.ptr(self._win),
            self.n_fft, self.n_win, self.n_hop, F, T, max_n, ctypes.c_float(dither), ctypes.c_uint64(seed),
            ctypes.c_float(0.97), int(self._psf), int(pad_to), int(self._per_feature), ctypes.c_float(self._gain),
            L.ptr(self._fixed_mean), L.ptr(self._fixed_std), p_masks, nm, self._ftype, L.ptr(self._post), self._n_filt,
            L.ptr(absmax), L.ptr(raw), L.ptr(out), None, L.ptr(out_lens), dt, st), "os2s_features_forward_p")
        self.h2d_bytes = wave_bytes + meta_bytes
        self._last = dict(B=B, T=T, max_n=max_n, dither=dither, aug=aug, nm=nm, n_out_sum=int(n_out.sum()),
                          ptrs=(p_wave, p_off, p_ooff, p_n, p_nout, p_srn, p_noise, p_masks), keep=dwave)
        return out, out_lens

    def _device_buffer_small(self, B):
        b = self._dev_buf.get("absmax")
        if b is None or b.numel() < B:
            b = torch.zeros(max(B, 64), dtype=torch.int32, device=self._dev)
            self._dev_buf["absmax"] = b
        return b

    def featurize_resident(self, seed=0):
        """Re-run augmentation + featurizer on the waveforms of the previous featurize() call, which are still
        resident in HBM (no host->device copy).  Used by bench.py for the device-resident timing."""
        from openseq2seq_b200 import _lib as L
        lib = L.load()
        p = self.params
        s = self._last
        B, T, max_n = s["B"], s["T"], s["max_n"]
        (p_wave, p_off, p_ooff, p_n, p_nout, p_srn, p_noise, p_masks) = s["ptrs"]
        F = p["num_audio_features"]
        raw, out, out_lens = self._out_buffers(B, T)
        absmax = self._device_buffer_small(B)
        st = L.stream_ptr()
        vp = ctypes.c_void_p
        dt = 1 if self.feature_dtype == "fp16" else 0
        sig_ptr, sigoff_ptr, n_ptr = vp(0), vp(0), p_n
        if s["aug"]:
            sig = self._device_buffer("sig", s["n_out_sum"] + 16, torch.float32)
            if self._gain <= 0.0:
                L.check(lib.os2s_wave_absmax(p_wave, p_off, p_n, B, L.ptr(absmax), st), "os2s_wave_absmax")
            L.check(lib.os2s_augment_signal(
                p_wave, p_off, p_n, B, L.ptr(absmax), ctypes.c_float(self._gain),
                p_srn if self._resample_tab is not None else vp(0), int(p["sample_freq"]),
                L.ptr(self._resample_tab), int(self._resample_tab.numel()) if self._resample_tab is not None else 0,
                int(self._num_table) if self._resample_tab is not None else 0, p_noise,
                ctypes.c_uint64((seed * 2654435761 + 97) & 0xFFFFFFFFFFFFFFFF), L.ptr(sig), p_ooff, p_nout, max_n, st),
                "os2s_augment_signal")
            sig_ptr, sigoff_ptr, n_ptr = L.ptr(sig), p_ooff, p_nout
        L.check(lib.os2s_features_forward_p(
            p_wave, sig_ptr, sigoff_ptr, p_off, n_ptr, B, L.ptr(self._mel), L.ptr(self._band), L.ptr(self._win),
            self.n_fft, self.n_win, self.n_hop, F, T, max_n, ctypes.c_float(s["dither"]), ctypes.c_uint64(seed),
            ctypes.c_float(0.97), int(self._psf), int(p.get("pad_to", 8)), int(self._per_feature),
            ctypes.c_float(self._gain), L.ptr(self._fixed_mean), L.ptr(self._fixed_std), p_masks, s["nm"], self._ftype,
            L.ptr(self._post), self._n_filt, L.ptr(absmax), L.ptr(raw), L.ptr(out), None, L.ptr(out_lens), dt, st),
            "os2s_features_forward_p")
        return out, out_lens

    # ------------------------------------------------------------------ batching
    def _load(self, entry):
        fn, tr = entry
        p = self.params
        if fn.startswith("synthetic:"):
            hit = self._syn_cache.get(fn)
            if hit is not None:
                return hit
            _, i, secs, seed = fn.split(":")
            g = np.random.default_rng(int(seed) + 7919 * int(i))
            n = int(float(secs) * p["sample_freq"])
            sig = np.clip(3000.0 * g.standard_normal(n), -32768, 32767).astype(np.int16)
            chars = list(p["char2idx"].keys())
            L = int(g.integers(int(12 * float(secs)), int(17.3 * float(secs)) + 1))
            ids = g.integers(1, len(chars), size=L)  # no leading/trailing constraints needed for CTC
            out = (sig, np.asarray(ids, dtype=np.int32))
            if len(self._syn_cache) < 4096:
                self._syn_cache[fn] = out
            return out
        sig = speech_utils.read_wav(fn, p["sample_freq"])
        ids = np.array([p["char2idx"][c] for c in tr], dtype=np.int32) if tr is not None else np.zeros(0, np.int32)
        return sig, ids

    def build_graph(self):
        """Creates the batch iterator (the reference builds a tf.data graph here, speech2text.py:217-324)."""
        self._stop_producer()
        if self.params.get("interactive", False) or not torch.cuda.is_available() or self.prefetch <= 0:
            self._iterator = self._batches_sync()
        else:
            self._iterator = self._batches_async()

    def _index_batches(self, rng):
        """Yields (epoch, [dataset indices]) in the reference's order: shuffle -> repeat -> duration filters ->
        batch (the duration filter needs the decoded length, so it is applied after reading)."""
        p = self.params
        order = np.arange(len(self._files))
        epoch = 0
        while True:
            if p["shuffle"]:
                rng.shuffle(order)
            yield epoch, list(order)
            epoch += 1
            if not p.get("repeat", p["mode"] == "train"):
                return

    def _keep(self, n_samples):
        p = self.params
        dur = n_samples / float(p["sample_freq"])
        if p["max_duration"] > 0 and dur > p["max_duration"]:
            return False
        if p["min_duration"] > 0 and dur < p["min_duration"]:
            return False
        return True

    def _seed0(self):
        return (self._model.params.get("random_seed", 0) if self._model is not None else 0) + (self._worker_id or 0)

    def _batches_sync(self):
        p = self.params
        B = p["batch_size"]
        rng = np.random.default_rng(self._seed0())
        aug_rng = np.random.RandomState(self._seed0() + 17)
        for epoch, order in self._index_batches(rng):
            batch = []
            for idx in order:
                sig, ids = self._load(self._files[idx])
                draw = self._draw_one(len(sig), aug_rng)
                if not self._keep(draw[2]):
                    continue
                batch.append((sig, ids, idx, draw))
                if len(batch) == B:
                    yield self._collate(batch, epoch, aug_rng)
                    batch = []
            if batch and p["mode"] != "train":
                yield self._collate(batch, epoch, aug_rng)

    def _collate(self, batch, epoch, aug_rng=None):
        if self.params["mode"] == "train":
            # longest first (the order inside a batch carries no meaning): utterances of similar length sit next
            # to each other, so the tail tiles of the short ones pair up and are skipped by the conv kernels
            batch = sorted(batch, key=lambda b: -(b[3][2] if len(b) > 3 else len(b[0])))
        sigs = [b[0] for b in batch]
        Lmax = max(1, max(len(b[1]) for b in batch))
        y = np.zeros((len(batch), Lmax), dtype=np.int32)
        ylen = np.zeros(len(batch), dtype=np.int32)
        for i, b in enumerate(batch):
            y[i, :len(b[1])] = b[1]
            ylen[i] = len(b[1])
        draws = None
        if self._aug is not None:
            draws = (np.asarray([b[3][0] for b in batch], dtype=np.int32), np.asarray([b[3][1] for b in batch], dtype=np.float32),
                     np.asarray([b[3][2] for b in batch], dtype=np.int64))
        feats, lens = self.featurize(sigs, seed=epoch * 1000003 + int(batch[0][2]), rng=aug_rng, draws=draws)
        out = {"source_tensors": [feats, lens]}
        if self.params["mode"] != "infer":
            host = self._pinned.get(len(batch) * (Lmax + 1) * 4 + 64)
            hy = host[:y.nbytes].numpy().view(np.int32).reshape(y.shape)
            hy[...] = y
            hl = host[y.nbytes:y.nbytes + ylen.nbytes].numpy().view(np.int32)
            hl[...] = ylen
            ty = torch.empty(y.shape, dtype=torch.int32, device=feats.device)
            tl = torch.empty(ylen.shape, dtype=torch.int32, device=feats.device)
            ty.copy_(host[:y.nbytes].view(torch.int32).view(y.shape), non_blocking=True)
            tl.copy_(host[y.nbytes:y.nbytes + ylen.nbytes].view(torch.int32), non_blocking=True)
            self._pinned.mark()
            out["target_tensors"] = [ty, tl]
        else:
            out["source_ids"] = [np.array([b[2] for b in batch])]
        self._input_tensors = out
        return out

    # -- asynchronous pipeline
    def _stop_producer(self):
        pr = self._producer
        if pr is not None:
            pr["stop"].set()
            try:
                while True:
                    pr["q"].get_nowait()
            except queue.Empty:
                pass
            pr["thread"].join(timeout=5.0)
            self._producer = None

    def _batches_async(self):
        """Producer thread: reader pool -> collate into pinned buffers -> H2D + augmentation + featurizer on a side
        stream; the consumer makes its stream wait on the batch's event."""
        from concurrent.futures import ThreadPoolExecutor
        p = self.params
        B = p["batch_size"]
        if self._dev is None:
            self._setup_device_tables()
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()
        side = torch.cuda.Stream()
        device = torch.cuda.current_device()
        # released[k]: event on the CONSUMER's stream after everything it enqueued for batch k.  The output ring
        # has prefetch + 2 slots and the queue holds at most `prefetch` batches, so when the producer starts
        # batch n the consumer has already come back for batch n - (prefetch + 1), i.e. released[n - ring] exists;
        # the side stream waits on it before the slot is overwritten (the consumer's GPU work may lag far behind
        # its host thread).
        released = []
        ring = self.prefetch + 2
        n_made = [0]

        def guard_slot():
            k = n_made[0] - ring
            if k >= 0:
                side.wait_event(released[k])
            n_made[0] += 1

        def produce():
            try:
                torch.cuda.set_device(device)
                rng = np.random.default_rng(self._seed0())
                aug_rng = np.random.RandomState(self._seed0() + 17)
                with ThreadPoolExecutor(max_workers=self.reader_threads) as pool:
                    for epoch, order in self._index_batches(rng):
                        batch = []
                        # read ahead in chunks of 2 batches so that the pool always has work
                        for c0 in range(0, len(order), 2 * B):
                            chunk = order[c0:c0 + 2 * B]
                            loaded = list(pool.map(lambda i: self._load(self._files[i]), chunk))
                            for idx, (sig, ids) in zip(chunk, loaded):
                                if stop.is_set():
                                    return
                                draw = self._draw_one(len(sig), aug_rng)
                                if not self._keep(draw[2]):
                                    continue
                                batch.append((sig, ids, idx, draw))
                                if len(batch) == B:
                                    with torch.cuda.stream(side):
                                        guard_slot()
                                        out = self._collate(batch, epoch, aug_rng)
                                        ev = torch.cuda.Event()
                                        ev.record(side)
                                    q.put((out, ev))
                                    batch = []
                        if batch and p["mode"] != "train":
                            with torch.cuda.stream(side):
                                guard_slot()
                                out = self._collate(batch, epoch, aug_rng)
                                ev = torch.cuda.Event()
                                ev.record(side)
                            q.put((out, ev))
                q.put((None, None))
            except BaseException as e:  # surfaced in the consumer
                q.put((e, None))

        th = threading.Thread(target=produce, daemon=True, name="os2s-data-producer")
        self._producer = {"q": q, "stop": stop, "thread": th}
        th.start()
        try:
            while True:
                out, ev = q.get()
                if out is None:
                    return
                if isinstance(out, BaseException):
                    raise out
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                for t in out.get("target_tensors", []):
                    t.record_stream(cur)      # allocated on the side stream, consumed on this one
                self._input_tensors = out
                yield out
                rel = torch.cuda.Event()
                rel.record(torch.cuda.current_stream())
                released.append(rel)
        finally:
            stop.set()

    @property
    def iterator(self):
        return self._iterator

    @property
    def input_tensors(self):
        return self._input_tensors
