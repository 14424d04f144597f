"""Speech2TextDataLayer (open_seq2seq/data/speech2text/speech2text.py:25-485) over the GPU featurizer.

Same constructor signature and params schema; CSV (wav_filename, transcript) datasets, vocab file
-> char2idx with `tgt_vocab_size = len(vocab) + 1` (blank last), duration filters, per-worker
seeding / eval sharding (`split_data`), padded batches with `pad_to`.  Instead of a tf.data graph of
py_func featurizer threads, `iterator` yields batches whose waveforms are staged in pinned host
memory, copied to the device and featurised by ONE call of os2s_logmel_forward per batch.

Extension used by bench.py and tests: dataset_files may contain entries of the form
"synthetic:<n_utts>:<seconds>[:<seed>]" which generate band-limited noise utterances and random
transcripts in memory -- there is no network / dataset in the build image."""
import ctypes
import math

import numpy as np
import torch

from open_seq2seq.data.data_layer import DataLayer
from open_seq2seq.data.utils import load_pre_existing_vocabulary
from . import speech_utils


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
        p["char2idx"] = load_pre_existing_vocabulary(p["vocab_file"], read_chars=True)
        p["idx2char"] = {i: w for w, i in p["char2idx"].items()}
        p["tgt_vocab_size"] = len(p["char2idx"]) + 1  # + blank (speech2text.py:120-125)
        self.target_pad_value = 0
        p["min_duration"] = p.get("min_duration", -1.0)
        p["max_duration"] = p.get("max_duration", -1.0)
        p["window_size"] = p.get("window_size", 20e-3)
        p["window_stride"] = p.get("window_stride", 10e-3)
        p["sample_freq"] = p.get("sample_freq", 16000)
        self._files = None
        self._load_file_list()
        self._dev = None
        self._iterator = None
        self._input_tensors = None

    # ------------------------------------------------------------------ dataset
    def _load_file_list(self):
        p = self.params
        rows = []
        self._synthetic = []
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
        self._psf = p.get("backend", "psf") == "psf"
        if p["input_type"] != "logfbank":
            raise NotImplementedError("Speech2TextDataLayer: the GPU featurizer implements input_type='logfbank' "
                                      "(both backends); 'spectrogram' / 'mfcc' are not built")
        # psf normalises with one mean / std per utterance; librosa per feature unless norm_per_feature=False
        self._per_feature = (not self._psf) and bool(p.get("norm_per_feature", False))
        sr = p["sample_freq"]
        self.n_win = int(sr * p["window_size"])
        self.n_hop = int(sr * p["window_stride"])
        F = p["num_audio_features"]
        if self._psf:
            # psf.logfbank(nfft=512, winfunc = rectangular) (speech_utils.py:514-522)
            self.n_fft = 512
            mel = speech_utils.psf_filterbank(F, self.n_fft, sr, 0.0, sr / 2.0)
            win = np.ones(self.n_win)
        else:
            self.n_fft = p.get("num_fft") or speech_utils.num_fft_for(p["window_size"], sr)
            mel = speech_utils.mel_filterbank(sr, self.n_fft, F, 0.0, int(sr / 2))
            win_name = p.get("window", "hanning")
            win = {"hanning": np.hanning, "hamming": np.hamming, "none": np.ones}[win_name](self.n_win)
        self._dev = torch.device("cuda")
        self._mel = torch.tensor(mel, dtype=torch.float32, device=self._dev)
        band = [[int(np.nonzero(r)[0].min()), int(np.nonzero(r)[0].max()) + 1] if np.any(r) else [0, 0] for r in mel]
        self._band = torch.tensor(band, dtype=torch.int32, device=self._dev)
        self._win = torch.tensor(win, dtype=torch.float32, device=self._dev)
        self._ws = {}

    def featurize(self, waves, seed=0):
        """waves: list of int16 numpy arrays (or a pinned int16 tensor + lengths tuple).
        Returns (features bf16 [B,T_pad,F] device, lengths int32 [B] device)."""
        from openseq2seq_b200 import _lib as L
        if self._dev is None:
            self._setup_device_tables()
        lib = L.load()
        p = self.params
        if isinstance(waves, tuple):
            host, lens = waves
        else:
            lens = [len(w) for w in waves]
            host = torch.empty(int(sum(lens)), dtype=torch.int16).pin_memory()
            np.concatenate(waves, out=host.numpy())
        B = len(lens)
        F = p["num_audio_features"]
        max_n = int(max(lens))
        pad_to = p.get("pad_to", 8)
        if self._psf:
            T = speech_utils.psf_num_frames(max_n, self.n_win, self.n_hop, pad_to)
        else:
            T = 1 + max_n // self.n_hop
        if pad_to > 0 and T % pad_to:
            T += pad_to - T % pad_to
        key = (B, T, int(host.numel()))
        ws = self._ws.get(key)
        if ws is None:
            dev = self._dev
            ws = {"wave": torch.empty(host.numel(), dtype=torch.int16, device=dev),
                  "off": torch.empty(B, dtype=torch.int64, device=dev),
                  "n": torch.empty(B, dtype=torch.int32, device=dev),
                  "absmax": torch.zeros(B, dtype=torch.int32, device=dev),
                  "raw": torch.empty(B * T * F, dtype=torch.float32, device=dev),
                  "out": torch.empty(B, T, F, dtype=torch.bfloat16, device=dev),
                  "lens": torch.empty(B, dtype=torch.int32, device=dev)}
            self._ws[key] = ws
        ws["wave"].copy_(host, non_blocking=True)
        offs = np.zeros(B, dtype=np.int64)
        offs[1:] = np.cumsum(lens[:-1])
        ws["off"].copy_(torch.from_numpy(offs), non_blocking=True)
        ws["n"].copy_(torch.tensor(lens, dtype=torch.int32), non_blocking=True)
        dither = float(p.get("dither", 0.0)) if p["mode"] == "train" or p.get("dither", 0.0) else 0.0
        if self._psf:
            dither = 0.0   # get_speech_features_psf takes no dither (speech_utils.py:444-449): silently unused
        L.check(lib.os2s_features_forward(
            L.ptr(ws["wave"]), L.ptr(ws["off"]), L.ptr(ws["n"]), B, L.ptr(self._mel), L.ptr(self._band), L.ptr(self._win),
            self.n_fft,
            self.n_win, self.n_hop, F, T, max_n, ctypes.c_float(dither), ctypes.c_uint64(seed),
            ctypes.c_float(0.97), int(self._psf), int(self.params.get("pad_to", 8)), int(self._per_feature),
            L.ptr(ws["absmax"]), L.ptr(ws["raw"]), L.ptr(ws["out"]), None, L.ptr(ws["lens"]),
            L.stream_ptr()), "os2s_features_forward")
        self.h2d_bytes = host.numel() * 2 + B * 12
        self._last = (ws, B, T, max_n, dither)
        return ws["out"], ws["lens"]

    def featurize_resident(self, seed=0):
        """Re-run the featurizer on the waveforms of the previous featurize() call, which are still
        resident in HBM (no host->device copy).  Used by bench.py for the device-resident timing."""
        from openseq2seq_b200 import _lib as L
        lib = L.load()
        ws, B, T, max_n, dither = self._last
        F = self.params["num_audio_features"]
        L.check(lib.os2s_features_forward(
            L.ptr(ws["wave"]), L.ptr(ws["off"]), L.ptr(ws["n"]), B, L.ptr(self._mel), L.ptr(self._band), L.ptr(self._win),
            self.n_fft,
            self.n_win, self.n_hop, F, T, max_n, ctypes.c_float(dither), ctypes.c_uint64(seed),
            ctypes.c_float(0.97), int(self._psf), int(self.params.get("pad_to", 8)), int(self._per_feature),
            L.ptr(ws["absmax"]), L.ptr(ws["raw"]), L.ptr(ws["out"]), None, L.ptr(ws["lens"]),
            L.stream_ptr()), "os2s_features_forward")
        return ws["out"], ws["lens"]

    # ------------------------------------------------------------------ batching
    def _load(self, entry, rng):
        fn, tr = entry
        p = self.params
        if fn.startswith("synthetic:"):
            _, i, secs, seed = fn.split(":")
            g = np.random.default_rng(int(seed) + 7919 * int(i))
            n = int(float(secs) * p["sample_freq"])
            sig = np.clip(3000.0 * g.standard_normal(n), -32768, 32767).astype(np.int16)
            chars = list(p["char2idx"].keys())
            L = int(g.integers(int(12 * float(secs)), int(17.3 * float(secs)) + 1))
            ids = g.integers(1, len(chars), size=L)  # no leading/trailing constraints needed for CTC
            return sig, np.asarray(ids, dtype=np.int32)
        sig = speech_utils.read_wav(fn, p["sample_freq"])
        ids = np.array([p["char2idx"][c] for c in tr], dtype=np.int32) if tr is not None else np.zeros(0, np.int32)
        return sig, ids

    def build_graph(self):
        """Creates the batch iterator (the reference builds a tf.data graph here, speech2text.py:217-324)."""
        self._iterator = self._batches()

    def _batches(self):
        p = self.params
        B = p["batch_size"]
        seed = (self._model.params.get("random_seed", 0) if self._model is not None else 0) + (self._worker_id or 0)
        rng = np.random.default_rng(seed)
        order = np.arange(len(self._files))
        epoch = 0
        sr = p["sample_freq"]
        while True:
            if p["shuffle"]:
                rng.shuffle(order)
            batch = []
            for idx in order:
                sig, ids = self._load(self._files[idx], rng)
                dur = len(sig) / float(sr)
                if p["max_duration"] > 0 and dur > p["max_duration"]:
                    continue
                if p["min_duration"] > 0 and dur < p["min_duration"]:
                    continue
                batch.append((sig, ids, idx))
                if len(batch) == B:
                    yield self._collate(batch, epoch)
                    batch = []
            if batch and p["mode"] != "train":
                yield self._collate(batch, epoch)
            epoch += 1
            if not p.get("repeat", p["mode"] == "train"):
                return

    def _collate(self, batch, epoch):
        sigs = [b[0] for b in batch]
        Lmax = max(1, max(len(b[1]) for b in batch))
        y = np.zeros((len(batch), Lmax), dtype=np.int32)
        ylen = np.zeros(len(batch), dtype=np.int32)
        for i, b in enumerate(batch):
            y[i, :len(b[1])] = b[1]
            ylen[i] = len(b[1])
        feats, lens = self.featurize(sigs, seed=epoch * 1000003 + int(batch[0][2]))
        out = {"source_tensors": [feats, lens]}
        if self.params["mode"] != "infer":
            out["target_tensors"] = [torch.from_numpy(y).to(feats.device, non_blocking=True),
                                     torch.from_numpy(ylen).to(feats.device, non_blocking=True)]
        else:
            out["source_ids"] = [np.array([b[2] for b in batch])]
        self._input_tensors = out
        return out

    @property
    def iterator(self):
        return self._iterator

    @property
    def input_tensors(self):
        return self._input_tensors
