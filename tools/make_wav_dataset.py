"""Write a file-backed speech dataset of synthetic utterances (16 kHz int16 wav files + the CSV the
Speech2TextDataLayer reads): `python tools/make_wav_dataset.py <dir> <n_utts> <seconds>`.
Used to exercise the asynchronous input pipeline (reader threads -> pinned ring -> side-stream featurizer)
with real file I/O: `OS2S_DATASET_CSV=<dir>/data.csv python run.py --config_file=configs/jasper10x5_files.py
--benchmark --bench_steps=40` must report the audio-seconds/sec of bench.py's end-to-end figure."""
import os
import sys

import numpy as np
import scipy.io.wavfile as wavfile


def main():
    out, n, secs = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    os.makedirs(os.path.join(out, "wav"), exist_ok=True)
    g = np.random.default_rng(77)
    vocab = " abcdefghijklmnopqrstuvwxyz'"
    rows = ["wav_filename,wav_filesize,transcript"]
    for i in range(n):
        sig = np.clip(3000.0 * g.standard_normal(int(secs * 16000)), -32768, 32767).astype(np.int16)
        fn = os.path.join(out, "wav", "utt%05d.wav" % i)
        wavfile.write(fn, 16000, sig)
        L = int(g.integers(int(12 * secs), int(17.3 * secs) + 1))
        tr = "".join(vocab[int(c)] for c in g.integers(1, 27, size=L))   # letters only: no CSV quoting issues
        rows.append("%s,%d,%s" % (fn, os.path.getsize(fn), tr))
    with open(os.path.join(out, "data.csv"), "w") as f:
        f.write("\n".join(rows) + "\n")
    print(os.path.join(out, "data.csv"))


if __name__ == "__main__":
    main()
