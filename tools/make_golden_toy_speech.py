"""Copy the reference's own toy speech fixtures (8 WSJ-like wavs, the 10-row CSV, the vocabulary)
into tests/golden/toy_speech_data so the GPU box -- which has no /root/reference -- can run the
reference's train-to-convergence integration test (open_seq2seq/models/speech2text_test.py:89-103,
speech2text_w2l_test.py:23-24) against this implementation.  Run in the build container only."""
import os
import shutil

import pandas as pd

SRC = "/root/reference/open_seq2seq/test_utils/toy_speech_data"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "toy_speech_data")
os.makedirs(os.path.join(DST, "wav_files"), exist_ok=True)
for f in sorted(os.listdir(os.path.join(SRC, "wav_files"))):
    shutil.copy(os.path.join(SRC, "wav_files", f), os.path.join(DST, "wav_files", f))
shutil.copy(os.path.join(SRC, "vocab.txt"), os.path.join(DST, "vocab.txt"))
csv = pd.read_csv(os.path.join(SRC, "toy_data.csv"))
# paths relative to the fixture directory (the reference's are relative to its repo root)
csv["wav_filename"] = [os.path.join("wav_files", os.path.basename(p)) for p in csv["wav_filename"]]
csv.to_csv(os.path.join(DST, "toy_data.csv"), index=False)
print(csv.shape, sorted(os.listdir(os.path.join(DST, "wav_files"))))
