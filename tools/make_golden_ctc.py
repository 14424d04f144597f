"""Generate tests/golden/ctc_test_logits.npz from the reference's own golden vector
/root/reference/ctc_decoder_with_lm/ctc-test.pickle (run in the build container only; the
GPU box has no /root/reference).  The pinned answers come from
ctc_decoder_with_lm/ctc-test.py:64-67,73 and scripts/ctc_decoders_test.py:58-61."""
import pickle
import numpy as np

with open("/root/reference/ctc_decoder_with_lm/ctc-test.pickle", "rb") as f:
    seq, label = pickle.load(f, encoding="bytes")
seq = np.asarray(seq, dtype=np.float32)
label = label.decode() if isinstance(label, bytes) else str(label)
vocab = [line[0] for line in open("/root/reference/open_seq2seq/test_utils/toy_speech_data/vocab.txt")]
np.savez_compressed(
    "tests/golden/ctc_test_logits.npz",
    logits=seq, label=np.array(label), vocab=np.array(vocab),
    greedy_text=np.array("then seconds"), greedy_neg_sum_logits=np.float64(7079.117),
    ctc_log_prob_then_seconds=np.float64(-1.1842575))
print(seq.shape, seq.dtype, label, len(vocab))
