"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the CTC pieces of the reference's speech2text path:

  open_seq2seq/losses/ctc_loss.py:12-16   dense_to_sparse (labels = first tgt_length[b] ids of row b)
  open_seq2seq/losses/ctc_loss.py:77-89   tf.nn.ctc_loss(..., ignore_longer_outputs_than_inputs=True),
                                          mask_nans, reduce_mean over the batch
  open_seq2seq/utils/utils.py:366-370     mask_nans
  open_seq2seq/decoders/fc_decoders.py:244-251  tf.nn.ctc_greedy_decoder(merge_repeated=True)

tf.nn.ctc_loss itself is TensorFlow 1.x (tensorflow/core/util/ctc, not vendored, not installed).
Its published algorithm is restated: unnormalised time-major logits [T,B,V], softmax inside,
blank = V-1, standard Graves alpha/beta recursion in log space over the blank-augmented label
sequence l' (length 2L+1) with the repeated-label skip rule, per-utterance loss -log p(l|x), gradient
wrt logits  softmax - (1/p) * sum_{s: l'_s = v} alpha_t(s) beta_t(s)  for t < len, zero afterwards;
an utterance with no valid alignment (needs L + repeats > len) contributes loss 0 / grad 0 when
ignore_longer_outputs_than_inputs=True.

Pinned against the reference's own golden vector ctc_decoder_with_lm/ctc-test.pickle
(tests/golden/ctc_test_logits.npz): greedy -> "then seconds", sum of max logits 7079.117
(ctc-test.py:64-67) and -log p("then seconds") = 1.1842575 (ctc-test.py:73).
"""
import numpy as np


def log_softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    y = x - m
    return y - np.log(np.sum(np.exp(y), axis=axis, keepdims=True))


def _logaddexp(a, b):
    return np.logaddexp(a, b)


def ctc_loss_and_grad(logits, labels, label_lens, input_lens, blank=None):
    """logits [T,B,V] (float), labels [B,Lmax] int, label_lens [B], input_lens [B].

    Returns (loss [B] float64, grad [T,B,V] float64 = d sum_b loss_b / d logits)."""
    logits = np.asarray(logits, dtype=np.float64)
    T, B, V = logits.shape
    if blank is None:
        blank = V - 1
    logp = log_softmax(logits, axis=2)
    loss = np.zeros(B)
    grad = np.zeros_like(logits)
    NEG = -np.inf
    for b in range(B):
        Tb = int(input_lens[b])
        L = int(label_lens[b])
        lab = np.asarray(labels[b][:L], dtype=np.int64)
        repeats = int(np.sum(lab[1:] == lab[:-1])) if L > 1 else 0
        if Tb == 0 or L + repeats > Tb:
            continue  # ignore_longer_outputs_than_inputs=True -> zero loss, zero grad
        S = 2 * L + 1
        ext = np.full(S, blank, dtype=np.int64)
        ext[1::2] = lab
        lp = logp[:Tb, b, :]
        alpha = np.full((Tb, S), NEG)
        alpha[0, 0] = lp[0, ext[0]]
        if S > 1:
            alpha[0, 1] = lp[0, ext[1]]
        skip_ok = np.zeros(S, dtype=bool)
        skip_ok[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        for t in range(1, Tb):
            a = alpha[t - 1]
            acc = a.copy()
            acc[1:] = _logaddexp(acc[1:], a[:-1])
            sk = np.full(S, NEG)
            sk[2:] = a[:-2]
            acc = np.where(skip_ok, _logaddexp(acc, sk), acc)
            alpha[t] = acc + lp[t, ext]
        beta = np.full((Tb, S), NEG)
        beta[Tb - 1, S - 1] = lp[Tb - 1, ext[S - 1]]
        if S > 1:
            beta[Tb - 1, S - 2] = lp[Tb - 1, ext[S - 2]]
        skip_ok_b = np.zeros(S, dtype=bool)
        skip_ok_b[:-2] = (ext[:-2] != blank) & (ext[:-2] != ext[2:])
        for t in range(Tb - 2, -1, -1):
            bt = beta[t + 1]
            acc = bt.copy()
            acc[:-1] = _logaddexp(acc[:-1], bt[1:])
            sk = np.full(S, NEG)
            sk[:-2] = bt[2:]
            acc = np.where(skip_ok_b, _logaddexp(acc, sk), acc)
            beta[t] = acc + lp[t, ext]
        ll = alpha[Tb - 1, S - 1]
        if S > 1:
            ll = _logaddexp(ll, alpha[Tb - 1, S - 2])
        loss[b] = -ll
        if not np.isfinite(ll):
            continue
        # grad = softmax - exp(log sum_{s in class v} alpha*beta - lp - ll)
        ab = alpha + beta  # includes lp twice
        g = np.exp(lp)
        for v in np.unique(ext):
            sel = ext == v
            lab_sum = np.logaddexp.reduce(ab[:, sel], axis=1)
            g[:, v] -= np.exp(lab_sum - lp[:, v] - ll)
        grad[:Tb, b, :] = g
    return loss, grad


def mask_nans(x):
    """utils/utils.py:366-370: NaN -> 0."""
    x = np.asarray(x, dtype=np.float64).copy()
    x[np.isnan(x)] = 0.0
    return x


def ctc_loss_mean(logits, labels, label_lens, input_lens):
    """CTCLoss._compute_loss (ctc_loss.py:77-89): (scalar mean loss, grad of the mean wrt logits)."""
    loss, grad = ctc_loss_and_grad(logits, labels, label_lens, input_lens)
    B = logits.shape[1]
    nan = np.isnan(loss)
    loss = mask_nans(loss)
    grad[:, nan, :] = 0.0
    return float(np.mean(loss)), grad / B


def ctc_greedy_decode(logits, input_lens, blank=None, merge_repeated=True):
    """tf.nn.ctc_greedy_decoder: returns (list of token lists, neg_sum_logits [B])."""
    logits = np.asarray(logits)
    T, B, V = logits.shape
    if blank is None:
        blank = V - 1
    out, score = [], np.zeros(B)
    for b in range(B):
        prev = -1
        toks = []
        for t in range(int(input_lens[b])):
            row = logits[t, b]
            c = int(np.argmax(row))  # first maximum on ties
            score[b] -= float(row[c])
            if c != blank and not (merge_repeated and c == prev):
                toks.append(c)
            prev = c
        out.append(toks)
    return out, score
