"""Model check of the two-phase gradient exchange of csrc/peer.cu (no GPU): N ranks, each with a compute stream
(backward produces the buckets in order, the optimizer consumes them) and a side stream (the exact operation list
of os2s_peer_exchange_bucket / os2s_peer_finish), executed under RANDOM interleavings -- any stream of any rank
whose head operation is enabled may run next, so ranks drift up to a whole step apart.  Checked on every run:

  * no deadlock (some operation is always enabled until every rank has finished every step);
  * the optimizer of every rank and step sees exactly sum_r g_r for every bucket (identical on all ranks);
  * the reuse argument of the file header: a staging slot is never refilled before the slice sum has read it, and
    a gradient slice is never overwritten by a peer's summed slice before its own contribution has left.

The slicing (split_bucket) is the host-side mirror of peer.cu's PeerExchange::slice.
"""
import random

import numpy as np
import pytest

from openseq2seq_b200.dist import split_bucket


class Rank(object):
    def __init__(self, r, world, buckets, steps, seed):
        self.r, self.world, self.buckets, self.steps = r, world, buckets, steps
        total = max(e for _, e in buckets)
        self.g = np.zeros(total)
        self.rng = np.random.RandomState(seed)
        self.flags = np.zeros((2, len(buckets), world), dtype=np.int64)      # raised by the peers
        self.expected = np.zeros((2, len(buckets)), dtype=np.int64)
        self.stage = {}          # (bucket, src) -> array: the staging slots
        self.stage_unread = set()
        self.contrib_left = set()    # (step, bucket, dst): my contribution to dst's slice has been copied out
        self.produced = {}       # step -> list of per-bucket local gradients (for the expected sums)
        self.main = self._main_ops()
        self.side = []           # filled as backward hands buckets over (event record -> side stream)
        self.main_pos = self.side_pos = 0
        self.side_done_step = -1     # last step whose os2s_peer_finish has completed
        self.seen = {}           # step -> copy of g at optimizer time

    def _main_ops(self):
        ops = []
        for s in range(self.steps):
            for b in range(len(self.buckets)):
                ops.append(("produce", s, b))
            ops.append(("optimizer", s))
        return ops


def run(world, buckets, steps, seed):
    sched = random.Random(seed)
    ranks = [Rank(r, world, buckets, steps, 100 * seed + r) for r in range(world)]
    slices = [split_bucket(s, e, world) for (s, e) in buckets]

    def side_ops(s, b, r):
        ops = [("push0", s, b, p) for p in range(world) if p != r]
        ops += [("signal", s, 0, b), ("wait", s, 0, b), ("sum", s, b)]
        ops += [("push1", s, b, p) for p in range(world) if p != r]
        ops += [("signal", s, 1, b)]
        if b == len(buckets) - 1:
            ops += [("finish", s)]
        return ops

    def enabled(R, op):
        kind = op[0]
        if kind == "wait":
            _, s, ph, b = op
            return all(R.flags[ph, b, p] >= R.expected[ph, b] + 1 for p in range(world) if p != R.r)
        if kind == "finish":
            return all(R.flags[1, b, p] >= R.expected[1, b] + 1
                       for b in range(len(buckets)) for p in range(world) if p != R.r)
        if kind == "optimizer":
            return R.side_done_step >= op[1]      # the compute stream waits for the side stream
        return True

    def execute(R, op):
        kind = op[0]
        if kind == "produce":
            _, s, b = op
            lo, hi = buckets[b]
            R.g[lo:hi] = R.rng.randint(-8, 9, size=hi - lo)        # wgrad overwrites (beta = 0)
            R.produced.setdefault(s, {})[b] = R.g[lo:hi].copy()
            R.side.extend(side_ops(s, b, R.r))
        elif kind == "push0":
            _, s, b, p = op
            lo, hi = slices[b][p]
            P = ranks[p]
            assert (b, R.r) not in P.stage_unread, "staging slot refilled before the slice sum read it"
            P.stage[(b, R.r)] = R.g[lo:hi].copy()
            P.stage_unread.add((b, R.r))
            R.contrib_left.add((s, b, p))
        elif kind == "signal":
            _, s, ph, b = op
            for p in range(world):
                if p != R.r:
                    ranks[p].flags[ph, b, R.r] += 1
        elif kind == "wait":
            _, s, ph, b = op
            R.expected[ph, b] += 1
        elif kind == "sum":
            _, s, b = op
            lo, hi = slices[b][R.r]
            for p in range(world):
                if p != R.r:
                    R.g[lo:hi] += R.stage[(b, p)]
                    R.stage_unread.discard((b, p))
        elif kind == "push1":
            _, s, b, p = op
            lo, hi = slices[b][R.r]
            P = ranks[p]
            assert (s, b, R.r) in P.contrib_left, "gradient slice overwritten before the peer's contribution left"
            P.g[lo:hi] = R.g[lo:hi]
        elif kind == "finish":
            R.expected[1, :] += 1
            R.side_done_step = op[1]
        elif kind == "optimizer":
            R.seen[op[1]] = R.g.copy()

    while True:
        ready = []
        for R in ranks:
            if R.main_pos < len(R.main) and enabled(R, R.main[R.main_pos]):
                ready.append((R, "main"))
            if R.side_pos < len(R.side) and enabled(R, R.side[R.side_pos]):
                ready.append((R, "side"))
        if not ready:
            break
        R, which = sched.choice(ready)
        if which == "main":
            execute(R, R.main[R.main_pos])
            R.main_pos += 1
        else:
            execute(R, R.side[R.side_pos])
            R.side_pos += 1
    for R in ranks:
        assert R.main_pos == len(R.main) and R.side_pos == len(R.side), "deadlock"
    for s in range(steps):
        for b, (lo, hi) in enumerate(buckets):
            want = sum(R.produced[s][b] for R in ranks)
            for R in ranks:
                assert np.array_equal(R.seen[s][lo:hi], want), (s, b, R.r)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_two_phase_exchange_is_deadlock_free_and_sums_under_random_interleavings(world):
    # tail-first buckets, one of them smaller than world * 4 elements (some ranks own an empty slice)
    buckets = [(96, 163), (64, 96), (12, 64), (0, 12)]
    for seed in range(25):
        run(world, buckets, steps=3, seed=seed)
