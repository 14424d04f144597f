"""Multi-process check of the peer-memory gradient exchange (csrc/peer.cu) against NCCL's all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/peer_check.py [--floats 40000000] [--iters 6] [--time]

Every rank fills a flat fp32 buffer with its own random numbers, cut into uneven buckets (a tiny one, one that
is not a multiple of world * 4, a large one), sums it through both transports, and checks (a) peer == NCCL to fp32
rounding of a differently-ordered sum, (b) bit-identical results on every rank, (c) several rounds in a row (the
flags count up, the staging slots are reused).  --time prints the duration of one exchange of the whole buffer.
Prints one JSON line on rank 0; exit code 1 on a mismatch.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch
import torch.distributed as dist

from openseq2seq_b200.dist import PeerGradExchange, TorchDistHvd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--floats", type=int, default=40_000_000)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    hvd = TorchDistHvd.init()
    rank, world = hvd.rank(), hvd.size()
    dev = torch.device("cuda", hvd.local_rank())
    n = a.floats // 4 * 4 + 3          # the last bucket ends off a 16-byte boundary
    g = torch.zeros(n, dtype=torch.float32, device=dev)
    # buckets from the tail, as backward produces them: 67 floats (some ranks own an empty slice), 4 * (4 world + 3),
    # a large ragged one, the rest in two halves
    top = n // 4 * 4 - 64
    cuts = [n, top, top - 4 * (4 * world + 3), top - 4 * (4 * world + 3) - 4 * 1000003, n // 2 // 4 * 4, 0]
    buckets = [(cuts[i + 1], cuts[i]) for i in range(len(cuts) - 1)]
    px = PeerGradExchange(hvd, g, buckets, 30.0)
    out = {"world": world, "floats": n, "ok": bool(px.ok), "why": px.why}
    if not px.ok:
        if rank == 0:
            print(json.dumps(out))
        sys.exit(1)
    side = torch.cuda.Stream()
    worst, same = 0.0, True
    for it in range(a.iters):
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 * it + rank)
        g.copy_(torch.randn(n, device=dev, generator=gen) * (1.0 + rank))
        ref = g.clone()
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        # exchange on a side stream while the main stream is busy, as in the training step
        ev = torch.cuda.Event()
        ev.record()
        side.wait_event(ev)
        busy = torch.randn(4096, 4096, device=dev)
        for _ in range(4):
            busy = busy @ busy
            busy = busy / busy.abs().max()
        for b in range(len(buckets)):
            px.exchange_bucket(b, side)
        px.finish(side)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if px.timed_out():
            out["ok"], out["why"] = False, "time-out in round %d" % it
            break
        err = float((g - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        # identical bits everywhere: compare a checksum of the raw words
        chk = g.view(torch.int32).to(torch.int64).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = same and bool(int(lo[0]) == int(hi[0]))
    out["max_rel_err_vs_nccl"] = worst
    out["identical_bits_on_all_ranks"] = same
    out["ok"] = bool(out["ok"] and worst < 1e-6 and same)
    if a.time and out["ok"]:
        for name in ("peer", "nccl"):
            ts = []
            for _ in range(5):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if name == "peer":
                    px.allreduce_all()
                else:
                    dist.all_reduce(g)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name + "_ms"] = round(float(t[0]), 3)
            out[name + "_busbw_GBps"] = round(2 * (world - 1) / world * n * 4 / (float(t[0]) * 1e-3) / 1e9, 1)
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
