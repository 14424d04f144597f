"""Achieved HBM GB/s of the BN kernels on Jasper shapes (algorithmic bytes / CUDA-event time)."""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
from openseq2seq_b200 import _lib as L

lib = L.load()
_vp, _f = ctypes.c_void_p, ctypes.c_float
st = L.stream_ptr()
B, T = 32, 752
M = B * T


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C, nb in [(256, 1), (384, 1), (512, 1), (640, 1), (768, 1), (1024, 1), (768, 4), (768, 11), (512, 7)]:
    ys = [torch.randn(B, T, C, device="cuda").half() for _ in range(nb)]
    # rotate through several copies so the 126 MB L2 does not serve the reads
    stats = torch.zeros(nb, 2, C, device="cuda")
    for j in range(nb):
        lib.os2s_bn_stats(L.ptr(ys[j]), L.ptr(stats[j]), M, C, st)
    gam = [torch.ones(C, device="cuda") for _ in range(nb)]
    bet = [torch.zeros(C, device="cuda") for _ in range(nb)]
    mi = torch.zeros(nb, 2, C, device="cuda")
    out = torch.empty(B, T, C, dtype=torch.bfloat16, device="cuda")
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    arr = lambda ts: (_vp * nb)(*[t.data_ptr() for t in ts])
    a_y, a_s, a_g, a_b, a_mi = arr(ys), arr(list(stats)), arr(gam), arr(bet), arr(list(mi))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def fwd():
        lib.os2s_bn_apply_fwd(nb, a_y, a_s, a_g, a_b, a_mi, None, L.ptr(out), L.ptr(lens), B, T, C, _f(1e-3), _f(0.9),
                              _f(0.8), ctypes.c_uint64(1), 1, _f(0.0), 0, None, st)

    dA = torch.randn(B, T, C, device="cuda").bfloat16()
    dys = [torch.empty(B, T, C, dtype=torch.bfloat16, device="cuda") for _ in range(nb)]
    dg = [torch.zeros(C, device="cuda") for _ in range(nb)]
    db = [torch.zeros(C, device="cuda") for _ in range(nb)]
    red = torch.zeros((1 + nb) * C, device="cuda")
    a_dy, a_dg, a_db = arr(dys), arr(dg), arr(db)

    def bwd():
        lib.os2s_bn_bwd(nb, a_y, a_mi, a_g, a_dg, a_db, a_dy, L.ptr(dA), 0, L.ptr(out), L.ptr(red), M, C, _f(0.8), 1, st)

    def stats_only():
        lib.os2s_bn_stats(L.ptr(ys[0]), L.ptr(stats[0]), M, C, st)

    def with_flush(fn):
        def g():
            flush.zero_()
            fn()
        return g

    t_flush = timeit(lambda: flush.zero_())
    res = {"C": C, "branches": nb}
    for name, fn, nbytes in [("bn_apply_fwd", fwd, M * C * 2 * (nb + 1)),
                             ("bn_bwd", bwd, M * C * 2 * (2 + nb) + M * C * 2 * (2 + 2 * nb)),
                             ("bn_stats", stats_only, M * C * 2)]:
        ms = timeit(with_flush(fn)) - t_flush
        res[name] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 0)}
    print(json.dumps(res), flush=True)
