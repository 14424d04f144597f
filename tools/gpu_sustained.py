"""Sustained (power-capped) throughput of the conv kernel vs cuBLAS on the K25/768 Jasper shape:
run each for ~3 s and report TFLOP/s plus SM clock / power sampled during the run."""
import json
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from openseq2seq_b200 import _lib as L

lib = L.load()
st = L.stream_ptr()


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for l in self.p.stdout:
            try:
                a, b = l.split(",")
                self.rows.append((float(a), float(b)))
            except Exception:
                pass

    def stop(self):
        self.p.terminate()
        r = self.rows[5:] or self.rows
        return {"sm_mhz": sorted(x[0] for x in r)[len(r) // 2], "power_w": sorted(x[1] for x in r)[len(r) // 2]}


def sustained(fn, flops, secs=3.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d = s.stop()
    d["tflops"] = round(flops * n / ms / 1e9, 1)
    return d


B, T, K, dil = 32, 752, 25, 1
for Cin, Cout in [(768, 768), (640, 640), (512, 512)]:
    padl = (K - 1) // 2
    x = torch.randn(B, T, Cin, device="cuda").bfloat16()
    w = (torch.randn(K, Cin, Cout, device="cuda") / (K * Cin) ** 0.5).bfloat16()
    y = torch.empty(B, T, Cout, device="cuda").bfloat16()
    dy = torch.randn(B, T, Cout, device="cuda").bfloat16()
    dw = torch.empty(K, Cin, Cout, device="cuda")
    flops = 2.0 * B * T * K * Cin * Cout
    res = {"shape": [Cin, Cout, K]}
    res["fwd"] = sustained(lambda: lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 3, None, st), flops)
    res["dgrad"] = sustained(lambda: lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(x), B, T, Cin, Cout, K, dil, padl, 0, st), flops)
    res["wgrad"] = sustained(lambda: lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, padl, st), flops)
    a = torch.randn(B * T, K * Cin, device="cuda").bfloat16()
    bm = torch.randn(K * Cin, Cout, device="cuda").bfloat16()
    res["cublas_same_gemm"] = sustained(lambda: torch.matmul(a, bm), flops)
    print(json.dumps(res), flush=True)
a = torch.randn(8192, 8192, device="cuda").bfloat16()
b = torch.randn(8192, 8192, device="cuda").bfloat16()
print(json.dumps({"cublas_8192": sustained(lambda: torch.matmul(a, b), 2.0 * 8192 ** 3)}), flush=True)
