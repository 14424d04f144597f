"""CTA-pair (cta_group::2) conv kernel vs the single-CTA kernel: bitwise output comparison on a few
shapes, then sustained TFLOP/s of both.  The mode is read once per process (OS2S_CONV_PAIR), so every
leg runs in its own subprocess with a timeout (a hung kernel cannot eat the gpurun call).
Usage: python tools/gpu_pair_check.py            (parent)
"""
import json
import os
import subprocess
import sys
import time

SHAPES = [
    # B, T, Cin, Cout, K, dil
    (2, 300, 256, 256, 11, 1),
    (3, 752, 256, 384, 13, 1),
    (2, 752, 640, 640, 21, 1),
    (2, 1000, 768, 896, 29, 2),
    (4, 752, 896, 1024, 1, 1),
    (1, 129, 128, 256, 3, 1),
]
MODES = {
    "single": {"OS2S_CONV_PAIR": "0", "OS2S_CONV_HALO": "0"},
    "pair": {"OS2S_CONV_PAIR": "2", "OS2S_CONV_HALO": "0"},
    "halo": {"OS2S_CONV_PAIR": "2", "OS2S_CONV_HALO": "1"},
}
PERF = [
    (32, 752, 256, 256, 11, 1),
    (32, 752, 384, 384, 13, 1),
    (32, 752, 512, 512, 17, 1),
    (32, 752, 640, 640, 21, 1),
    (32, 752, 768, 768, 25, 1),
    (32, 752, 768, 896, 29, 2),
    (32, 752, 896, 1024, 1, 1),
]


def child_outputs(path):
    import torch
    sys.path.insert(0, ".")
    from openseq2seq_b200 import _lib as L
    lib = L.load()
    st = L.stream_ptr()
    out = {}
    for i, (B, T, Cin, Cout, K, dil) in enumerate(SHAPES):
        torch.manual_seed(100 + i)
        padl = ((K - 1) * dil) // 2
        x = torch.randn(B, T, Cin, device="cuda").bfloat16()
        w = (torch.randn(K, Cin, Cout, device="cuda") / (K * Cin) ** 0.5).bfloat16()
        dy = torch.randn(B, T, Cout, device="cuda").bfloat16()
        y = torch.full((B, T, Cout), float("nan"), device="cuda", dtype=torch.float16)
        stats = torch.zeros(2, Cout, device="cuda")
        L.check(lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 3, L.ptr(stats), st), "fwd")
        dx = torch.full((B, T, Cin), float("nan"), device="cuda").bfloat16()
        L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, padl, 0, st), "dgrad")
        acc = torch.ones(B, T, Cin, device="cuda")
        L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(acc), B, T, Cin, Cout, K, dil, padl, 2, st), "dgrad_acc")
        dw = torch.full((K, Cin, Cout), float("nan"), device="cuda")
        if Cin % 128 == 0:
            L.check(lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, padl, st), "wgrad")
            out["w%d" % i] = dw.cpu()
        torch.cuda.synchronize()
        out["y%d" % i] = y.cpu()
        out["s%d" % i] = stats.cpu()
        out["dx%d" % i] = dx.cpu()
        out["acc%d" % i] = acc.cpu()
    torch.save(out, path)


def child_perf():
    import torch
    sys.path.insert(0, ".")
    from openseq2seq_b200 import _lib as L
    lib = L.load()
    st = L.stream_ptr()
    for (B, T, Cin, Cout, K, dil) in PERF:
        padl = ((K - 1) * dil) // 2
        x = torch.randn(B, T, Cin, device="cuda").bfloat16()
        w = (torch.randn(K, Cin, Cout, device="cuda") / (K * Cin) ** 0.5).bfloat16()
        y = torch.empty(B, T, Cout, device="cuda", dtype=torch.float16)
        dy = torch.randn(B, T, Cout, device="cuda").bfloat16()
        dx = torch.empty(B, T, Cin, device="cuda").bfloat16()
        stats = torch.zeros(2, Cout, device="cuda")
        dw = torch.empty(K, Cin, Cout, device="cuda")
        flops = 2.0 * B * T * K * Cin * Cout
        res = {"pair": os.environ.get("OS2S_CONV_PAIR", "0"), "halo": os.environ.get("OS2S_CONV_HALO", "0"),
               "shape": [Cin, Cout, K]}
        for name, fn in (
            ("fwd", lambda: lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 3, L.ptr(stats), st)),
            ("dgrad", lambda: lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, padl, 0, st)),
            ("wgrad", lambda: lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, padl, st)),
        ):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # ~1.5 s of back-to-back launches: the clock settles at the power-capped value
            n = max(20, int(1.5 / (flops / 1.2e15)))
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = round(flops * n / e0.elapsed_time(e1) / 1e9, 1)
        print(json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "out":
        return child_outputs(sys.argv[2])
    if len(sys.argv) > 1 and sys.argv[1] == "perf":
        return child_perf()
    import torch
    ok = True
    for mode in MODES:
        env = dict(os.environ, **MODES[mode])
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "out", "/tmp/pair_%s.pt" % mode], env=env, timeout=240,
                               capture_output=True, text=True)
            if r.returncode != 0:
                print(json.dumps({"mode": mode, "error": r.stderr[-800:]}))
                ok = False
        except subprocess.TimeoutExpired:
            print(json.dumps({"mode": mode, "error": "timeout"}))
            ok = False
        print(json.dumps({"mode": mode, "secs": round(time.time() - t0, 1)}), flush=True)
    if not ok:
        return 1
    a = torch.load("/tmp/pair_single.pt")
    allsame = True
    for mode in list(MODES)[1:]:
        b = torch.load("/tmp/pair_%s.pt" % mode)
        print(json.dumps({"comparing": mode}))
        for k in sorted(a):
            if k.startswith("s") or k.startswith("w"):
                rel = ((a[k] - b[k]).abs().max() / a[k].abs().max()).item()
                same = rel < 1e-5 and not bool(torch.isnan(b[k]).any())
                print(json.dumps({"tensor": k, "max_rel": rel, "ok": same}))
            else:
                same = torch.equal(a[k].view(torch.int16) if a[k].element_size() == 2 else a[k].view(torch.int32),
                                   b[k].view(torch.int16) if b[k].element_size() == 2 else b[k].view(torch.int32))
                nbad = 0 if same else int((a[k].float() != b[k].float()).sum())
                print(json.dumps({"tensor": k, "bitwise_equal": same, "mismatches": nbad,
                                  "nan": bool(torch.isnan(b[k].float()).any())}))
            allsame &= same
    print(json.dumps({"pair_matches_single": allsame}), flush=True)
    if not allsame:
        return 1
    for mode in MODES:
        env = dict(os.environ, **MODES[mode])
        try:
            r = subprocess.run([sys.executable, __file__, "perf"], env=env, timeout=300, capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                print(json.dumps({"mode": mode, "perf_error": r.stderr[-500:]}))
        except subprocess.TimeoutExpired as e:
            print(json.dumps({"mode": mode, "perf_error": "timeout", "partial": (e.stdout or b"").decode()[-1500:]}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
