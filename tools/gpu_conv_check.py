"""Bring-up check of the tcgen05 conv kernels against torch (fp32 cuDNN) on a B200.
Run under gpurun; each case runs in its own subprocess with a timeout so a hung kernel
cannot eat the whole call.  Usage: python tools/gpu_conv_check.py [case ...]"""

import json
import subprocess
import sys
import time

CASES = {
    # name: (mode, B, T, Cin, Cout, K, dil)
    "fwd_1x1_tiny": ("fwd", 1, 128, 64, 64, 1, 1),
    "fwd_1x1": ("fwd", 2, 752, 256, 256, 1, 1),
    "fwd_k11": ("fwd", 2, 752, 256, 256, 11, 1),
    "fwd_k29d2": ("fwd", 2, 752, 768, 896, 29, 2),
    "fwd_k13_384": ("fwd", 2, 300, 256, 384, 13, 1),
    "fwd_k21_640": ("fwd", 2, 300, 640, 640, 21, 1),
    "dgrad_k11": ("dgrad", 2, 752, 256, 256, 11, 1),
    "dgrad_k13": ("dgrad", 2, 300, 256, 384, 13, 1),
    "dgrad_acc": ("dgrad_acc", 2, 300, 256, 384, 1, 1),
    "wgrad_tiny": ("wgrad", 1, 64, 128, 64, 1, 1),
    "wgrad_1x1": ("wgrad", 2, 752, 256, 256, 1, 1),
    "wgrad_k11": ("wgrad", 4, 752, 256, 256, 11, 1),
    "wgrad_k29d2": ("wgrad", 2, 300, 768, 896, 29, 2),
    "wgrad_k13": ("wgrad", 3, 300, 256, 384, 13, 1),
}
PERF = {
    # Jasper 10x5 shapes at B=32, T=752
    "perf_k11_256": (32, 752, 256, 256, 11, 1),
    "perf_k13_384": (32, 752, 384, 384, 13, 1),
    "perf_k17_512": (32, 752, 512, 512, 17, 1),
    "perf_k21_640": (32, 752, 640, 640, 21, 1),
    "perf_k25_768": (32, 752, 768, 768, 25, 1),
    "perf_k29_896": (32, 752, 768, 896, 29, 2),
    "perf_1x1_768": (32, 752, 768, 768, 1, 1),
    "perf_1x1_1024": (32, 752, 896, 1024, 1, 1),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    from openseq2seq_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(0)
    dev = "cuda"
    st = L.stream_ptr()

    def same_pad(K, dil):
        tot = (K - 1) * dil
        return tot // 2

    if name in CASES:
        mode, B, T, Cin, Cout, K, dil = CASES[name]
        padl = same_pad(K, dil)
        x = torch.randn(B, T, Cin, device=dev).bfloat16()
        w = (torch.randn(K, Cin, Cout, device=dev) / (K * Cin) ** 0.5).bfloat16()
        dy = torch.randn(B, T, Cout, device=dev).bfloat16()
        wt = w.permute(0, 2, 1).contiguous()
        xf, wf, dyf = x.float(), w.float(), dy.float()
        # torch conv1d wants NCW and weight [Cout, Cin, K]
        wt_torch = wf.permute(2, 1, 0).contiguous()
        xin = xf.permute(0, 2, 1).contiguous().requires_grad_(True)
        wtt = wt_torch.clone().requires_grad_(True)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        yref = F.conv1d(xin, wtt, padding=padl, dilation=dil)
        yref.backward(dyf.permute(0, 2, 1).contiguous())
        y_ref = yref.detach().permute(0, 2, 1).contiguous()
        dx_ref = xin.grad.permute(0, 2, 1).contiguous()
        dw_ref = wtt.grad.permute(2, 1, 0).contiguous()  # [K,Cin,Cout]
        if mode == "fwd":
            y = torch.full((B, T, Cout), float("nan"), device=dev).bfloat16()
            L.check(lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 0, None, st), name)
            got, ref = y.float(), y_ref
        elif mode == "dgrad":
            dx = torch.full((B, T, Cin), float("nan"), device=dev).bfloat16()
            L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, padl, 0, st), name)
            got, ref = dx.float(), dx_ref
        elif mode == "dgrad_acc":
            base = torch.randn(B, T, Cin, device=dev)
            dx = base.clone()
            L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, padl, 2, st), name)
            got, ref = dx, dx_ref + base
        else:
            dw = torch.full((K, Cin, Cout), float("nan"), device=dev)
            L.check(lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, padl, st), name)
            got, ref = dw, dw_ref
        torch.cuda.synchronize()
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item()
        nan = int(torch.isnan(got).sum().item())
        rel = err / max(scale, 1e-9)
        out = {"case": name, "max_abs_err": err, "ref_max": scale, "rel": rel, "nan": nan,
               "ok": bool(rel < 2e-2 and nan == 0)}
        if not out["ok"]:
            # localise: which rows / cols are wrong
            bad = ((got - ref).abs() > 2e-2 * scale) | torch.isnan(got)
            idx = bad.nonzero()
            out["n_bad"] = int(bad.sum().item())
            out["first_bad"] = idx[:5].tolist()
            out["last_bad"] = idx[-5:].tolist()
            flat = idx[:, -1]
            out["bad_lastdim_hist"] = torch.bincount(flat // 16, minlength=4)[:32].tolist()
            if idx.shape[1] == 3:
                out["bad_dim1_hist"] = torch.bincount(idx[:, 1] // 8)[:48].tolist()
            i0 = tuple(idx[0].tolist())
            out["got_ref_first"] = [got[i0].item(), ref[i0].item()]
        print(json.dumps(out), flush=True)
        return
    B, T, Cin, Cout, K, dil = PERF[name]
    padl = same_pad(K, dil)
    x = torch.randn(B, T, Cin, device=dev).bfloat16()
    w = (torch.randn(K, Cin, Cout, device=dev) / (K * Cin) ** 0.5).bfloat16()
    wt = w.permute(0, 2, 1).contiguous()
    dy = torch.randn(B, T, Cout, device=dev).bfloat16()
    y = torch.empty(B, T, Cout, device=dev).bfloat16()
    dx = torch.empty(B, T, Cin, device=dev).bfloat16()
    dw = torch.empty(K, Cin, Cout, device=dev)
    stats = torch.zeros(2, Cout, device=dev)
    flops = 2.0 * B * T * K * Cin * Cout
    res = {"case": name, "gflop": flops / 1e9}
    fns = {
        "fwd": lambda: lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 0, None, st),
        "fwd_stats": lambda: lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 3,
                                                 L.ptr(stats), st),
        "fwd_wt": lambda: lib.os2s_conv1d_fwd_wt(L.ptr(x), L.ptr(wt), L.ptr(y), B, T, Cin, Cout, K, dil, padl, 0, st),
        "dgrad": lambda: lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, padl, 0, st),
        "wgrad": lambda: lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, padl, st),
    }
    # cuDNN / cuBLAS comparison (library baseline, bf16)
    xin = x.permute(0, 2, 1).contiguous()
    wtt = w.permute(2, 1, 0).contiguous()
    fns["torch_fwd"] = lambda: (F.conv1d(xin, wtt, padding=padl, dilation=dil), 0)[1]
    for key, fn in fns.items():
        for _ in range(3):
            L.check(fn(), key)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[key + "_ms"] = round(ms, 4)
        res[key + "_tflops"] = round(flops / ms / 1e9, 1)
    print(json.dumps(res), flush=True)


def main():
    names = sys.argv[1:] or (list(CASES) + list(PERF))
    if len(names) == 1 and names[0].startswith("@"):
        run_case(names[0][1:])
        return
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "@" + n], timeout=120, capture_output=True, text=True)
            out = r.stdout.strip().splitlines()
            print(out[-1] if out else json.dumps({"case": n, "error": r.stderr[-800:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": n, "error": "TIMEOUT (hang)"}), flush=True)
        sys.stderr.write("%s took %.1fs\n" % (n, time.time() - t0))


if __name__ == "__main__":
    sys.path.insert(0, ".")
    main()
