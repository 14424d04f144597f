"""Per-kernel GPU parity tests: every C-ABI entry point against the CPU oracle on the same seeded
inputs (sizes the oracle finishes in seconds) and against the committed golden fixtures."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

_vp = ctypes.c_void_p
_f = ctypes.c_float
_ll = ctypes.c_longlong


def _lib():
    from openseq2seq_b200 import _lib as L
    return L, L.load()


def _bf(x):
    return torch.as_tensor(x, dtype=torch.float32).bfloat16()


@pytest.mark.parametrize("B,T,Cin,Cout,K,dil", [
    (2, 70, 128, 64, 5, 1),      # ragged T (not a multiple of the 128-row tile)
    (1, 300, 256, 384, 13, 1),
    (2, 131, 128, 256, 9, 2),    # dilation
    (3, 64, 256, 128, 1, 1),     # 1x1 residual conv
    (2, 200, 640, 640, 3, 1),    # ragged N tiling 256 + 256 + 128 (forward, dgrad and wgrad)
    (1, 130, 384, 896, 1, 1),    # 896 = 3 x 256 + 128
])
def test_conv_fwd_dgrad_wgrad_vs_oracle(B, T, Cin, Cout, K, dil):
    from oracle import encoder as E
    L, lib = _lib()
    rng = np.random.default_rng(0)
    x = _bf(rng.standard_normal((B, T, Cin)))
    w = _bf(rng.standard_normal((K, Cin, Cout)) / np.sqrt(K * Cin))
    dy = _bf(rng.standard_normal((B, T, Cout)))
    pl = ((K - 1) * dil) // 2
    y_ref = E.conv1d_same(x.float().numpy(), w.float().numpy(), 1, dil)
    # oracle gradients by the adjoint identities of the same restated conv
    wf = w.float().numpy().astype(np.float64)
    dx_ref = E.conv1d_same(dy.float().numpy(), np.ascontiguousarray(wf[::-1].transpose(0, 2, 1)), 1, dil)
    xp = np.zeros((B, T + 2 * pl, Cin))
    xp[:, pl:pl + T] = x.float().numpy()
    dw_ref = np.stack([np.einsum("btc,bto->co", xp[:, k * dil:k * dil + T], dy.float().numpy().astype(np.float64))
                       for k in range(K)])
    xd, wd, dyd = x.cuda(), w.cuda(), dy.cuda()
    wt = wd.permute(0, 2, 1).contiguous()
    st = L.stream_ptr()
    y = torch.empty(B, T, Cout, dtype=torch.bfloat16, device="cuda")
    L.check(lib.os2s_conv1d_fwd(L.ptr(xd), L.ptr(wd), L.ptr(y), B, T, Cin, Cout, K, dil, pl, 0, None, st), "fwd")
    y16 = torch.empty(B, T, Cout, dtype=torch.float16, device="cuda")
    fused = torch.zeros(2, Cout, device="cuda")
    L.check(lib.os2s_conv1d_fwd(L.ptr(xd), L.ptr(wd), L.ptr(y16), B, T, Cin, Cout, K, dil, pl, 3, L.ptr(fused), st),
            "fwd16")
    alone = torch.zeros(2, Cout, device="cuda")
    L.check(lib.os2s_bn_stats(L.ptr(y16), L.ptr(alone), B * T, Cout, st), "bn_stats")
    ywt = torch.empty(B, T, Cout, dtype=torch.bfloat16, device="cuda")
    L.check(lib.os2s_conv1d_fwd_wt(L.ptr(xd), L.ptr(wt), L.ptr(ywt), B, T, Cin, Cout, K, dil, pl, 0, st), "fwd_wt")
    dx = torch.empty(B, T, Cin, dtype=torch.float32, device="cuda")
    L.check(lib.os2s_conv1d_dgrad(L.ptr(dyd), L.ptr(wd), L.ptr(dx), B, T, Cin, Cout, K, dil, pl, 1, st), "dgrad")
    if Cin % 128 == 0:
        dw = torch.empty(K, Cin, Cout, dtype=torch.float32, device="cuda")
        L.check(lib.os2s_conv1d_wgrad(L.ptr(xd), L.ptr(dyd), L.ptr(dw), B, T, Cin, Cout, K, dil, pl, st), "wgrad")
    torch.cuda.synchronize()
    assert np.abs(y.float().cpu().numpy() - y_ref).max() <= 1e-2 * np.abs(y_ref).max()
    assert torch.equal(y, ywt)  # the two weight-operand layouts are the same arithmetic
    # BN statistics fused into the conv epilogue == the standalone statistics kernel == the oracle
    y16d = y16.double().cpu()
    assert torch.allclose(fused.cpu().double()[0], y16d.sum((0, 1)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(fused.cpu().double()[1], (y16d * y16d).sum((0, 1)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(fused, alone, rtol=1e-4, atol=1e-2)
    assert np.abs(y16.float().cpu().numpy() - y_ref).max() <= 1.5e-3 * np.abs(y_ref).max()
    assert np.abs(dx.cpu().numpy() - dx_ref).max() <= 1e-4 * np.abs(dx_ref).max() + 1e-4
    if Cin % 128 == 0:
        assert np.abs(dw.cpu().numpy() - dw_ref).max() <= 1e-4 * np.abs(dw_ref).max() + 1e-4


@pytest.mark.parametrize("B,T,Cin,Cout,K,dil", [
    (2, 300, 256, 256, 11, 1),
    (3, 752, 256, 384, 13, 1),   # odd number of (tap, C_in tile) row blocks: the pair wgrad pads one
    (2, 1000, 768, 896, 29, 2),  # largest halo tile (128 + 56 rows), ragged last N tile
    (1, 129, 128, 256, 3, 1),    # second CTA of the pair owns a single valid row
    (2, 260, 896, 1024, 1, 1),
])
def test_conv_kernel_variants_agree(B, T, Cin, Cout, K, dil):
    """single-CTA tiles, CTA pairs (cta_group::2) and pairs with a shared halo tile are the same
    arithmetic in the same accumulation order: forward / data-gradient outputs must be bitwise equal
    (weight gradients differ only by the order of the stream-K fp32 reductions)."""
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, T, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K, Cin, Cout, device="cuda", generator=g) / (K * Cin) ** 0.5).bfloat16()
    dy = torch.randn(B, T, Cout, device="cuda", generator=g).bfloat16()
    pl = ((K - 1) * dil) // 2
    st = L.stream_ptr()
    outs = {}
    try:
        for name, (pm, hm) in {"single": (0, 0), "pair": (2, 0), "halo": (2, 1)}.items():
            assert lib.os2s_conv_tuning(pm, hm) == 0
            y = torch.full((B, T, Cout), float("nan"), dtype=torch.float16, device="cuda")
            stats = torch.zeros(2, Cout, device="cuda")
            L.check(lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), B, T, Cin, Cout, K, dil, pl, 3, L.ptr(stats), st), name)
            dx = torch.full((B, T, Cin), float("nan"), dtype=torch.bfloat16, device="cuda")
            L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), B, T, Cin, Cout, K, dil, pl, 0, st), name)
            acc = torch.ones(B, T, Cin, device="cuda")
            L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(acc), B, T, Cin, Cout, K, dil, pl, 2, st), name)
            dw = torch.full((K, Cin, Cout), float("nan"), device="cuda")
            L.check(lib.os2s_conv1d_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), B, T, Cin, Cout, K, dil, pl, st), name)
            torch.cuda.synchronize()
            outs[name] = (y, stats, dx, acc, dw)
    finally:
        lib.os2s_conv_tuning(1, 1)
    ref = outs["single"]
    assert not torch.isnan(ref[0].float()).any() and not torch.isnan(ref[4]).any()
    for name in ("pair", "halo"):
        y, stats, dx, acc, dw = outs[name]
        assert torch.equal(y.view(torch.int16), ref[0].view(torch.int16)), name
        assert torch.equal(dx.view(torch.int16), ref[2].view(torch.int16)), name
        assert torch.equal(acc, ref[3]), name
        assert torch.allclose(stats, ref[1], rtol=1e-5, atol=1e-3), name
        assert (dw - ref[4]).abs().max() <= 1e-5 * ref[4].abs().max(), name
    assert lib.os2s_conv_tuning(3, 0) != 0


@pytest.mark.parametrize("B,T,Cin,Cout,K,dil,modes", [
    (2, 300, 384, 384, 13, 1, [(0, 0), (1, 1)]),   # single-CTA and pair + halo epilogues
    (3, 131, 256, 128, 3, 1, [(1, 1)]),            # narrow layer: single-CTA kernel, ragged rows
    (2, 260, 640, 896, 1, 1, [(1, 0)]),            # 1x1, pair kernel without halo
])
def test_dgrad_with_fused_bn_backward_reductions(B, T, Cin, Cout, K, dil, modes):
    """os2s_conv1d_dgrad_bnred == os2s_conv1d_dgrad (bitwise dx) and its epilogue sums equal
    sum dz and sum dz*y computed from the stored tensors; os2s_bn_bwd_apply on those sums == os2s_bn_bwd."""
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = torch.randn(B, T, Cout, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K, Cin, Cout, device="cuda", generator=g) / (K * Cout) ** 0.5).bfloat16()
    y = torch.randn(B, T, Cin, device="cuda", generator=g).half()
    a = torch.relu(torch.randn(B, T, Cin, device="cuda", generator=g)).bfloat16()   # ~half zeros
    a[:, T - 7:] = 0                                                                   # masked tail rows
    keep = 0.8
    pl = ((K - 1) * dil) // 2
    st = L.stream_ptr()
    try:
        for pm, hm in modes:
            assert lib.os2s_conv_tuning(pm, hm) == 0
            dx0 = torch.empty(B, T, Cin, dtype=torch.bfloat16, device="cuda")
            L.check(lib.os2s_conv1d_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx0), B, T, Cin, Cout, K, dil, pl, 0, st), "dgrad")
            dx1 = torch.empty_like(dx0)
            red = torch.zeros(2, Cin, device="cuda")
            L.check(lib.os2s_conv1d_dgrad_bnred(L.ptr(dy), L.ptr(w), L.ptr(dx1), B, T, Cin, Cout, K, dil, pl, L.ptr(a),
                                                L.ptr(y), _f(keep), L.ptr(red), st), "dgrad_bnred")
            torch.cuda.synchronize()
            assert torch.equal(dx0.view(torch.int16), dx1.view(torch.int16))
            dz = dx1.double() * (a != 0).double() / keep
            ref0, ref1 = dz.sum((0, 1)), (dz * y.double()).sum((0, 1))
            scale0, scale1 = dz.abs().sum((0, 1)).max(), (dz * y.double()).abs().sum((0, 1)).max()
            assert (red[0].double() - ref0).abs().max() <= 1e-5 * scale0
            assert (red[1].double() - ref1).abs().max() <= 1e-5 * scale1
    finally:
        lib.os2s_conv_tuning(1, 1)
    # apply-only BN backward on the fused sums == the two-pass kernel
    M = B * T
    mi = torch.stack([y.float().mean((0, 1)), 1.0 / torch.sqrt(y.float().var((0, 1), unbiased=False) + 1e-3)]).contiguous()
    gam = (1.0 + 0.1 * torch.randn(Cin, device="cuda", generator=g)).contiguous()
    outs = []
    for fused in (False, True):
        dyo = torch.empty(B, T, Cin, dtype=torch.bfloat16, device="cuda")
        dg, db = torch.zeros(Cin, device="cuda"), torch.zeros(Cin, device="cuda")
        if fused:
            L.check(lib.os2s_bn_bwd_apply(L.ptr(y), L.ptr(mi), L.ptr(gam), L.ptr(dg), L.ptr(db), L.ptr(dyo), L.ptr(dx1),
                                          L.ptr(a), L.ptr(red), M, Cin, _f(keep), st), "bn_bwd_apply")
        else:
            scratch = torch.zeros(2 * Cin, device="cuda")
            arr = lambda t: (_vp * 1)(t.data_ptr())
            L.check(lib.os2s_bn_bwd(1, arr(y), arr(mi), arr(gam), arr(dg), arr(db), arr(dyo), L.ptr(dx1), 0, L.ptr(a),
                                    L.ptr(scratch), M, Cin, _f(keep), 1, st), "bn_bwd")
        torch.cuda.synchronize()
        outs.append((dyo.float(), dg, db))
    assert (outs[0][0] - outs[1][0]).abs().max() <= 2e-2 * outs[0][0].abs().max()
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-4 * float(outs[0][1].abs().max()))
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-4, atol=1e-4 * float(outs[0][2].abs().max()))


def test_conv_rejects_unsupported_shapes_loudly():
    L, lib = _lib()
    x = torch.zeros(1, 16, 48, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(1, 48, 64, dtype=torch.bfloat16, device="cuda")
    y = torch.zeros(1, 16, 64, dtype=torch.bfloat16, device="cuda")
    rc = lib.os2s_conv1d_fwd(L.ptr(x), L.ptr(w), L.ptr(y), 1, 16, 48, 64, 1, 1, 0, 0, None, L.stream_ptr())
    assert rc == -3 and b"multiple of 64" in lib.os2s_last_error()
    assert lib.os2s_conv1d_fwd(None, None, None, 1, 16, 64, 64, 1, 1, 0, 0, None, L.stream_ptr()) == -1


def test_batchnorm_residual_relu_dropout_mask_fwd_bwd_vs_oracle():
    from oracle import torch_twin as TT
    L, lib = _lib()
    B, T, C, nb = 3, 50, 128, 3
    g = torch.Generator().manual_seed(0)
    ys = [(torch.randn(B, T, C, generator=g) * (1 + j) + 0.3 * j).half() for j in range(nb)]
    gam = [1 + 0.3 * torch.randn(C, generator=g) for _ in range(nb)]
    bet = [0.2 * torch.randn(C, generator=g) for _ in range(nb)]
    lens = torch.tensor([50, 31, 7], dtype=torch.int32)
    dA = torch.randn(B, T, C, generator=g)
    # oracle (fp64 autograd over the restated BN)
    yd = [y.double().requires_grad_(True) for y in ys]
    gd = [x.double().requires_grad_(True) for x in gam]
    bd = [x.double().requires_grad_(True) for x in bet]
    tot = 0
    for j in range(nb):
        tot = tot + TT.batch_norm_train(yd[j], gd[j], bd[j], 1e-3)[0]
    out_ref = torch.relu(tot) * TT.sequence_mask(lens.long(), T, torch.float64)
    out_ref.backward(dA.double())
    # device
    st = L.stream_ptr()
    dev = "cuda"
    yc = [y.to(dev) for y in ys]
    stats = torch.zeros(nb, 2, C, device=dev)
    for j in range(nb):
        L.check(lib.os2s_bn_stats(L.ptr(yc[j]), L.ptr(stats[j]), B * T, C, st), "stats")
    arr = lambda ts: (_vp * nb)(*[t.data_ptr() for t in ts])
    gc = [x.to(dev) for x in gam]
    bc = [x.to(dev) for x in bet]
    mi = torch.zeros(nb, 2, C, device=dev)
    mv = torch.zeros(nb, 2, C, device=dev)
    mv[:, 1] = 1
    out = torch.empty(B, T, C, dtype=torch.bfloat16, device=dev)
    lens_c = lens.to(dev)
    L.check(lib.os2s_bn_apply_fwd(nb, arr(yc), arr(list(stats)), arr(gc), arr(bc), arr(list(mi)), arr(list(mv)),
                                  L.ptr(out), L.ptr(lens_c), B, T, C, _f(1e-3), _f(0.9), _f(1.0),
                                  ctypes.c_uint64(1), 1, _f(0.0), 0, None, st), "apply")
    dgam = [torch.zeros(C, device=dev) for _ in range(nb)]
    dbet = [torch.zeros(C, device=dev) for _ in range(nb)]
    dy = [torch.empty(B, T, C, dtype=torch.bfloat16, device=dev) for _ in range(nb)]
    red = torch.zeros((1 + nb) * C, device=dev)
    dAc = dA.to(dev)
    L.check(lib.os2s_bn_bwd(nb, arr(yc), arr(list(mi)), arr(gc), arr(dgam), arr(dbet), arr(dy), L.ptr(dAc), 1,
                            L.ptr(out), L.ptr(red), B * T, C, _f(1.0), 1, st), "bwd")
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
    assert rel(out, out_ref.detach()) < 1e-2
    for j in range(nb):
        assert rel(dy[j], yd[j].grad) < 2e-2
        assert rel(dgam[j], gd[j].grad) < 2e-2
        assert rel(dbet[j], bd[j].grad) < 2e-2
    # moving statistics (momentum 0.9, Bessel-corrected variance)
    n = B * T
    m0 = ys[0].double().mean((0, 1))
    v0 = ys[0].double().var((0, 1), unbiased=True)
    assert rel(mv[0, 0], 0.1 * m0) < 1e-3 and rel(mv[0, 1], 0.9 + 0.1 * v0) < 1e-3


def test_dropout_statistics_and_backward_mask():
    L, lib = _lib()
    B, T, C = 4, 64, 256
    y = torch.randn(B, T, C, device="cuda").half()
    stats = torch.zeros(2, C, device="cuda")
    st = L.stream_ptr()
    L.check(lib.os2s_bn_stats(L.ptr(y), L.ptr(stats), B * T, C, st), "stats")
    one = lambda t: (_vp * 1)(t.data_ptr())
    gam = torch.ones(C, device="cuda")
    bet = torch.full((C,), 3.0, device="cuda")  # keep everything positive: zeros == dropped
    mi = torch.zeros(2, C, device="cuda")
    outs = []
    for seed in (7, 7, 8):
        out = torch.empty(B, T, C, dtype=torch.bfloat16, device="cuda")
        L.check(lib.os2s_bn_apply_fwd(1, one(y), one(stats), one(gam), one(bet), one(mi), None, L.ptr(out), None,
                                      B, T, C, _f(1e-3), _f(0.9), _f(0.7), ctypes.c_uint64(seed), 1, _f(0.0), 0, None, st),
                "apply")
        outs.append(out.float())
    torch.cuda.synchronize()
    kept = (outs[0] != 0).float().mean().item()
    assert abs(kept - 0.7) < 0.01
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
    nodrop = torch.empty(B, T, C, dtype=torch.bfloat16, device="cuda")
    L.check(lib.os2s_bn_apply_fwd(1, one(y), one(stats), one(gam), one(bet), one(mi), None, L.ptr(nodrop), None,
                                  B, T, C, _f(1e-3), _f(0.9), _f(1.0), ctypes.c_uint64(0), 1, _f(0.0), 0, None, st), "apply")
    torch.cuda.synchronize()
    m = outs[0] != 0
    assert torch.allclose(outs[0][m], nodrop.float()[m] / 0.7, rtol=2e-2)


def test_fc_fwd_bwd_vs_oracle():
    from oracle import encoder as E
    L, lib = _lib()
    M, H, V = 77, 256, 29
    rng = np.random.default_rng(1)
    x = _bf(rng.standard_normal((M, H)))
    w = torch.tensor(rng.standard_normal((H, V)) / 16, dtype=torch.float32)
    b = torch.tensor(rng.standard_normal(V), dtype=torch.float32)
    dl = torch.tensor(rng.standard_normal((M, V)), dtype=torch.float32)
    ref = E.fc_decode(x.float().numpy()[None], w.numpy(), b.numpy())[:, 0]
    st = L.stream_ptr()
    xc, wc, bc, dlc = x.cuda(), w.cuda(), b.cuda(), dl.cuda()
    logits = torch.empty(M, V, device="cuda")
    L.check(lib.os2s_fc_fwd(L.ptr(xc), L.ptr(wc), L.ptr(bc), L.ptr(logits), M, H, V, st), "fc_fwd")
    dx = torch.empty(M, H, dtype=torch.bfloat16, device="cuda")
    dw = torch.empty(H, V, device="cuda")
    db = torch.empty(V, device="cuda")
    L.check(lib.os2s_fc_bwd(L.ptr(xc), L.ptr(dlc), L.ptr(wc), L.ptr(dx), L.ptr(dw), L.ptr(db), M, H, V, st), "fc_bwd")
    torch.cuda.synchronize()
    assert np.abs(logits.cpu().numpy() - ref).max() < 1e-4
    xd = x.double().numpy()
    assert np.abs(dx.float().cpu().numpy() - dl.double().numpy() @ w.double().numpy().T).max() < 2e-2
    assert np.abs(dw.cpu().numpy() - xd.T @ dl.double().numpy()).max() < 1e-3
    assert np.abs(db.cpu().numpy() - dl.double().numpy().sum(0)).max() < 1e-4


def _run_ctc(logits_tbv, labels, label_lens, in_lens, scale=None):
    """logits_tbv: numpy [T,B,V]; device call uses a batch-major copy with explicit strides."""
    L, lib = _lib()
    T, B, V = logits_tbv.shape
    lg = torch.tensor(np.ascontiguousarray(logits_tbv.transpose(1, 0, 2)), dtype=torch.float32, device="cuda")
    Lmax = labels.shape[1]
    lab = torch.tensor(labels, dtype=torch.int32, device="cuda")
    ll = torch.tensor(label_lens, dtype=torch.int32, device="cuda")
    il = torch.tensor(in_lens, dtype=torch.int32, device="cuda")
    grad = torch.empty(B, T, V, device="cuda")
    loss = torch.empty(B, device="cuda")
    need = lib.os2s_ctc_workspace_bytes(B, T, Lmax)
    ws = torch.empty(int(need), dtype=torch.uint8, device="cuda")
    sc = torch.tensor([scale], dtype=torch.float32, device="cuda") if scale else None
    L.check(lib.os2s_ctc_loss_fwd_bwd(L.ptr(lg), L.ptr(lab), L.ptr(ll), L.ptr(il), L.ptr(grad), L.ptr(loss),
                                      L.ptr(ws), ctypes.c_size_t(int(need)), L.ptr(sc), B, T, V, Lmax,
                                      _ll(T * V), _ll(V), L.stream_ptr()), "ctc")
    torch.cuda.synchronize()
    return loss.cpu().numpy(), grad.cpu().numpy().transpose(1, 0, 2), lg


def test_ctc_loss_and_greedy_on_reference_golden_vector(golden_dir):
    """The reference's own known answers: ctc_decoder_with_lm/ctc-test.py:64-67,73."""
    L, lib = _lib()
    g = np.load(os.path.join(golden_dir, "ctc_test_logits.npz"))
    lg = g["logits"]  # [184,1,29]
    vocab = list(g["vocab"])
    lab = np.array([[vocab.index(c) for c in "then seconds"]])
    loss, grad, lgd = _run_ctc(lg, lab, [lab.shape[1]], [lg.shape[0]])
    assert abs(loss[0] - (-float(g["ctc_log_prob_then_seconds"]))) < 1e-3
    T, B, V = lg.shape
    il = torch.tensor([T], dtype=torch.int32, device="cuda")
    toks = torch.zeros(B, T, dtype=torch.int32, device="cuda")
    tl = torch.zeros(B, dtype=torch.int32, device="cuda")
    ns = torch.zeros(B, device="cuda")
    L.check(lib.os2s_ctc_greedy(L.ptr(lgd), L.ptr(il), L.ptr(toks), L.ptr(tl), L.ptr(ns), B, T, V, _ll(T * V), _ll(V),
                                1, L.stream_ptr()), "greedy")
    torch.cuda.synchronize()
    text = "".join(vocab[c] for c in toks[0, :int(tl[0])].cpu().tolist())
    assert text == str(g["greedy_text"]) == "then seconds"
    assert abs(float(ns[0]) + float(g["greedy_neg_sum_logits"])) < 1e-2 or \
        abs(float(ns[0]) - float(g["greedy_neg_sum_logits"])) < 1e-2


def test_ctc_loss_grad_vs_oracle_random_ragged_with_repeats_and_infeasible():
    from oracle import ctc as OC
    rng = np.random.default_rng(3)
    T, B, V, Lmax = 60, 5, 29, 20
    lg = rng.standard_normal((T, B, V)).astype(np.float32) * 2
    labels = rng.integers(0, V - 1, size=(B, Lmax))
    labels[1, :6] = [3, 3, 3, 5, 5, 7]        # repeats
    label_lens = [20, 6, 1, 0, 20]
    in_lens = [60, 41, 5, 17, 21]             # last: 20 labels + repeats > 21 frames? made infeasible below
    labels[4, :20] = 4                        # 19 repeats -> needs 39 frames > 21 -> skipped
    ref_loss, ref_grad = OC.ctc_loss_and_grad(lg, labels, label_lens, in_lens)
    ref_mean, ref_gmean = OC.ctc_loss_mean(lg, labels, label_lens, in_lens)
    loss, grad, _ = _run_ctc(lg, labels, label_lens, in_lens, scale=8.0)
    assert ref_loss[4] == 0 and loss[4] == 0
    assert np.abs(loss - ref_loss).max() < 1e-3 * max(1.0, np.abs(ref_loss).max())
    assert np.abs(grad / 8.0 - ref_gmean).max() < 1e-4
    assert np.all(grad[41:, 1] == 0) and np.all(grad[:, 4] == 0)


def test_ctc_greedy_vs_oracle_ragged():
    from oracle import ctc as OC
    L, lib = _lib()
    rng = np.random.default_rng(4)
    T, B, V = 75, 4, 29
    lg = rng.standard_normal((T, B, V)).astype(np.float32)
    lg[:, :, V - 1] += 1.0  # more blanks
    lg[10:14, 0, :] = lg[9, 0, :]  # repeated frames -> merged
    in_lens = [75, 33, 1, 64]
    ref, ref_score = OC.ctc_greedy_decode(lg, in_lens)
    lgd = torch.tensor(np.ascontiguousarray(lg.transpose(1, 0, 2)), device="cuda")
    il = torch.tensor(in_lens, dtype=torch.int32, device="cuda")
    toks = torch.zeros(B, T, dtype=torch.int32, device="cuda")
    tl = torch.zeros(B, dtype=torch.int32, device="cuda")
    ns = torch.zeros(B, device="cuda")
    L.check(lib.os2s_ctc_greedy(L.ptr(lgd), L.ptr(il), L.ptr(toks), L.ptr(tl), L.ptr(ns), B, T, V, _ll(T * V), _ll(V),
                                1, L.stream_ptr()), "greedy")
    torch.cuda.synchronize()
    for b in range(B):
        assert toks[b, :int(tl[b])].cpu().tolist() == ref[b]
    assert np.abs(ns.cpu().numpy() - ref_score).max() < 1e-3


def test_optimizer_chain_vs_oracle_including_overflow_skip():
    """LARC + Backoff scaler + NovoGrad(as written) + poly_decay over several steps, 2 ranks' worth of
    summed gradients, with an injected Inf on step 2 (skip, halve the scale, no step increment)."""
    from oracle import optimizer as OO
    from openseq2seq_b200.engine import JasperEngine
    from tests.common_cfg import MINI_JASPER
    eng = JasperEngine(MINI_JASPER, 64, 29, world_size=2,
                       opt=dict(algo="novograd", beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.001,
                                larc_eta=0.001, learning_rate=0.02, min_lr=1e-5, power=2.0, decay_steps=10,
                                loss_scaling=True))
    names = [n for n, _ in eng.named_parameters()]
    w_ref = [eng.param_view(n).detach().cpu().numpy().copy() for n in names]
    state = OO.NovoGradState(len(names))
    scaler = OO.BackoffScaler()
    step = 0
    rng = np.random.default_rng(0)
    lr_fn = lambda s: OO.poly_decay(s, 0.02, 10, power=2.0, min_lr=1e-5)
    for it in range(5):
        scale = scaler.scale
        assert abs(float(eng.fstate[0]) - scale) < 1e-6
        per_rank = [[(rng.standard_normal(w.shape) * 0.01 * scale).astype(np.float32) for w in w_ref] for _ in range(2)]
        if it == 2:
            per_rank[0][3].flat[5] = np.inf
        eng.grad.zero_()
        for i, n in enumerate(names):
            eng.param_view(n, eng.grad).copy_(torch.tensor(per_rank[0][i] + per_rank[1][i]))
        eng.optimizer_step()
        torch.cuda.synchronize()
        skipped, lr, step = OO.train_step(w_ref, per_rank, state, scaler, step, lr_fn,
                                          dict(beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.001),
                                          larc_params=dict(larc_eta=0.001))
        assert int(eng.istate[3]) == int(skipped)
        assert int(eng.istate[2]) == step
        if not skipped:
            assert abs(float(eng.fstate[1]) - lr) < 1e-6 * max(1.0, lr)
        for i, n in enumerate(names):
            got = eng.param_view(n).cpu().numpy()
            assert np.abs(got - w_ref[i]).max() <= 2e-5 * max(1.0, np.abs(w_ref[i]).max()), (it, n)
    assert int(eng.istate[4]) == 1 and abs(float(eng.fstate[0]) - scaler.scale) < 1e-6
    # the bf16 working copy follows the master
    s = eng.by_name["conv21/kernel"]
    K, R, C = s["shape"]
    wb = eng.wb[s["half_offset"]:s["half_offset"] + s["size"]].view(K, R, C).float()
    m = eng.param_view("conv21/kernel")
    assert torch.equal(wb, m.bfloat16().float())


@pytest.mark.parametrize("algo,policy", [("adam", "cosine_decay"), ("momentum", "exp_decay"), ("novograd", "fixed_lr")])
def test_optimizer_variants_vs_oracle(algo, policy):
    """Adam / Momentum / NovoGrad with the cosine / exponential / fixed policies, an L2 regulariser on
    kernels and BN gammas (added to the unscaled gradient before LARC) and a static loss scale."""
    from oracle import optimizer as OO
    from openseq2seq_b200.engine import JasperEngine
    from tests.common_cfg import MINI_JASPER
    opt_kw = {"adam": dict(beta1=0.9, beta2=0.999, epsilon=1e-8),
              "momentum": dict(momentum=0.9),
              "novograd": dict(beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.001)}[algo]
    pol_kw = {"cosine_decay": dict(decay_steps=8, min_lr=0.1, warmup_steps=3),
              "exp_decay": dict(decay_steps=2, decay_rate=0.5, use_staircase_decay=True, begin_decay_at=1, min_lr=1e-4),
              "fixed_lr": dict()}[policy]
    lr0, reg = 0.01, 1e-3
    eng = JasperEngine(MINI_JASPER, 64, 29, world_size=1,
                       opt=dict(algo=algo, larc_eta=0.001, learning_rate=lr0, loss_scaling=False, initial_scale=128.0,
                                lr_policy=policy, l2_regularizer_scale=reg, **opt_kw, **pol_kw))
    names = [n for n, _ in eng.named_parameters()]
    kinds = [eng.by_name[n]["kind"] for n in names]
    # non-zero betas / bias: with warm-up the first step has lr = 0 and LARC's ratio for an all-zero
    # variable would be 0/0 (the reference would skip that step on the NaN; not what is under test here)
    for n, k in zip(names, kinds):
        if k in ("beta", "fc_b"):
            eng.param_view(n).fill_(0.05)
    # conv / dense kernels and BN gammas carry the regularizer; the 1x1 residual kernels are built without one
    # (conv_blocks.py:80-86)
    reg_scales = [reg if (k in ("conv", "gamma", "fc_w") and not (k == "conv" and "/res" in n)) else 0.0
                  for n, k in zip(names, kinds)]
    assert reg_scales == [float(np.float32(v)) if v else 0.0 for v in eng._reg.tolist()] or \
        np.allclose(reg_scales, eng._reg.cpu().numpy())
    w_ref = [eng.param_view(n).detach().cpu().numpy().copy() for n in names]
    state = OO.AdamState(len(names)) if algo == "adam" else OO.NovoGradState(len(names))

    class _Static(object):
        scale = 128.0

        def update(self, has_nan, amax):
            return bool(has_nan) or bool(np.isinf(amax))

    lr_fn = {"cosine_decay": lambda s: OO.cosine_decay(s, lr0, 8, min_lr=0.1, warmup_steps=3),
             "exp_decay": lambda s: OO.exp_decay(s, lr0, 2, 0.5, True, begin_decay_at=1, min_lr=1e-4),
             "fixed_lr": lambda s: OO.fixed_lr(s, lr0)}[policy]
    rng = np.random.default_rng(1)
    step = 0
    for it in range(6):
        per_rank = [[(rng.standard_normal(w.shape) * 0.01 * 128.0).astype(np.float32) for w in w_ref]]
        eng.grad.zero_()
        for i, n in enumerate(names):
            eng.param_view(n, eng.grad).copy_(torch.tensor(per_rank[0][i]))
        eng.optimizer_step()
        torch.cuda.synchronize()
        skipped, lr, step = OO.train_step(w_ref, per_rank, state, _Static(), step, lr_fn, opt_kw,
                                          larc_params=dict(larc_eta=0.001), algo=algo, reg_scales=reg_scales)
        assert not skipped and int(eng.istate[2]) == step
        assert abs(float(eng.fstate[1]) - lr) < 1e-6 * max(1.0, lr), (it, float(eng.fstate[1]), lr)
        for i, n in enumerate(names):
            got = eng.param_view(n).cpu().numpy()
            assert np.abs(got - w_ref[i]).max() <= 3e-5 * max(1.0, np.abs(w_ref[i]).max()), (it, n)


def test_logmel_featurizer_vs_oracle():
    from oracle import featurizer as FZ
    L, lib = _lib()
    rng = np.random.default_rng(1234)
    sigs = [np.clip(3000 * rng.standard_normal(n), -32768, 32767).astype(np.int16) for n in (16000, 12345, 4000)]
    # speech-like spectral tilt so the log-mel features are not flat noise
    sigs = [np.clip(np.convolve(s.astype(np.float64), np.ones(8) / 8, mode="same"), -32768, 32767).astype(np.int16)
            for s in sigs]
    ref, ref_lens = FZ.batch_features(sigs, pad_to=16)
    B, T_pad, F = ref.shape
    wave = torch.tensor(np.concatenate(sigs), dtype=torch.int16, device="cuda")
    offs = torch.tensor(np.cumsum([0] + [len(s) for s in sigs[:-1]]), dtype=torch.int64, device="cuda")
    ns = torch.tensor([len(s) for s in sigs], dtype=torch.int32, device="cuda")
    mel = torch.tensor(FZ.mel_filterbank(), dtype=torch.float32, device="cuda")
    win = torch.tensor(np.hanning(320), dtype=torch.float32, device="cuda")
    absmax = torch.zeros(B, dtype=torch.int32, device="cuda")
    raw = torch.zeros(B * T_pad * F, device="cuda")
    out = torch.zeros(B, T_pad, F, device="cuda")
    outb = torch.zeros(B, T_pad, F, dtype=torch.bfloat16, device="cuda")
    lens = torch.zeros(B, dtype=torch.int32, device="cuda")
    melnp = FZ.mel_filterbank()
    band = torch.tensor([[int(np.nonzero(r)[0].min()), int(np.nonzero(r)[0].max()) + 1] for r in melnp],
                        dtype=torch.int32, device="cuda")
    L.check(lib.os2s_logmel_forward(L.ptr(wave), L.ptr(offs), L.ptr(ns), B, L.ptr(mel), L.ptr(band), L.ptr(win), 512, 320, 160, F,
                                    T_pad, max(len(s) for s in sigs), _f(0.0), ctypes.c_uint64(0), _f(0.97),
                                    L.ptr(absmax), L.ptr(raw), L.ptr(outb), L.ptr(out), L.ptr(lens), L.stream_ptr()),
            "logmel")
    torch.cuda.synchronize()
    assert lens.cpu().tolist() == ref_lens.tolist()
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() < 2e-2  # normalised features are O(1)
    assert np.abs(outb.float().cpu().numpy() - ref).max() < 5e-2
    for b in range(B):
        assert np.all(got[b, ref_lens[b]:] == 0)


@pytest.mark.parametrize("mode", ["psf", "librosa_global_norm"])
def test_featurizer_psf_backend_and_global_normalisation_vs_oracle(mode):
    """os2s_features_forward: python_speech_features conventions (int16 re-quantisation, frame count a
    multiple of pad_to with the padding pre-emphasised, rectangular frames, power / n_fft, HTK filterbank,
    one mean/std per utterance) and the librosa backend with norm_per_feature=False."""
    from oracle import featurizer as FZ
    L, lib = _lib()
    rng = np.random.default_rng(99)
    sigs = [np.clip(3000 * rng.standard_normal(n), -32768, 32767).astype(np.int16) for n in (16000, 12345, 4001, 2000)]
    sigs = [np.clip(np.convolve(s.astype(np.float64), np.ones(8) / 8, mode="same"), -32768, 32767).astype(np.int16)
            for s in sigs]
    F = 64
    if mode == "psf":
        feats = [FZ.psf_logfbank_features(s, num_features=F, pad_to=8)[0] for s in sigs]
        melnp, win = FZ.psf_mel_filterbank(F, 512, 16000, 0.0, 8000.0), np.ones(320)
        psf, per_feature = 1, 0
    else:
        feats = [FZ.logfbank_features(s, num_features=F, norm_per_feature=False)[0] for s in sigs]
        melnp, win = FZ.mel_filterbank(), np.hanning(320)
        psf, per_feature = 0, 0
    ref_lens = [f.shape[0] for f in feats]
    if mode == "psf":
        assert all(n % 8 == 0 for n in ref_lens)
        assert ref_lens[0] == 104   # 1 + ceil((16000 - 320) / 160) = 99 -> 104
    B, T_pad = len(sigs), -(-max(ref_lens) // 16) * 16
    wave = torch.tensor(np.concatenate(sigs), dtype=torch.int16, device="cuda")
    offs = torch.tensor(np.cumsum([0] + [len(s) for s in sigs[:-1]]), dtype=torch.int64, device="cuda")
    ns = torch.tensor([len(s) for s in sigs], dtype=torch.int32, device="cuda")
    mel = torch.tensor(melnp, dtype=torch.float32, device="cuda")
    band = torch.tensor([[int(np.nonzero(r)[0].min()), int(np.nonzero(r)[0].max()) + 1] if np.any(r) else [0, 0]
                         for r in melnp], dtype=torch.int32, device="cuda")
    wint = torch.tensor(win, dtype=torch.float32, device="cuda")
    absmax = torch.zeros(B, dtype=torch.int32, device="cuda")
    raw = torch.zeros(B * T_pad * F, device="cuda")
    out = torch.full((B, T_pad, F), float("nan"), device="cuda")
    lens = torch.zeros(B, dtype=torch.int32, device="cuda")
    L.check(lib.os2s_features_forward(L.ptr(wave), L.ptr(offs), L.ptr(ns), B, L.ptr(mel), L.ptr(band), L.ptr(wint), 512, 320,
                                      160, F, T_pad, max(len(s) for s in sigs), _f(0.0), ctypes.c_uint64(0), _f(0.97),
                                      psf, 8, per_feature, L.ptr(absmax), L.ptr(raw), None, L.ptr(out), L.ptr(lens),
                                      L.stream_ptr()), "features")
    torch.cuda.synchronize()
    assert lens.cpu().tolist() == ref_lens
    got = out.cpu().numpy()
    for b in range(B):
        n = ref_lens[b]
        assert np.abs(got[b, :n] - feats[b]).max() < 2e-2, (mode, b, np.abs(got[b, :n] - feats[b]).max())
        assert np.all(got[b, n:] == 0)
        # the reference's own pin for this backend (speech_utils_test.py:72-73): mean 0, std 1
        assert abs(float(got[b, :n].mean())) < 1e-4 and abs(float(got[b, :n].std()) - 1.0) < 1e-4


def test_speed_perturbation_and_noise_vs_oracle(golden_dir):
    """os2s_augment_signal (speech_utils.py:245-266): resampy 'kaiser_best' band-limited interpolation at the
    ratios of the Jasper recipe (0.9, 1.0, 1.1) on a real toy-speech wav vs the oracle's restatement, and the
    additive noise at the drawn level (statistics)."""
    import scipy.io.wavfile as wavfile
    from oracle import augment as AU
    from open_seq2seq.data.speech2text import speech_utils as SU
    L, lib = _lib()
    sr, sig = wavfile.read(os.path.join(golden_dir, "toy_speech_data", "wav_files", "46gc040q.wav"))
    sig = sig.astype(np.int16)[:40000]
    sigs = [sig, sig[:30001], sig[5000:33333]]
    sr_new = np.array([14400, 0, 17600], dtype=np.int32)
    n_out = np.array([AU.resample_out_len(len(s), sr, r) if r else len(s) for s, r in zip(sigs, sr_new)], dtype=np.int32)
    B = len(sigs)
    dev = "cuda"
    wave = torch.tensor(np.concatenate(sigs), dtype=torch.int16, device=dev)
    offs = torch.tensor(np.cumsum([0] + [len(s) for s in sigs[:-1]]), dtype=torch.int64, device=dev)
    ns = torch.tensor([len(s) for s in sigs], dtype=torch.int32, device=dev)
    ooffs = torch.tensor(np.cumsum([0] + list(n_out[:-1])), dtype=torch.int64, device=dev)
    tab, num_table = SU.kaiser_best_table()
    assert np.array_equal(tab, AU.kaiser_best_table()[0])   # product table == oracle table
    tabd = torch.tensor(tab, dtype=torch.float32, device=dev)
    absmax = torch.zeros(B, dtype=torch.int32, device=dev)
    out = torch.full((int(n_out.sum()),), float("nan"), device=dev)
    srd, nod = torch.tensor(sr_new, device=dev), torch.tensor(n_out, device=dev)
    st = L.stream_ptr()
    L.check(lib.os2s_wave_absmax(L.ptr(wave), L.ptr(offs), L.ptr(ns), B, L.ptr(absmax), st), "absmax")
    L.check(lib.os2s_augment_signal(L.ptr(wave), L.ptr(offs), L.ptr(ns), B, L.ptr(absmax), _f(0.0),
                                    L.ptr(srd), sr, L.ptr(tabd), tabd.numel(), num_table,
                                    None, ctypes.c_uint64(1), L.ptr(out), L.ptr(ooffs),
                                    L.ptr(nod), int(n_out.max()), st), "augment")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert absmax.cpu().tolist() == [int(np.abs(s.astype(np.int32)).max()) for s in sigs]
    o = 0
    for b, s in enumerate(sigs):
        x = s.astype(np.float32) * (1.0 / (np.abs(s.astype(np.float32)).max() + 1e-5))
        ref = AU.resample(x, sr, int(sr_new[b])) if sr_new[b] else x.astype(np.float64)
        assert len(ref) == n_out[b]
        err = np.abs(got[o:o + n_out[b]] - ref)
        # fp32 accumulation of ~130 taps of an O(1) signal.  When downsampling, resampy's table step is
        # int(0.9 * 512) = 460 while the fraction spans 460.8 entries: at output samples whose time register is
        # an exact integer (every 9th) the result depends on the rounding of the ACCUMULATED register (numpy's
        # cumsum here, repeated addition in resampy, t * increment on the device) -- isolated 1e-3 differences
        assert np.percentile(err, 85) < 2e-5 and err.max() < 5e-3, (b, float(np.percentile(err, 85)), float(err.max()))
        o += n_out[b]
    # noise: out - clean has the drawn amplitude, zero mean, unit-variance Gaussian shape
    amp = torch.tensor([0.01, 0.0, 0.002], device=dev)
    noisy = torch.empty_like(out)
    L.check(lib.os2s_augment_signal(L.ptr(wave), L.ptr(offs), L.ptr(ns), B, L.ptr(absmax), _f(0.0),
                                    L.ptr(srd), sr, L.ptr(tabd), tabd.numel(), num_table,
                                    L.ptr(amp), ctypes.c_uint64(7), L.ptr(noisy), L.ptr(ooffs),
                                    L.ptr(nod), int(n_out.max()), st), "augment+noise")
    torch.cuda.synchronize()
    d = (noisy - out).cpu().numpy()
    o = 0
    for b in range(B):
        seg = d[o:o + n_out[b]]
        if float(amp[b]) == 0:
            assert np.all(seg == 0)
        else:
            z = seg / float(amp[b])
            assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
            assert abs(np.mean(z ** 4) - 3.0) < 0.3      # Gaussian kurtosis
        o += n_out[b]


def test_featurizer_on_augmented_signal_with_spec_masks_gain_and_fixed_normalisation():
    """os2s_features_forward_p: (a) features of the speed-perturbed float signal + spec-augment masks vs the
    oracle (speech_utils.py:354-433); (b) params['gain'] with features_mean / features_std_dev instead of
    the computed statistics; (c) fp16 output format."""
    from oracle import augment as AU
    from oracle import featurizer as FZ
    L, lib = _lib()
    rng = np.random.default_rng(5)
    sigs = [np.clip(3000 * rng.standard_normal(n), -32768, 32767).astype(np.int16) for n in (24000, 17777)]
    sigs = [np.clip(np.convolve(s.astype(np.float64), np.ones(8) / 8, mode="same"), -32768, 32767).astype(np.int16)
            for s in sigs]
    sr, F, hop = 16000, 64, 160
    sr_new = [17600, 14400]
    xs = [s.astype(np.float32) * (1.0 / (np.abs(s.astype(np.float32)).max() + 1e-5)) for s in sigs]
    aug_sig = [AU.resample(x, sr, r) for x, r in zip(xs, sr_new)]
    r = np.random.RandomState(11)
    augp = {"n_freq_mask": 2, "n_time_mask": 2, "width_freq_mask": 6, "width_time_mask": 6}
    masks_l = [AU.draw_spec_masks(1 + len(a) // hop, F, augp, r) for a in aug_sig]
    # the normalised features do not depend on the signal's scale: the oracle featurizer may renormalise
    ref = [AU.apply_spec_masks(FZ.logfbank_features(a, dither=0.0)[0], m) for a, m in zip(aug_sig, masks_l)]
    B = 2
    ref_lens = [f.shape[0] for f in ref]
    T_pad = -(-max(ref_lens) // 16) * 16
    dev = "cuda"
    n_out = np.array([len(a) for a in aug_sig], dtype=np.int32)
    sig = torch.tensor(np.concatenate(aug_sig), dtype=torch.float32, device=dev)
    soff = torch.tensor([0, n_out[0]], dtype=torch.int64, device=dev)
    nsd = torch.tensor(n_out, device=dev)
    melnp = FZ.mel_filterbank()
    mel = torch.tensor(melnp, dtype=torch.float32, device=dev)
    band = torch.tensor([[int(np.nonzero(q)[0].min()), int(np.nonzero(q)[0].max()) + 1] for q in melnp],
                        dtype=torch.int32, device=dev)
    win = torch.tensor(np.hanning(320), dtype=torch.float32, device=dev)
    nm = 4
    mk = np.zeros((B, nm, 3), dtype=np.int32)
    for b, ml in enumerate(masks_l):
        for i, m in enumerate(ml):
            mk[b, i] = m
    mkd = torch.tensor(mk, device=dev)
    absmax = torch.zeros(B, dtype=torch.int32, device=dev)
    raw = torch.zeros(B * T_pad * F, device=dev)
    out = torch.full((B, T_pad, F), float("nan"), device=dev)
    out16 = torch.zeros(B, T_pad, F, dtype=torch.float16, device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    L.check(lib.os2s_features_forward_p(None, L.ptr(sig), L.ptr(soff), L.ptr(soff), L.ptr(nsd), B, L.ptr(mel), L.ptr(band),
                                        L.ptr(win), 512, 320, hop, F, T_pad, int(n_out.max()), _f(0.0), ctypes.c_uint64(0),
                                        _f(0.97), 0, 16, 1, _f(0.0), None, None, L.ptr(mkd), nm, 0, None, 0, L.ptr(absmax), L.ptr(raw),
                                        L.ptr(out16), L.ptr(out), L.ptr(lens), 1, L.stream_ptr()), "features_p")
    torch.cuda.synchronize()
    assert lens.cpu().tolist() == ref_lens
    got = out.cpu().numpy()
    for b in range(B):
        n = ref_lens[b]
        assert np.abs(got[b, :n] - ref[b]).max() < 2e-2, (b, np.abs(got[b, :n] - ref[b]).max())
        assert np.all(got[b, n:] == 0)
        for kind, base, width in masks_l[b]:     # the bands are exact zeros
            blk = got[b, :n, base:base + width] if kind == 0 else got[b, base:base + width, :]
            assert np.all(blk == 0)
    assert np.abs(out16.float().cpu().numpy() - got).max() < 4e-3      # fp16 rounding of O(1) values
    # (b) fixed gain + given mean / std: (log-mel - mean) / std of the un-normalised features
    gain = 1.0 / 20000.0
    s0 = sigs[0]
    wave = torch.tensor(s0, dtype=torch.int16, device=dev)
    off0 = torch.zeros(1, dtype=torch.int64, device=dev)
    n0 = torch.tensor([len(s0)], dtype=torch.int32, device=dev)
    x = FZ.preemphasis(s0.astype(np.float32) * np.float32(gain))
    logmel = np.log(melnp @ FZ.stft_power(x) + 1e-20).T
    fm = rng.standard_normal(F).astype(np.float32)
    fs = (1.0 + rng.random(F)).astype(np.float32)
    want = (logmel - fm) / fs
    T1 = -(-logmel.shape[0] // 16) * 16
    out1 = torch.full((1, T1, F), float("nan"), device=dev)
    raw1 = torch.zeros(T1 * F, device=dev)
    fmd, fsd = torch.tensor(fm, device=dev), torch.tensor(fs, device=dev)
    L.check(lib.os2s_features_forward_p(L.ptr(wave), None, None, L.ptr(off0), L.ptr(n0), 1, L.ptr(mel), L.ptr(band), L.ptr(win),
                                        512, 320, hop, F, T1, len(s0), _f(0.0), ctypes.c_uint64(0), _f(0.97), 0, 16, 1,
                                        _f(gain), L.ptr(fmd), L.ptr(fsd),
                                        None, 0, 0, None, 0, L.ptr(absmax), L.ptr(raw1), None, L.ptr(out1), L.ptr(lens), 0,
                                        L.stream_ptr()), "features_p fixed")
    torch.cuda.synchronize()
    g1 = out1.cpu().numpy()[0, :logmel.shape[0]]
    assert np.abs(g1 - want).max() < 2e-2 * max(1.0, np.abs(want).max() / 10)


@pytest.mark.parametrize("ftype", ["spectrogram", "mfcc"])
def test_psf_spectrogram_and_mfcc_feature_types_vs_oracle(ftype):
    """get_speech_features_psf(features_type='spectrogram' | 'mfcc') (speech_utils.py:490-512) on the GPU vs the
    oracle restatement (pinned on the reference's shape / mean 0 / std 1 test in tests/test_oracle.py)."""
    from oracle import featurizer as FZ
    from open_seq2seq.data.speech2text import speech_utils as SU
    L, lib = _lib()
    rng = np.random.default_rng(7)
    sigs = [np.clip(3000 * rng.standard_normal(n), -32768, 32767).astype(np.int16) for n in (16000, 9999, 4001)]
    sigs = [np.clip(np.convolve(s.astype(np.float64), np.ones(8) / 8, mode="same"), -32768, 32767).astype(np.int16)
            for s in sigs]
    if ftype == "spectrogram":
        F, code = 161, 1
        feats = [FZ.psf_spectrogram_features(s, num_features=F, pad_to=8)[0] for s in sigs]
        melnp, win, post, n_filt = np.zeros((1, 257)), np.hanning(320), None, 0
    else:
        F, code = 40, 2
        feats = [FZ.psf_mfcc_features(s, num_features=F, pad_to=8)[0] for s in sigs]
        melnp, win = FZ.psf_mel_filterbank(2 * F, 512, 16000, 0.0, 8000.0), np.ones(320)
        post, n_filt = SU.psf_mfcc_matrix(F, 2 * F, 2 * F), 2 * F
    ref_lens = [f.shape[0] for f in feats]
    B, T_pad = len(sigs), -(-max(ref_lens) // 8) * 8
    dev = "cuda"
    wave = torch.tensor(np.concatenate(sigs), dtype=torch.int16, device=dev)
    offs = torch.tensor(np.cumsum([0] + [len(s) for s in sigs[:-1]]), dtype=torch.int64, device=dev)
    ns = torch.tensor([len(s) for s in sigs], dtype=torch.int32, device=dev)
    mel = torch.tensor(melnp, dtype=torch.float32, device=dev)
    band = None
    if ftype == "mfcc":
        band = torch.tensor([[int(np.nonzero(r)[0].min()), int(np.nonzero(r)[0].max()) + 1] if np.any(r) else [0, 0]
                             for r in melnp], dtype=torch.int32, device=dev)
    wint = torch.tensor(win, dtype=torch.float32, device=dev)
    postd = torch.tensor(post, dtype=torch.float32, device=dev) if post is not None else None
    absmax = torch.zeros(B, dtype=torch.int32, device=dev)
    raw = torch.zeros(B * T_pad * F, device=dev)
    out = torch.full((B, T_pad, F), float("nan"), device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    L.check(lib.os2s_features_forward_p(L.ptr(wave), None, None, L.ptr(offs), L.ptr(ns), B, L.ptr(mel), L.ptr(band), L.ptr(wint),
                                        512, 320, 160, F, T_pad, max(len(s) for s in sigs), _f(0.0), ctypes.c_uint64(0),
                                        _f(0.97), 1, 8, 0, _f(0.0), None, None, None, 0, code, L.ptr(postd), n_filt,
                                        L.ptr(absmax), L.ptr(raw), None, L.ptr(out), L.ptr(lens), 0, L.stream_ptr()), ftype)
    torch.cuda.synchronize()
    assert lens.cpu().tolist() == ref_lens
    got = out.cpu().numpy()
    for b in range(B):
        n = ref_lens[b]
        err = np.abs(got[b, :n] - feats[b]).max()
        assert err < 3e-2, (ftype, b, err)
        assert np.all(got[b, n:] == 0)
        assert abs(float(got[b, :n].mean())) < 1e-3 and abs(float(got[b, :n].std()) - 1.0) < 1e-3
