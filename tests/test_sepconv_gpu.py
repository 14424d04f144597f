"""GPU parity of the separable-convolution path (sep_conv1d: QuartzNet 15x5, Jasper-Mini): the CUDA-core
depthwise kernels and the composed-kernel helpers against the oracle's tf.layers.separable_conv1d restatement
(oracle/torch_twin.py::sep_conv1d_same, reference parts/cnns/conv_blocks.py:27-40,79-85,180-193), and a
QuartzNet-style stack through JasperEngine: logits to 1e-2 (L2), every parameter gradient to 2e-2 at the
engine's own forward state, a few training steps."""
import ctypes

import numpy as np
import pytest
import torch

from tests.common_cfg import MINI_QUARTZ

pytestmark = pytest.mark.gpu
_f = ctypes.c_float


def _lib():
    from openseq2seq_b200 import _lib as L
    return L, L.load()


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,T,C,K,dil,half", [
    (2, 300, 256, 33, 1, "bf16"),
    (3, 131, 128, 11, 1, "fp16"),     # ragged rows
    (2, 752, 512, 87, 2, "bf16"),     # QuartzNet's widest span: 87 taps, dilation 2
    (1, 64, 64, 1, 1, "bf16"),
])
def test_depthwise_conv_fwd_dgrad_wgrad_vs_oracle(B, T, C, K, dil, half):
    import torch.nn.functional as F
    from oracle import torch_twin as TT
    L, lib = _lib()
    hdt = torch.float16 if half == "fp16" else torch.bfloat16
    flags = 1 if half == "fp16" else 0
    g = torch.Generator().manual_seed(K + C)
    x = torch.randn(B, T, C, generator=g).to(hdt)
    D = (torch.randn(K, C, 1, generator=g) / K ** 0.5).float()
    dz = torch.randn(B, T, C, generator=g).to(hdt)
    pad = ((K - 1) * dil) // 2
    # oracle: separable conv with an identity pointwise stage == the depthwise stage; gradients by autograd
    xr = x.double().requires_grad_(True)
    Dr = D.double().requires_grad_(True)
    eye = torch.eye(C, dtype=torch.float64)[None]
    z_ref = TT.sep_conv1d_same(xr, Dr, eye, 1, dil)
    z_ref.backward(dz.double())
    xd, dzd, Dd = x.cuda(), dz.cuda(), D.reshape(K, C).contiguous().cuda()
    st = L.stream_ptr()
    z = torch.full((B, T, C), float("nan"), dtype=torch.float32, device="cuda")
    L.check(lib.os2s_depthwise_conv1d(L.ptr(xd), L.ptr(Dd), L.ptr(z), B, T, C, K, -pad, dil, 1, flags, st), "dw fwd")
    z16 = torch.zeros(B, T, C, dtype=hdt, device="cuda")
    L.check(lib.os2s_depthwise_conv1d(L.ptr(xd), L.ptr(Dd), L.ptr(z16), B, T, C, K, -pad, dil, 4 if flags else 0, flags, st),
            "dw fwd16")
    dx = torch.ones(B, T, C, dtype=torch.float32, device="cuda")      # accumulate mode: + 1
    L.check(lib.os2s_depthwise_conv1d(L.ptr(dzd), L.ptr(Dd), L.ptr(dx), B, T, C, K, pad, -dil, 2, flags, st), "dw dgrad")
    dD = torch.full((K, C), float("nan"), device="cuda")
    L.check(lib.os2s_depthwise_conv1d_wgrad(L.ptr(xd), L.ptr(dzd), L.ptr(dD), B, T, C, K, dil, pad, flags, st), "dw wgrad")
    torch.cuda.synchronize()
    zr = z_ref.detach()
    assert float((z.cpu().double() - zr).abs().max()) <= 1e-5 * float(zr.abs().max()) + 1e-5
    assert float((z16.cpu().double() - zr).abs().max()) <= (2e-3 if flags else 1e-2) * float(zr.abs().max())
    assert float((dx.cpu().double() - 1.0 - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max()) + 1e-5
    gref = Dr.grad.reshape(K, C)
    assert float((dD.cpu().double() - gref).abs().max()) <= 1e-4 * float(gref.abs().max()) + 1e-4


def test_composed_separable_kernel_and_gradient_fold_back():
    L, lib = _lib()
    K, C, Co = 11, 64, 256
    g = torch.Generator().manual_seed(1)
    D = torch.randn(K, C, generator=g).cuda()
    P = torch.randn(C, Co, generator=g).cuda()
    dW = torch.randn(K, C, Co, generator=g).cuda()
    w16 = torch.zeros(K, C, Co, dtype=torch.float16, device="cuda")
    dD = torch.zeros(K, C, device="cuda")
    dP = torch.zeros(C, Co, device="cuda")
    st = L.stream_ptr()
    L.check(lib.os2s_sepconv_compose(L.ptr(D), L.ptr(P), L.ptr(w16), K, C, Co, 1, st), "compose")
    L.check(lib.os2s_sepconv_decompose_grad(L.ptr(dW), L.ptr(D), L.ptr(P), L.ptr(dD), L.ptr(dP), K, C, Co, st), "decompose")
    torch.cuda.synchronize()
    W = D[:, :, None] * P[None]
    assert torch.equal(w16, W.half())
    assert torch.allclose(dD, (dW * P[None]).sum(2), rtol=1e-4, atol=1e-3)
    assert torch.allclose(dP, (dW * D[:, :, None]).sum(0), rtol=1e-4, atol=1e-3)


def _setup(half="bf16", B=3, T=160, F=64, V=29):
    from openseq2seq_b200.engine import JasperEngine
    from oracle import torch_twin as TT
    torch.manual_seed(0)
    lens = torch.tensor([T, T - 22, T - 57][:B], dtype=torch.int32)
    feats = (torch.randn(B, T, F) * TT.sequence_mask(lens.long(), T, torch.float32)).bfloat16().float()
    params = TT.init_params(MINI_QUARTZ, F, V, seed=3)
    gen = torch.Generator().manual_seed(11)
    for k in params:
        if k.endswith("/gamma"):
            params[k] = 1.0 + 0.2 * torch.randn(params[k].shape, generator=gen)
        if k.endswith("/beta"):
            params[k] = 0.1 * torch.randn(params[k].shape, generator=gen)
    eng = JasperEngine(MINI_QUARTZ, F, V, training=True, dropout_keep_default=1.0, act_dtype=half,
                       opt=dict(loss_scaling=False, learning_rate=0.01))
    eng.load_parameters(params)
    return eng, params, feats, lens


def _device_weights(eng, params):
    """The oracle multiplies with what the device multiplies with: 16-bit roundings of the pointwise / dense
    kernels and of the composed products D * P; depthwise taps of split layers stay fp32."""
    hdt = eng.act_torch
    out = {k: v.clone() for k, v in params.items()}
    for s in eng.specs:
        n = s["name"]
        if s["kind"] == "conv" and not n.endswith("@composed"):
            out[n] = params[n].to(hdt).float()
    return out


@pytest.mark.parametrize("half", ["bf16", "fp16"])
def test_quartznet_style_stack_forward_and_backward_vs_oracle(half):
    from oracle import torch_twin as TT
    eng, params, feats, lens = _setup(half)
    modes = [(l.sep, l.sep_mode) for l in eng.layers]
    assert modes[0] == (True, "compose") and (True, "split") in modes and modes[-2] == (True, "compose") and modes[-1] == (False, None)
    logits, out_lens = eng.forward(feats.cuda(), lens.cuda())
    torch.cuda.synchronize()
    pw = {k: v.double() for k, v in _device_weights(eng, params).items()}
    with torch.no_grad():
        enc, ref_len = TT.tdnn_encode(feats.double(), lens.long(), MINI_QUARTZ, pw)
        ref = TT.fc_decode(enc, pw["fc/kernel"], pw["fc/bias"]).transpose(0, 1)
    assert out_lens.cpu().tolist() == ref_len.tolist()
    for b in range(feats.shape[0]):
        n = int(ref_len[b])
        e = _rel_l2(logits[b, :n], ref[b, :n])
        assert e < (1e-2 if half == "fp16" else 2e-2), (half, b, e)   # composed kernels round D*P once more in bf16
    # backward at the engine's own forward state
    g = torch.Generator().manual_seed(9)
    R = torch.randn(logits.shape, generator=g)
    R = R * TT.sequence_mask(out_lens.cpu().long(), logits.shape[1], R.dtype)
    eng.backward_from_dlogits(R.cuda())
    torch.cuda.synchronize()
    ws = eng._last_ws
    conv, out = {}, {}
    for li, l in enumerate(eng.layers):
        conv[l.name] = ws.Y[li].float().cpu()
        out[l.name] = ws.A[li].float().cpu()
        for n in range(len(l.res_sources)):
            j, col = eng.res_col[(li, n)]
            conv[eng.res_name(l, n)] = ws.YRcat[j][:, :, col:col + l.c_out].float().cpu()
    refg = TT.backward_with_saved_forward(_device_weights(eng, params), MINI_QUARTZ, feats, lens.long(), conv, out, R)
    names = [n for n, _ in eng.named_parameters()]
    assert any(n.endswith("/depthwise_kernel") for n in names) and not any("@composed" in n for n in names)
    worst = {n: _rel_l2(eng.param_view(n, eng.grad), refg[n]) for n in names}
    bad = {k: round(v, 4) for k, v in worst.items() if v > 3e-2}
    assert not bad, "gradient mismatch: %r" % bad


def test_quartznet_style_stack_trains():
    eng, params, feats, lens = _setup("bf16")
    eng.set_optimizer(algo="novograd", beta1=0.95, beta2=0.5, weight_decay=0.001, learning_rate=0.01,
                      lr_policy="cosine_decay", decay_steps=200, loss_scaling=True)
    eng.load_parameters(params)
    B = feats.shape[0]
    gl = torch.Generator().manual_seed(5)
    labels = torch.randint(0, 28, (B, 12), generator=gl, dtype=torch.int32)
    label_lens = torch.tensor([12, 9, 7][:B], dtype=torch.int32)
    x = feats.cuda()
    losses = [float(eng.train_step(x, lens.cuda(), labels.cuda(), label_lens.cuda()).mean()) for _ in range(40)]
    assert np.isfinite(losses).all() and losses[-1] < 0.7 * losses[0], losses
    # the composed pseudo-kernels follow their factors and are never updated themselves
    for s in eng.specs:
        if s["name"].endswith("@composed"):
            assert float(eng.param_view(s["name"]).abs().max()) == 0.0
