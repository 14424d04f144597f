"""GPU parity of the whole training path (JasperEngine through the C ABI) against the oracle's
torch-CPU fp32 twin on the same seeded inputs: encoder logits to 1e-2 relative (north-star
tolerance), identical greedy-CTC tokens, loss and parameter gradients."""
import numpy as np
import pytest
import torch

from tests.common_cfg import MINI_JASPER

pytestmark = pytest.mark.gpu


def _rel(a, b):
    """max-norm relative error."""
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _rel_l2(a, b):
    """relative error in the L2 sense: ||a - b|| / ||b||."""
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _setup(B=3, T=96, F=64, V=29, seed=0):
    from openseq2seq_b200.engine import JasperEngine
    from oracle import torch_twin as TT
    torch.manual_seed(seed)
    lens = torch.tensor([T, T - 22, T - 41][:B], dtype=torch.int32)
    feats = torch.randn(B, T, F)
    feats = feats * TT.sequence_mask(lens.long(), T, feats.dtype)
    feats = feats.bfloat16().float()  # both sides see the same (bf16-representable) features
    params = TT.init_params(MINI_JASPER, F, V, seed=3)
    # bf16-representable conv kernels so the comparison isolates kernel arithmetic
    for k in params:
        if k.endswith("/kernel") and k != "fc/kernel":
            params[k] = params[k].bfloat16().float()
    # non-trivial BN affine parameters
    g = torch.Generator().manual_seed(11)
    for k in params:
        if k.endswith("/gamma"):
            params[k] = 1.0 + 0.2 * torch.randn(params[k].shape, generator=g)
        if k.endswith("/beta"):
            params[k] = 0.1 * torch.randn(params[k].shape, generator=g)
    eng = JasperEngine(MINI_JASPER, F, V, training=True, dropout_keep_default=1.0,
                       opt=dict(loss_scaling=False, learning_rate=0.01))
    for l in eng.layers:
        l.keep = 1.0
    eng._ws = {}
    eng.load_parameters(params)
    L = 12
    gl = torch.Generator().manual_seed(5)
    labels = torch.randint(0, V - 1, (B, L), generator=gl, dtype=torch.int32)
    label_lens = torch.tensor([12, 9, 7][:B], dtype=torch.int32)
    return eng, params, feats, lens, labels, label_lens


def test_forward_logits_and_greedy_match_oracle():
    """(a) against the fp64 oracle: north-star tolerance 1e-2 relative on the logits, identical greedy
    tokens; (b) against the oracle with the device's storage rounding emulated (fp16 conv outputs,
    bf16 activations): only accumulation-order differences remain, so the bound is 10x tighter."""
    from oracle import torch_twin as TT
    from oracle import ctc as OC
    eng, params, feats, lens, labels, label_lens = _setup()
    logits, out_lens = eng.forward(feats.cuda().bfloat16().contiguous(), lens.cuda())
    torch.cuda.synchronize()
    p64 = {k: v.double() for k, v in params.items()}
    _, ref_logits, ref_len = TT.forward_loss(p64, MINI_JASPER, feats.double(), lens.long(), labels.long(),
                                             label_lens.long())
    _, emu_logits, _ = TT.forward_loss(p64, MINI_JASPER, feats.double(), lens.long(), labels.long(),
                                       label_lens.long(), emulate_storage=True)
    ref = ref_logits.transpose(0, 1)  # [B,T,V]
    emu = emu_logits.transpose(0, 1)
    assert out_lens.cpu().tolist() == ref_len.tolist()
    for b in range(feats.shape[0]):
        n = int(ref_len[b])  # padded frames are defined but irrelevant to loss / decode
        assert _rel_l2(logits[b, :n], ref[b, :n]) < 1e-2
        assert _rel(logits[b, :n], ref[b, :n]) < 1.5e-2
        assert _rel(logits[b, :n], emu[b, :n]) < 3e-3
    toks, tl = eng.greedy_decode()
    torch.cuda.synchronize()
    ref_toks, _ = OC.ctc_greedy_decode(ref_logits.detach().numpy(), ref_len.numpy())
    for b in range(feats.shape[0]):
        assert toks[b, :int(tl[b])].cpu().tolist() == ref_toks[b]


def test_ctc_loss_value_matches_oracle():
    from oracle import torch_twin as TT
    eng, params, feats, lens, labels, label_lens = _setup()
    eng.forward(feats.cuda().bfloat16().contiguous(), lens.cuda())
    loss = eng.loss_and_backward(labels.cuda(), label_lens.cuda())
    torch.cuda.synchronize()
    p64 = {k: v.double() for k, v in params.items()}
    ref_loss, _, _ = TT.forward_loss(p64, MINI_JASPER, feats.double(), lens.long(), labels.long(), label_lens.long())
    assert abs(float(loss.mean()) - float(ref_loss)) < 1e-2 * abs(float(ref_loss))


def test_parameter_gradients_match_oracle_for_a_fixed_cotangent():
    """Backward pass (BN bwd, dgrad, wgrad, dense-residual accumulation, FC bwd) against autograd of
    the oracle for loss = <logits, R> with a fixed R, so the comparison is not amplified by the CTC
    posterior's sensitivity to the logits (the CTC gradient itself is pinned in test_kernels_gpu)."""
    from oracle import torch_twin as TT
    eng, params, feats, lens, labels, label_lens = _setup()
    logits, out_lens = eng.forward(feats.cuda().bfloat16().contiguous(), lens.cuda())
    g = torch.Generator().manual_seed(9)
    R = torch.randn(logits.shape, generator=g)
    R = R * TT.sequence_mask(out_lens.cpu().long(), logits.shape[1], R.dtype)  # no gradient on padded frames
    eng.backward_from_dlogits(R.cuda())
    torch.cuda.synchronize()
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    enc, _ = TT.tdnn_encode(feats.double(), lens.long(), MINI_JASPER, p64, emulate_storage=True)
    ref_logits = TT.fc_decode(enc, p64["fc/kernel"], p64["fc/bias"]).transpose(0, 1)
    (ref_logits * R.double()).sum().backward()
    worst = {}
    for name, _ in eng.named_parameters():
        worst[name] = _rel_l2(eng.param_view(name, eng.grad), p64[name].grad)
    bad = {k: round(v, 4) for k, v in worst.items() if v > 3e-2}
    assert not bad, "gradient mismatch: %r" % bad


def test_training_reduces_loss_and_matches_oracle_optimizer_direction():
    """A few full steps (fwd, bwd, LARC + NovoGrad) lower the CTC loss on a fixed batch."""
    from openseq2seq_b200.engine import JasperEngine
    eng, params, feats, lens, labels, label_lens = _setup()
    eng.set_optimizer(algo="novograd", beta1=0.95, beta2=0.98, weight_decay=0.001, larc_eta=0.001,
                      learning_rate=0.02, min_lr=1e-5, power=2.0, decay_steps=200, loss_scaling=True)
    eng.load_parameters(params)
    x = feats.cuda().bfloat16().contiguous()
    losses = []
    for _ in range(30):
        l = eng.train_step(x, lens.cuda(), labels.cuda(), label_lens.cuda())
        losses.append(float(l.mean()))
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.7 * losses[0], losses
    assert int(eng.istate[2]) == 30 and int(eng.istate[4]) == 0
