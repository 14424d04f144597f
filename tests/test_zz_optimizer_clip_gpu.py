"""Global-norm clipping (`max_grad_norm`, optimizers.py:380-480) and `freeze_variables_regex` (models/model.py:502-507)
on the device optimizer against the oracle, whose clipping is pinned on the executed reference
(tests/test_reference_config_executed_cpu.py).

Written after this round's GPU budget was spent: the test has NOT run on a B200 at commit time, so it is marked
`xfail(strict=False)` -- it reports XPASS when the kernel agrees with the oracle and cannot turn the suite red on
account of a mistake in the test itself.  (The file sorts last for the same reason.)
"""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="not yet run on a B200 (round-2 GPU budget exhausted); non-gating")]


@pytest.mark.parametrize("algo", ["novograd", "momentum"])
def test_global_norm_clipping_and_frozen_variables_vs_oracle(algo):
    from oracle import optimizer as OO
    from openseq2seq_b200.engine import JasperEngine
    from tests.common_cfg import MINI_JASPER
    opt_kw = {"novograd": dict(beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.001),
              "momentum": dict(momentum=0.9)}[algo]
    lr0, clip = 0.01, 0.05
    eng = JasperEngine(MINI_JASPER, 64, 29, world_size=1,
                       opt=dict(algo=algo, learning_rate=lr0, lr_policy="fixed_lr", loss_scaling=False,
                                initial_scale=64.0, max_grad_norm=clip,
                                freeze_variables_regex="ForwardPass/w2l_encoder/conv2.*", **opt_kw))
    names = [n for n, _ in eng.named_parameters()]
    frozen = [any(n == f or n.startswith(f) for f in eng.frozen_names) or n in eng.frozen_names for n in names]
    assert any(frozen) and not all(frozen) and all(n.startswith("conv2") for n, f in zip(names, frozen) if f)
    w_ref = [eng.param_view(n).detach().cpu().numpy().copy() for n in names]
    live = [i for i, f in enumerate(frozen) if not f]
    state = OO.NovoGradState(len(live))

    class _Static(object):
        scale = 64.0

        def update(self, has_nan, amax):
            return bool(has_nan) or bool(np.isinf(amax))

    rng = np.random.default_rng(7)
    step = 0
    for it in range(4):
        # iteration 3 stays below the threshold: the gradients must pass unchanged
        mag = 0.01 if it < 3 else 1e-6
        scaled = [(rng.standard_normal(w.shape) * mag * 64.0).astype(np.float32) for w in w_ref]
        eng.grad.zero_()
        for i, n in enumerate(names):
            eng.param_view(n, eng.grad).copy_(torch.tensor(scaled[i]))
        eng.optimizer_step()
        torch.cuda.synchronize()
        # the reference never sees the frozen variables: they are not in the optimizer's var_list
        w_live = [w_ref[i] for i in live]
        gnorm = OO.clip_by_global_norm([scaled[i] / np.float32(64.0) for i in live], clip)[1]
        skipped, lr, step = OO.train_step(w_live, [[scaled[i] for i in live]], state, _Static(), step,
                                          lambda s: OO.fixed_lr(s, lr0), opt_kw, larc_params=None, algo=algo,
                                          max_grad_norm=clip)
        assert not skipped and int(eng.istate[2]) == step
        assert abs(float(eng.fstate[2]) - float(gnorm)) <= 1e-4 * float(gnorm), (it, float(eng.fstate[2]), gnorm)
        assert (float(gnorm) > clip) == (it < 3)
        for i, n in enumerate(names):
            got = eng.param_view(n).cpu().numpy()
            assert np.abs(got - w_ref[i]).max() <= 3e-5 * max(1.0, np.abs(w_ref[i]).max()), (it, n, frozen[i])
