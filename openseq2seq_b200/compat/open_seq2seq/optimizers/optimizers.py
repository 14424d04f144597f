"""optimize_loss's configuration surface (open_seq2seq/optimizers/optimizers.py:36-44,107-286),
translated into JasperEngine.set_optimizer keyword arguments."""
from .lr_policies import cosine_decay, exp_decay, fixed_lr, poly_decay
from .novograd import NovoGrad


class _Momentum(object):
    engine_algo = "momentum"


class _Adam(object):
    engine_algo = "adam"


OPTIMIZER_CLS_NAMES = {
    "Adam": _Adam,
    "Momentum": _Momentum,
    "NovoGrad": NovoGrad,
}


def optimizer_engine_kwargs(params, last_step):
    """Model params -> dict for JasperEngine.set_optimizer (models/model.py:475-525 wiring:
    decay_steps defaults to last_step - begin_decay_at)."""
    opt = params.get("optimizer", "Momentum")
    if isinstance(opt, str):
        if opt not in OPTIMIZER_CLS_NAMES:
            raise ValueError("Optimizer name should be one of [{}], you provided {}.".format(
                ", ".join(OPTIMIZER_CLS_NAMES), opt))
        opt = OPTIMIZER_CLS_NAMES[opt]
    algo = getattr(opt, "engine_algo", None)
    if algo is None:
        raise NotImplementedError("optimizer %r has no fused B200 step (NovoGrad, Momentum and Adam are built)" % (opt,))
    op = dict(params.get("optimizer_params", {}))
    kw = {"algo": algo}
    if algo == "novograd":
        for k in ("beta1", "beta2", "epsilon", "weight_decay", "grad_averaging"):
            if k in op:
                kw[k] = op[k]
    elif algo == "adam":
        # tf.train.AdamOptimizer defaults
        kw["beta1"], kw["beta2"] = op.get("beta1", 0.9), op.get("beta2", 0.999)
        kw["epsilon"] = op.get("epsilon", 1e-8)
    else:
        kw["momentum"] = op.get("momentum", 0.9)
        if "weight_decay" in op:
            kw["weight_decay"] = op["weight_decay"]
    kw["iter_size"] = int(params.get("iter_size", 1))
    # tf.contrib.layers.l2_regularizer(scale) from the model params (every layer of this path is built
    # with the same regularizer: encoder.py / decoder.py read it from the model when not overridden)
    reg, rp = params.get("regularizer"), params.get("regularizer_params") or {}
    if reg is not None:
        if getattr(reg, "__name__", "") != "l2_regularizer":
            raise NotImplementedError("regularizer %r: only tf.contrib.layers.l2_regularizer is built" % (reg,))
        kw["l2_regularizer_scale"] = float(rp.get("scale", 0.0))
    if params.get("max_grad_norm") is not None and params.get("larc_params") is not None:
        raise AttributeError("LARC and gradient norm clipping should not be used together")
    if params.get("max_grad_norm") is not None:
        kw["max_grad_norm"] = float(params["max_grad_norm"])      # optimizers.py:408-433, global-norm clipping
    if params.get("freeze_variables_regex") is not None:
        kw["freeze_variables_regex"] = params["freeze_variables_regex"]   # models/model.py:502-507
    larc = params.get("larc_params")
    if larc is not None:
        kw["larc_eta"] = larc["larc_eta"]
        kw["larc_mode"] = larc.get("larc_mode", "clip")
        kw["larc_min_update"] = larc.get("min_update", 1e-7)
        kw["larc_eps"] = larc.get("epsilon", 1e-7)
    policy = params.get("lr_policy")
    lp = dict(params.get("lr_policy_params", {}))
    kw["learning_rate"] = lp.get("learning_rate", 0.01)
    if policy is None or getattr(policy, "__name__", "") == "fixed_lr":
        kw["decay_steps"] = 0
        kw["lr_policy"] = "fixed_lr"
    elif getattr(policy, "__name__", "") == "poly_decay" or policy is poly_decay:
        begin = lp.get("begin_decay_at", 0)
        kw["begin_decay_at"] = begin
        kw["decay_steps"] = lp.get("decay_steps", max(int(last_step) - begin, 1))
        kw["power"] = lp.get("power", 1.0)
        kw["min_lr"] = lp.get("min_lr", 0.0)
        kw["warmup_steps"] = lp.get("warmup_steps", 0)
    elif getattr(policy, "__name__", "") == "cosine_decay" or policy is cosine_decay:
        begin = lp.get("begin_decay_at", 0)
        kw["lr_policy"] = "cosine_decay"
        kw["begin_decay_at"] = begin
        kw["decay_steps"] = lp.get("decay_steps", max(int(last_step) - begin, 1))
        kw["min_lr"] = lp.get("min_lr", 0.0)
        kw["warmup_steps"] = lp.get("warmup_steps", 0)
    elif getattr(policy, "__name__", "") == "exp_decay" or policy is exp_decay:
        begin = lp.get("begin_decay_at", 0)
        kw["lr_policy"] = "exp_decay"
        kw["begin_decay_at"] = begin
        kw["decay_steps"] = lp["decay_steps"]
        kw["decay_rate"] = lp["decay_rate"]
        kw["use_staircase_decay"] = lp["use_staircase_decay"]
        kw["min_lr"] = lp.get("min_lr", 0.0)
    else:
        raise NotImplementedError("lr_policy %r is not fused on device (fixed_lr, poly_decay, cosine_decay and "
                                  "exp_decay are)" % (policy,))
    ls = params.get("loss_scaling", 1.0)
    if isinstance(ls, str):
        if ls.lower() != "backoff":
            raise NotImplementedError("loss_scaling %r: only 'Backoff' is built" % ls)
        kw["loss_scaling"] = True
        lsp = params.get("loss_scaling_params", {}) or {}
        for k in ("scale_min", "scale_max", "step_factor", "step_window"):
            if k in lsp:
                kw[k] = lsp[k]
    else:
        kw["loss_scaling"] = False
        kw["initial_scale"] = float(ls)
    return kw
