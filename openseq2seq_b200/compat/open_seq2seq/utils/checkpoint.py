"""Flat checkpoints (fp32 masters + momentum + BN moving stats + scaler/step state) with the
reference's directory conventions (logdir/model.ckpt-<step>, `num_checkpoints` kept;
open_seq2seq/utils/funcs.py:71-82).  The TF checkpoint FORMAT is out of scope."""
import glob
import os
import re

import torch


def latest_checkpoint(ckpt_dir):
    """Highest-step checkpoint of a directory (model.ckpt-<step>.pt, or val_loss=<x>-step-<step>.pt in
    best_models)."""
    files = glob.glob(os.path.join(ckpt_dir, "*-*.pt"))
    if not files:
        return None
    step = lambda f: int(re.search(r"-(\d+)\.pt$", f).group(1))
    return max(files, key=step)


def save(engine, logdir, step, keep=5, extra=None, prefix="model.ckpt"):
    """prefix: file stem; the evaluation hook's best-model copies use "val_loss=<loss>-step"
    (utils/hooks.py:228-236) inside logdir/best_models."""
    os.makedirs(logdir, exist_ok=True)
    path = os.path.join(logdir, "%s-%d.pt" % (prefix, step))
    state = {
        "params": {n: v.detach().cpu() for n, v in engine.named_parameters()},
        "momentum": {n: engine.param_view(n, engine.mom).detach().cpu() for n, _ in engine.named_parameters()},
        "moving": {k: v.detach().cpu() for k, v in engine.moving.items()},
        # Adam second moments / a partially filled iter_size accumulator, when the run has them
        "momentum2": ({n: engine.param_view(n, engine.mom2).detach().cpu() for n, _ in engine.named_parameters()}
                      if engine.mom2 is not None else None),
        "grad_acc": engine.grad_acc.detach().cpu() if engine.grad_acc is not None else None,
        "micro": engine._micro,
        "fstate": engine.fstate.cpu(), "istate": engine.istate.cpu(),
        "ema": engine._opt["ema"].cpu(), "step": step, "extra": extra or {},
    }
    torch.save(state, path)
    # keep the `keep` most recently written files of this directory (tf.train.Saver(max_to_keep))
    files = sorted(glob.glob(os.path.join(logdir, "*-*.pt")), key=os.path.getmtime)
    for f in files[:-keep]:
        if f != path:
            os.remove(f)
    return path


def restore_partial(engine, load_model_dir):
    """`load_model` (utils/funcs.py:117-144): every variable of the latest checkpoint in load_model_dir whose name
    and shape match a variable of this model is loaded (weights, BN moving statistics, optimizer slots);
    global_step and the loss-scaler state are not.  Returns the number of restored tensors."""
    path = load_model_dir if os.path.isfile(load_model_dir) else latest_checkpoint(load_model_dir)
    if path is None:
        raise IOError("load_model: there is no checkpoint in {}".format(load_model_dir))
    state = torch.load(path, map_location="cpu")
    n = 0
    ok = {}
    for name, v in state["params"].items():
        if name in engine.by_name and tuple(engine.by_name[name]["shape"]) == tuple(v.shape):
            ok[name] = v
    engine.load_parameters(ok)
    n += len(ok)
    for name in ok:
        m = state.get("momentum", {}).get(name)
        if m is not None:
            engine.param_view(name, engine.mom).copy_(m)
    for k, v in state.get("moving", {}).items():
        if k in engine.moving and tuple(engine.moving[k].shape) == tuple(v.shape):
            engine.moving[k].copy_(v)
            n += 1
    return n


def restore(engine, path):
    state = torch.load(path, map_location="cpu")
    engine.load_parameters(state["params"])
    for n, v in state["momentum"].items():
        engine.param_view(n, engine.mom).copy_(v)
    for k, v in state["moving"].items():
        engine.moving[k].copy_(v)
    if state.get("momentum2") is not None and engine.mom2 is not None:
        for n, v in state["momentum2"].items():
            engine.param_view(n, engine.mom2).copy_(v)
    if state.get("grad_acc") is not None and engine.grad_acc is not None:
        engine.grad_acc.copy_(state["grad_acc"])
        engine._micro = int(state.get("micro", 0))
    engine.fstate.copy_(state["fstate"])
    engine.istate.copy_(state["istate"])
    engine._opt["ema"].copy_(state["ema"])
    return state["step"]
