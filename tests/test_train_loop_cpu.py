"""Host-side behaviour of the training driver (open_seq2seq/utils/funcs.py:150-218 and the hooks it
installs, utils/hooks.py:166-245) with a stub model: no GPU, no kernels -- only the loop's bookkeeping:
checkpoint cadence and retention, evaluation every eval_steps AND at the last step, best-validation
copies under logdir/best_models, resuming from the restored global step, the objects/s meter."""
import argparse
import os

import torch

import openseq2seq_b200.compat as compat

compat.install()
from open_seq2seq.utils import checkpoint as ckpt  # noqa: E402
from open_seq2seq.utils.funcs import train  # noqa: E402
from open_seq2seq.utils.utils import check_logdir  # noqa: E402


class _Engine(object):
    """The attributes checkpoint.save / restore and train() touch."""

    def __init__(self):
        self.w = {"a/kernel": torch.zeros(3), "a/bn/gamma": torch.ones(2)}
        self.mom = {k: torch.zeros_like(v) for k, v in self.w.items()}
        self.moving = {"a/bn": torch.zeros(2, 2)}
        self.fstate, self.istate = torch.zeros(8), torch.zeros(8, dtype=torch.int64)
        self._opt = {"ema": torch.zeros(2)}
        self.mom2, self.grad_acc, self._micro = None, None, 0
        self.training = True

    def named_parameters(self):
        return list(self.w.items())

    def param_view(self, n, buf=None):
        return (self.w if buf is None else buf)[n]

    def load_parameters(self, params):
        for k, v in params.items():
            self.w[k].copy_(v)

    def set_training(self, flag):
        self.training = flag

    def greedy_decode(self):
        return None


class _DataLayer(object):
    def __init__(self, n):
        self.n = n
        self.iterator = None
        self.build_graph()

    def build_graph(self):
        self.iterator = iter(range(10 ** 9)) if self.n is None else iter(range(self.n))


class _Model(object):
    on_horovod, hvd = False, None

    def __init__(self, params, engine, eval_losses=None, n_eval_batches=2):
        self.params, self.engine, self.last_step = params, engine, params["max_steps"]
        self._dl = _DataLayer(None if eval_losses is None else n_eval_batches)
        self._eval_losses = iter(eval_losses or [])
        self.evals = 0

    def get_data_layer(self):
        return self._dl

    def train_step(self, batch):
        self.engine.istate[2] += 1
        self.engine.w["a/kernel"] += 1.0
        return torch.tensor(1.0), 100.0

    def maybe_print_logs(self, *a):
        pass

    # eval side
    def eval_step(self, batch):
        return self._cur, {"outputs": [None]}

    def evaluate(self, batch, out):
        return 0

    def finalize_evaluation(self, results):
        self.evals += 1
        return {}


class _EvalModel(_Model):
    def __init__(self, params, engine, losses):
        _Model.__init__(self, params, engine, eval_losses=losses)
        self._losses = list(losses)

    def get_data_layer(self):
        # a new "epoch": next validation loss
        if self._dl.iterator is None or getattr(self, "_fresh", True):
            self._cur = self._losses.pop(0) if self._losses else 9.0
            self._fresh = False
        return self._dl

    def finalize_evaluation(self, results):
        self._fresh = True
        return _Model.finalize_evaluation(self, results)


def test_train_loop_checkpoints_evaluates_and_keeps_best_models(tmp_path):
    logdir = str(tmp_path / "log")
    params = {"max_steps": 10, "print_loss_steps": None, "print_samples_steps": None, "save_checkpoint_steps": 4,
              "eval_steps": 5, "num_checkpoints": 2, "logdir": logdir, "bench_start": 2}
    eng = _Engine()
    tm = _Model(params, eng)
    # RunEvaluationHook fires on the first run, then every eval_steps runs, and on the last step (global steps 0, 5
    # and 9); the losses are 7.0, 3.0 (an improvement, saved as global step + 1 = 6) and 5.0
    em = _EvalModel(params, eng, losses=[7.0, 3.0, 5.0])
    rate = train(tm, em)
    assert int(eng.istate[2]) == 10 and float(eng.w["a/kernel"][0]) == 10.0
    assert em.evals == 3
    names = sorted(os.listdir(logdir))
    assert names == ["best_models", "model.ckpt-10.pt", "model.ckpt-8.pt"]   # 4 was rotated out (keep 2)
    assert sorted(os.listdir(os.path.join(logdir, "best_models"))) == ["val_loss=3.0000-step-6.pt",
                                                                       "val_loss=7.0000-step-1.pt"]
    assert rate is not None and rate > 0
    # eval / infer pick the latest or, with restore_best_checkpoint, the best checkpoint (utils.py:676-690)
    args = argparse.Namespace(mode="eval", enable_logs=False, continue_learning=False, no_dir_check=False,
                              benchmark=False)
    assert check_logdir(args, {"logdir": logdir}).endswith("model.ckpt-10.pt")
    assert check_logdir(args, {"logdir": logdir}, True).endswith(os.path.join("best_models", "val_loss=3.0000-step-6.pt"))
    # resume: a restored engine continues from its global step instead of starting over
    eng2 = _Engine()
    assert ckpt.restore(eng2, os.path.join(logdir, "model.ckpt-8.pt")) == 8
    assert int(eng2.istate[2]) == 8 and float(eng2.w["a/kernel"][0]) == 8.0
    params2 = dict(params, logdir=str(tmp_path / "log2"), eval_steps=None)
    train(_Model(params2, eng2))
    assert int(eng2.istate[2]) == 10 and float(eng2.w["a/kernel"][0]) == 10.0


def test_evaluate_restores_the_checkpoint_and_skipped_steps_do_not_count(tmp_path):
    """(a) funcs.evaluate(model, checkpoint) runs on the restored weights (utils/funcs.py:205-218) -- not on a
    freshly initialised engine; (b) the loop stops on the DEVICE global step, which an overflow-skipped step
    does not advance (mp_wrapper.py:115-120); (c) load_model restores name- and shape-matching variables only
    and leaves the step counters alone (utils/funcs.py:117-144)."""
    from open_seq2seq.utils.funcs import evaluate
    logdir = str(tmp_path / "log")
    eng = _Engine()
    eng.w["a/kernel"] += 7.0
    eng.istate[2] = 3
    path = ckpt.save(eng, logdir, 3)
    fresh = _Engine()
    em = _EvalModel({"max_steps": 1}, fresh, losses=[1.0])
    out = evaluate(em, path)
    assert float(fresh.w["a/kernel"][0]) == 7.0 and em.evals == 1 and out["Eval loss"] == 1.0

    class _Skippy(_Model):
        def train_step(self, batch):
            self.attempts = getattr(self, "attempts", 0) + 1
            if self.attempts % 3 != 0:          # every third attempt overflows: global_step stays
                self.engine.istate[2] += 1
            return torch.tensor(1.0), 100.0
    e2 = _Engine()
    m = _Skippy({"max_steps": 6, "print_loss_steps": None, "print_samples_steps": None, "save_checkpoint_steps": None,
                 "eval_steps": None, "logdir": None, "bench_start": 0}, e2)
    train(m)
    assert int(e2.istate[2]) == 6 and m.attempts == 8       # 6 applied + 2 skipped attempts

    class _E3(_Engine):
        def __init__(self):
            _Engine.__init__(self)
            self.w["b/kernel"] = torch.zeros(5)              # not in the checkpoint
            self.w["a/bn/gamma"] = torch.ones(3)             # shape differs from the checkpoint's
            self.mom = {k: torch.zeros_like(v) for k, v in self.w.items()}
            self.by_name = {k: {"shape": tuple(v.shape)} for k, v in self.w.items()}
    e3 = _E3()
    n = ckpt.restore_partial(e3, logdir)
    assert float(e3.w["a/kernel"][0]) == 7.0 and float(e3.w["a/bn/gamma"][0]) == 1.0 and int(e3.istate[2]) == 0
    assert n == 2    # a/kernel + the BN moving statistics
