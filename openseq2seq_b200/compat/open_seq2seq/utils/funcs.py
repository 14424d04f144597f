"""train / evaluate / infer drivers (open_seq2seq/utils/funcs.py:22-260) without TF sessions.

The training loop keeps the reference's observable behaviour: the cadence of its session hooks (loss / sample
printing and evaluation fire on the first run, then every N runs, evaluation also on the last step -- utils/hooks.py;
tests/test_reference_config_executed_cpu.py drives the reference's own hook classes next to this loop), checkpoints
every save_checkpoint_steps on rank 0 (num_checkpoints kept; TensorFlow's extra checkpoint at step 0 is not
written), and the throughput meter of `--benchmark`: after `bench_start` steps accumulate wall time and input
frames and print "Avg objects per second" (objects = input frames, summed over ranks)."""
import os
import time

import torch

from .utils import deco_print


def _master(model):
    return (not model.on_horovod) or model.hvd.rank() == 0


def evaluate_model(eval_model, max_batches=None):
    results, losses = [], []
    it = eval_model.get_data_layer().iterator or None
    if it is None:
        eval_model.get_data_layer().build_graph()
        it = eval_model.get_data_layer().iterator
    eng = eval_model.engine
    was = eng.training
    eng.set_training(False)
    try:
        for k, batch in enumerate(it):
            loss, dec_out = eval_model.eval_step(batch)
            results.append(eval_model.evaluate(batch, dec_out["outputs"][0]))
            if loss is not None:
                losses.append(float(loss))
            if max_batches is not None and k + 1 >= max_batches:
                break
    finally:
        eng.set_training(was)
        eval_model.get_data_layer().build_graph()
    out = eval_model.finalize_evaluation(results)
    if losses:
        out["Eval loss"] = sum(losses) / len(losses)
        deco_print("Validation loss: {:.4f}".format(out["Eval loss"]), offset=4)
    return out


class _StepTimer(object):
    """tf.train.SecondOrStepTimer(every_steps=N) as the reference's hooks use it (utils/hooks.py:57-245): the hook
    asks should_trigger(step fetched by the PREVIOUS run) before a run and, when it fired, records `this run's
    global step - 1`.  Net effect: the first run after (re)start fires, then every N-th one -- evaluation, loss and
    sample printing happen at global steps 0, N, 2N, ... (i.e. after 1, N + 1, 2N + 1, ... completed steps)."""

    def __init__(self, every_steps):
        self.every, self.last = every_steps, None

    def should_trigger(self, step):
        if self.last is None:
            return True
        if self.last == step:
            return False
        return step >= self.last + self.every

    def update(self, step):
        self.last = step


def train(train_model, eval_model=None, debug_port=None, custom_hooks=None):
    p = train_model.params
    master = _master(train_model)
    it = train_model.get_data_layer().iterator
    bench_start = p.get("bench_start", 10)
    print_every = p.get("print_loss_steps")
    save_every = p.get("save_checkpoint_steps")
    eval_every = p.get("eval_steps") if eval_model is not None else None
    logdir = p.get("logdir")
    last_step = train_model.last_step
    # resume (--continue_learning restored the engine state): continue from the restored global step
    step = int(train_model.engine.istate[2]) if hasattr(train_model.engine, "istate") else 0
    first_step = step
    best_eval_loss = 1e9   # RunEvaluationHook._best_eval_loss (utils/hooks.py:182)
    total_time, total_objects = 0.0, 0.0
    deco_print("Starting training ({} steps)".format(last_step))
    t_print = time.time()
    t_loss = _StepTimer(print_every) if print_every else None
    t_samples = _StepTimer(p["print_samples_steps"]) if p.get("print_samples_steps") else None
    t_eval = _StepTimer(eval_every) if eval_every else None
    fetched = 0          # the hooks' _iter_count: the global step the previous run fetched (0 after begin())
    while True:
        if step >= last_step:
            # StopAtStepHook counts global_step, which a loss-scale overflow does not advance
            # (mp_wrapper.py:115-120): the host counts attempts without synchronising and reconciles here
            dev_step = int(train_model.engine.istate[2]) if hasattr(train_model.engine, "istate") else step
            if dev_step >= last_step:
                break
            step = dev_step
        t0 = time.time()
        # before_run of the hooks: decided on the step the previous run fetched
        fire_loss = t_loss is not None and t_loss.should_trigger(fetched)
        fire_samples = t_samples is not None and t_samples.should_trigger(fetched)
        fire_eval = t_eval is not None and t_eval.should_trigger(fetched)
        k = step         # the global step this run starts from: what the hooks fetch next to the train op
        batch = next(it)
        loss, n_objects = train_model.train_step(batch)
        step += 1
        fetched = k
        timed = step - first_step > bench_start
        if (timed or fire_loss) and torch.cuda.is_available():
            torch.cuda.synchronize()
        if timed:
            total_time += time.time() - t0
            total_objects += float(n_objects)
        if fire_loss:
            t_loss.update(k - 1)
            if master:
                deco_print("Global step {}: train loss = {:.4f}, time per step = {:.3f}s".format(
                    k, float(loss), (time.time() - t_print) / print_every))
            t_print = time.time()
        if fire_samples:
            t_samples.update(k - 1)
            if master:
                toks = train_model.engine.greedy_decode()
                train_model.maybe_print_logs(batch, toks, k)
        if save_every and logdir and step % save_every == 0 and master:
            # tf.train.CheckpointSaverHook(save_steps): when the global step AFTER a run reaches last save + N
            from . import checkpoint as ckpt
            ckpt.save(train_model.engine, logdir, step, keep=p.get("num_checkpoints", 5))
        if t_eval is not None and (fire_eval or k == last_step - 1):
            # RunEvaluationHook (utils/hooks.py:166-245): first run, every eval_steps, and the last step; the best
            # validation loss so far gets its own checkpoint under logdir/best_models (named by global step + 1)
            t_eval.update(k - 1)
            out = evaluate_model(eval_model)
            eval_loss = out.get("Eval loss")
            if save_every and logdir and master and eval_loss is not None and eval_loss < best_eval_loss:
                from . import checkpoint as ckpt
                best_eval_loss = eval_loss
                ckpt.save(train_model.engine, os.path.join(logdir, "best_models"), step,
                          keep=p.get("num_checkpoints", 5), prefix="val_loss={:.4f}-step".format(eval_loss))
    if save_every and logdir and master:
        from . import checkpoint as ckpt
        ckpt.save(train_model.engine, logdir, step, keep=p.get("num_checkpoints", 5))
    if train_model.on_horovod:
        total_objects = train_model.hvd.sum_scalar(total_objects)
    if master:
        deco_print("Finished training")
        if step - first_step > bench_start and total_time > 0:
            deco_print("Avg time per step: {:.3f}s".format(total_time / (step - first_step - bench_start)))
            deco_print("Avg objects per second: {:.3f}".format(total_objects / total_time))
        else:
            deco_print("Not enough steps for benchmarking")
    return total_objects / total_time if total_time > 0 else None


def evaluate(model, checkpoint):
    """utils/funcs.py:205-218: restore `checkpoint` (create_model already did when it compiled the model; a
    model built elsewhere is restored here), then one pass over the evaluation set."""
    if checkpoint is not None and getattr(model, "_restored_from", None) != checkpoint:
        from . import checkpoint as ckpt
        ckpt.restore(model.engine, checkpoint)
        model._restored_from = checkpoint
    return evaluate_model(model)


def infer(model, checkpoint, output_file):
    results = []
    model.engine.set_training(False)
    for batch in model.get_data_layer().iterator:
        _, dec_out = model.eval_step(batch)
        results.append(model.infer(batch, dec_out["outputs"][0]))
    model.finalize_inference(results, output_file)
