"""Learning-rate policies as plain Python (open_seq2seq/optimizers/lr_policies.py).

Each policy is `f(global_step, learning_rate, **params) -> float`; poly_decay is the one Jasper
uses and the one evaluated ON DEVICE by os2s_opt_step (so the skip-step history never needs a host
sync); the host-side functions here are for logging / tests and for policies not fused yet."""
import math


def fixed_lr(global_step, learning_rate):
    return learning_rate


def poly_decay(global_step, learning_rate, decay_steps, power=1.0, begin_decay_at=0, min_lr=0.0,
               warmup_steps=0):
    """lr_policies.py:95-131 (tf.train.polynomial_decay, cycle=False)."""
    if warmup_steps > 0 and global_step < warmup_steps:
        learning_rate = learning_rate * float(global_step) / float(warmup_steps)
    if global_step < begin_decay_at:
        return learning_rate
    s = min(global_step - begin_decay_at, decay_steps)
    return (learning_rate - min_lr) * (1.0 - float(s) / float(decay_steps)) ** power + min_lr


def cosine_decay(global_step, learning_rate, decay_steps, power=1.0, begin_decay_at=0, min_lr=0.0,
                 warmup_steps=0):
    """lr_policies.py:134-172 (tf.train.cosine_decay with alpha = min_lr)."""
    if warmup_steps > 0 and global_step < warmup_steps:
        learning_rate = learning_rate * float(global_step) / float(warmup_steps)
    if global_step < begin_decay_at:
        return learning_rate
    s = min(global_step - begin_decay_at, decay_steps)
    cosine = 0.5 * (1.0 + math.cos(math.pi * float(s) / float(decay_steps)))
    return learning_rate * ((1.0 - min_lr) * cosine + min_lr)


def exp_decay(global_step, learning_rate, decay_steps, decay_rate, use_staircase_decay, begin_decay_at=0,
              min_lr=0.0):
    """lr_policies.py:55-92."""
    if global_step < begin_decay_at:
        return learning_rate
    e = float(global_step - begin_decay_at) / float(decay_steps)
    if use_staircase_decay:
        e = math.floor(e)
    return max(min_lr, learning_rate * decay_rate ** e)
