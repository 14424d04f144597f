"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the reference's optimizer chain for the Jasper config
(optimizer=NovoGrad, larc_params, dtype="mixed", loss_scaling="Backoff", lr_policy=poly_decay):

  open_seq2seq/optimizers/optimizers.py:107-286      optimize_loss: order of operations
  open_seq2seq/optimizers/optimizers.py:77-104       reduce_gradients (Horovod mean)
  open_seq2seq/optimizers/optimizers.py:333-377      LARC (clip / scale mode)
  open_seq2seq/optimizers/mp_wrapper.py:44-122       loss scaling, fp32 masters, skip on overflow
  open_seq2seq/optimizers/automatic_loss_scaler.py:31-106  check_grads, BackoffScaler
  open_seq2seq/optimizers/novograd.py:93-126         NovoGrad on top of tf.train.MomentumOptimizer
  open_seq2seq/optimizers/lr_policies.py:95-131      poly_decay -> tf.train.polynomial_decay

TensorFlow semantics restated: polynomial_decay
  lr = (lr0 - end) * (1 - min(step, D)/D)**power + end;  MomentumOptimizer  accum <- m*accum + g,
  w <- w - lr*accum;  hvd.allreduce = mean over ranks.

Quirk kept on purpose (`ema_persist=False`, the reference as written): novograd.py:110-113 rebinds
a Python list entry to a tf.cond tensor and never assigns the variable, so the per-tensor second
moment is v_t = ||g_t||^2 at every step.  `ema_persist=True` is the corrected NovoGrad
(v_t = beta2*v_{t-1} + (1-beta2)*||g_t||^2 after the first step).
Parity status: unpinned by the reference's own tests except mp_wrapper_test.py (regulariser grad
1e-8) and optimizers_test.py (iter_size); see tests/test_oracle.py.  In the build container the reference's OWN
lr_policies.py, post_process_gradients (global-norm clipping + LARC) and NovoGrad class are executed over stand-ins
for the TensorFlow symbols they touch and compared with the functions below
(tests/test_reference_config_executed_cpu.py); the Backoff scaler and the MP wrapper are restated only.
"""
import numpy as np


def poly_decay(step, learning_rate, decay_steps, power=1.0, begin_decay_at=0, min_lr=0.0,
               warmup_steps=0):
    """lr_policies.py:95-131."""
    lr0 = float(learning_rate)
    if warmup_steps > 0 and step < warmup_steps:
        lr0 = lr0 * float(step) / float(warmup_steps)
    if step < begin_decay_at:
        return lr0
    s = min(step - begin_decay_at, decay_steps)
    return (lr0 - min_lr) * (1.0 - float(s) / float(decay_steps)) ** power + min_lr


def cosine_decay(step, learning_rate, decay_steps, power=1.0, begin_decay_at=0, min_lr=0.0, warmup_steps=0):
    """lr_policies.py:134-170: tf.train.cosine_decay with alpha = min_lr (a FRACTION of the learning
    rate -- that is what the reference passes)."""
    lr0 = float(learning_rate)
    if warmup_steps > 0 and step < warmup_steps:
        lr0 = lr0 * float(step) / float(warmup_steps)
    if step < begin_decay_at:
        return lr0
    s = min(step - begin_decay_at, decay_steps)
    cosd = 0.5 * (1.0 + np.cos(np.pi * float(s) / float(decay_steps)))
    return lr0 * ((1.0 - min_lr) * cosd + min_lr)


def exp_decay(step, learning_rate, decay_steps, decay_rate, use_staircase_decay, begin_decay_at=0, min_lr=0.0):
    """lr_policies.py:55-92: tf.train.exponential_decay from begin_decay_at on, floored at min_lr."""
    if step < begin_decay_at:
        return max(float(learning_rate), min_lr)
    e = float(step - begin_decay_at) / float(decay_steps)
    if use_staircase_decay:
        e = np.floor(e)
    return max(float(learning_rate) * decay_rate ** e, min_lr)


def fixed_lr(step, learning_rate):
    """lr_policies.py:15-27."""
    return float(learning_rate)


class BackoffScaler(object):
    """automatic_loss_scaler.py:50-110."""

    def __init__(self, scale_min=1.0, scale_max=2.0 ** 14, step_factor=2.0, step_window=2000):
        self.scale_min, self.scale_max = scale_min, scale_max
        self.step_factor, self.step_window = step_factor, step_window
        self.iteration = 0
        self.last_overflow_iteration = -1
        self.scale = float(scale_max)

    def update(self, has_nan, amax):
        overflow = bool(has_nan) or bool(np.isinf(amax))
        if overflow:
            self.scale = float(np.clip(self.scale / self.step_factor, self.scale_min, self.scale_max))
            self.last_overflow_iteration = self.iteration
        else:
            since = self.iteration - self.last_overflow_iteration
            if since % self.step_window == 0:
                self.scale = float(np.clip(self.scale * self.step_factor, self.scale_min, self.scale_max))
        self.iteration += 1
        return overflow


def larc(grads, weights, lr, larc_eta, larc_mode="clip", min_update=1e-7, eps=1e-7):
    """optimizers.py:333-377 applied to every (g, v) pair; fp32 arithmetic like the reference."""
    out = []
    for g, v in zip(grads, weights):
        v_norm = np.float32(np.sqrt(np.sum(np.square(v.astype(np.float32), dtype=np.float32), dtype=np.float32)))
        g_norm = np.float32(np.sqrt(np.sum(np.square(g.astype(np.float32), dtype=np.float32), dtype=np.float32)))
        if larc_mode == "clip":
            r = max(np.float32(larc_eta) * v_norm / (np.float32(lr) * (g_norm + np.float32(eps))), np.float32(min_update))
            r = min(r, np.float32(1.0))
        else:
            r = max(np.float32(larc_eta) * v_norm / (g_norm + np.float32(eps)), np.float32(min_update))
        out.append((np.float32(r) * g).astype(np.float32))
    return out


def clip_by_global_norm(grads, clip_norm):
    """optimizers.py:380-480 (_clip_gradients_by_norm / _clip_by_global_norm with the fp32 global norm of
    _global_norm_with_cast): every gradient times clip_norm * min(1 / global_norm, 1 / clip_norm)."""
    sq = np.float32(0.0)
    for g in grads:
        sq = np.float32(sq + np.sum(np.square(g.astype(np.float32), dtype=np.float32), dtype=np.float32))
    gnorm = np.float32(np.sqrt(sq))
    scale = np.float32(clip_norm) * min(np.float32(1.0) / gnorm, np.float32(1.0) / np.float32(clip_norm))
    return [(g.astype(np.float32) * np.float32(scale)).astype(np.float32) for g in grads], gnorm


class NovoGradState(object):
    def __init__(self, n):
        self.ema = [0.0] * n
        self.momentum = [None] * n


def novograd_step(weights, grads, state, lr, beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.0,
                  grad_averaging=False, ema_persist=False):
    """novograd.py:93-126 + MomentumOptimizer.  Updates weights/state in place (fp32)."""
    for i, (w, g) in enumerate(zip(weights, grads)):
        g = g.astype(np.float32)
        g2 = np.float32(np.sum(np.square(g, dtype=np.float32), dtype=np.float32))
        if state.ema[i] == 0.0:
            v = g2
        else:
            v = np.float32(state.ema[i] * beta2 + g2 * (1.0 - beta2))
        if ema_persist:
            state.ema[i] = float(v)
        g = g * np.float32(1.0 / np.sqrt(np.float32(v) + np.float32(epsilon)))
        if weight_decay > 0.0:
            g = g + np.float32(weight_decay) * w
        if grad_averaging:
            g = g * np.float32(1.0 - beta1)
        if state.momentum[i] is None:
            state.momentum[i] = np.zeros_like(w, dtype=np.float32)
        state.momentum[i] = np.float32(beta1) * state.momentum[i] + g
        w -= np.float32(lr) * state.momentum[i]


def momentum_step(weights, grads, state, lr, momentum=0.9, weight_decay=0.0):
    """tf.train.MomentumOptimizer (use_nesterov=False): accum <- m*accum + g; w <- w - lr*accum."""
    for i, (w, g) in enumerate(zip(weights, grads)):
        g = g.astype(np.float32)
        if weight_decay > 0.0:
            g = g + np.float32(weight_decay) * w
        if state.momentum[i] is None:
            state.momentum[i] = np.zeros_like(w, dtype=np.float32)
        state.momentum[i] = np.float32(momentum) * state.momentum[i] + g
        w -= np.float32(lr) * state.momentum[i]


class AdamState(object):
    def __init__(self, n):
        self.m = [None] * n
        self.v = [None] * n
        self.t = 0


def adam_step(weights, grads, state, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, weight_decay=0.0):
    """tf.train.AdamOptimizer (the reference's "Adam", optimizers.py:36-44):
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2; w <- w - lr_t m/(sqrt(v)+eps)."""
    state.t += 1
    lr_t = np.float32(lr) * np.float32(np.sqrt(1.0 - beta2 ** state.t) / (1.0 - beta1 ** state.t))
    for i, (w, g) in enumerate(zip(weights, grads)):
        g = g.astype(np.float32)
        if weight_decay > 0.0:
            g = g + np.float32(weight_decay) * w
        if state.m[i] is None:
            state.m[i] = np.zeros_like(w, dtype=np.float32)
            state.v[i] = np.zeros_like(w, dtype=np.float32)
        state.m[i] = np.float32(beta1) * state.m[i] + np.float32(1.0 - beta1) * g
        state.v[i] = np.float32(beta2) * state.v[i] + np.float32(1.0 - beta2) * g * g
        w -= lr_t * state.m[i] / (np.sqrt(state.v[i]) + np.float32(epsilon))


def check_grads(grads):
    """automatic_loss_scaler.py:31-47."""
    has_nan = any(bool(np.isnan(g).any()) for g in grads)
    amax = max(float(np.max(np.abs(g))) if g.size else 0.0 for g in grads)
    return has_nan, amax


def train_step(weights, scaled_grads_per_rank, state, scaler, step, lr_fn, opt_params, larc_params=None,
               ema_persist=False, algo="novograd", reg_scales=None, max_grad_norm=None):
    """One optimize_loss step (optimizers.py:208-281 + mp_wrapper.py:44-122).

    scaled_grads_per_rank: list over ranks of lists of gradients of (loss * scaler.scale) wrt the
    half-precision weights (already rounded as the backward pass produced them).
    Returns (skipped, lr, new_step)."""
    scale = np.float32(scaler.scale)
    n_rank = len(scaled_grads_per_rank)
    grads = []
    for i in range(len(weights)):
        g = [scaled_grads_per_rank[r][i].astype(np.float32) * (np.float32(1.0) / scale) for r in range(n_rank)]
        grads.append(np.sum(g, axis=0, dtype=np.float32) / np.float32(n_rank) if n_rank > 1 else g[0])
    if reg_scales is not None:
        # deferred regulariser gradient on the fp32 copy, added after the un-scaling (mp_wrapper.py:81-95)
        grads = [g + np.float32(r) * w if r else g for g, w, r in zip(grads, weights, reg_scales)]
    lr = lr_fn(step)
    if max_grad_norm:
        # post_process_gradients clips before LARC (optimize_loss refuses both at once, optimizers.py:161-164)
        grads, _ = clip_by_global_norm(grads, max_grad_norm)
    if larc_params is not None:
        grads = larc(grads, weights, lr, **larc_params)
    has_nan, amax = check_grads(grads)
    skipped = scaler.update(has_nan, amax)
    if skipped:
        return True, lr, step
    if algo == "novograd":
        novograd_step(weights, grads, state, lr, ema_persist=ema_persist, **opt_params)
    elif algo == "momentum":
        momentum_step(weights, grads, state, lr, **opt_params)
    elif algo == "adam":
        adam_step(weights, grads, state, lr, **opt_params)
    else:
        raise ValueError(algo)
    return False, lr, step + 1
