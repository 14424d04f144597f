"""NovoGrad symbol for configs (`"optimizer": NovoGrad`).  The update itself is the fused CUDA
step os2s_opt_step (open_seq2seq/optimizers/novograd.py:30-126 restated in optim.cu)."""


class NovoGrad(object):
    engine_algo = "novograd"

    def __init__(self, learning_rate=1.0, beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.0,
                 grad_averaging=False, use_locking=False, name="NovoGrad"):
        self.hparams = dict(beta1=beta1, beta2=beta2, epsilon=epsilon, weight_decay=weight_decay,
                            grad_averaging=grad_averaging)
