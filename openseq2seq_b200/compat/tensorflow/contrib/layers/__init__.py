"""tf.contrib.layers.xavier_initializer / l2_regularizer stand-ins (descriptors, no graph)."""


class _Xavier(object):
    def __init__(self, uniform=True, seed=None, dtype=None):
        self.uniform, self.seed = uniform, seed


def xavier_initializer(uniform=True, seed=None, dtype=None):
    return _Xavier(uniform=uniform, seed=seed, dtype=dtype)


class _L2(object):
    def __init__(self, scale):
        self.scale = float(scale)


def l2_regularizer(scale, scope=None):
    return _L2(scale)
