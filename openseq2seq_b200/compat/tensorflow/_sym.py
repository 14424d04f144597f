class DType(object):
    def __init__(self, name):
        self.name = name
        self.base_dtype = self

    def __repr__(self):
        return "tf." + self.name

    def __eq__(self, other):
        return isinstance(other, DType) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


class Sym(object):
    """Symbolic activation value: ('x',) -> ('relu',) -> ('relu_clip', c)."""

    def __init__(self, kind="x", clip=0.0):
        self.kind, self.clip = kind, clip


def relu(x):
    if isinstance(x, Sym):
        if x.kind != "x":
            raise NotImplementedError("nested activations are not built")
        return Sym("relu")
    raise TypeError("tf.nn.relu stand-in only traces activation functions")


def minimum(a, b):
    s, c = (a, b) if isinstance(a, Sym) else (b, a)
    if isinstance(s, Sym) and s.kind == "relu" and isinstance(c, (int, float)):
        return Sym("relu_clip", float(c))
    raise NotImplementedError("tf.minimum stand-in only supports min(relu(x), const)")


def resolve_activation(fn):
    """Returns (apply_relu, clip) for an activation callable from a config (None -> identity)."""
    if fn is None:
        return 0, 0.0
    out = fn(Sym("x"))
    if isinstance(out, Sym):
        if out.kind == "relu":
            return 1, 0.0
        if out.kind == "relu_clip":
            return 1, out.clip
        if out.kind == "x":
            return 0, 0.0
    raise NotImplementedError("activation function is not one of identity / relu / clipped relu")
