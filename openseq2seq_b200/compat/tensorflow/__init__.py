"""Minimal `tensorflow` stand-in for loading OpenSeq2Seq speech2text configs unchanged.

Only what example_configs/speech2text/*.py and test_utils/test_speech_configs/*.py touch:
tf.float16 / tf.float32, tf.nn.relu, tf.minimum, tf.contrib.layers.{xavier_initializer,
l2_regularizer}.  Activation lambdas such as `lambda x: tf.minimum(tf.nn.relu(x), 20.0)`
(ds2_toy_config.py:79) are resolved by tracing them on a symbolic probe (see resolve_activation).
This is NOT TensorFlow; anything else raises AttributeError.
"""
from . import contrib  # noqa: F401
from . import nn  # noqa: F401
from ._sym import DType, Sym, minimum, resolve_activation  # noqa: F401

__version__ = "0.0-os2s-b200-shim"
float16 = DType("float16")
float32 = DType("float32")
bfloat16 = DType("bfloat16")
int32 = DType("int32")
