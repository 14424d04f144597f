from .model import Model  # noqa: F401
from .encoder_decoder import EncoderDecoderModel  # noqa: F401
from .speech2text import Speech2Text  # noqa: F401
