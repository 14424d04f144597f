from .decoder import Decoder  # noqa: F401
from .fc_decoders import FullyConnectedCTCDecoder, FullyConnectedTimeDecoder  # noqa: F401
