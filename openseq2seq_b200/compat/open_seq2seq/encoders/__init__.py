from .encoder import Encoder  # noqa: F401
from .tdnn_encoder import TDNNEncoder  # noqa: F401


class DeepSpeech2Encoder(Encoder):
    """Symbol kept so DS2 configs import; the conv2d + RNN kernels are outside the built hot path
    (out of scope for kernels)."""

    @staticmethod
    def get_required_params():
        return None

    @staticmethod
    def get_optional_params():
        return None

    def _encode(self, input_dict):
        raise NotImplementedError("DeepSpeech2Encoder has no B200 kernels in this build (Jasper/TDNN path only)")
