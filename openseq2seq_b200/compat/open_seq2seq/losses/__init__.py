from .loss import Loss  # noqa: F401
from .ctc_loss import CTCLoss  # noqa: F401
