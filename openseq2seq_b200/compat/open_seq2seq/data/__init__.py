from .data_layer import DataLayer  # noqa: F401
from .speech2text.speech2text import Speech2TextDataLayer  # noqa: F401
