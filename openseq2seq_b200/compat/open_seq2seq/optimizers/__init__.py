from .optimizers import OPTIMIZER_CLS_NAMES, optimizer_engine_kwargs  # noqa: F401
