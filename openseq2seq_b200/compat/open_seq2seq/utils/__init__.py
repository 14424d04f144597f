from .utils import (check_params, deco_print, flatten_dict, get_base_config, nest_dict,  # noqa: F401
                    nested_update, create_model, create_logdir, check_logdir)
from .funcs import train, evaluate, infer  # noqa: F401
