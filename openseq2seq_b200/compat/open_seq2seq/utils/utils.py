"""Config / flag system of the reference, restated (open_seq2seq/utils/utils.py:326-545,633-882).

Behavioural contract kept: config file executed with runpy (globals {'tf': tf}); must define
`base_model` and `base_params`; every int/float/str/bool leaf of base_params becomes a
`--nested/key` command-line override; per-class parameter schemas are validated by check_params
(type, allowed-value list, unknown-key rejection -> ValueError)."""
from __future__ import print_function

import argparse
import ast
import copy
import os
import runpy

import tensorflow as tf

_SCALAR = (int, float, str, bool)


def deco_print(line, offset=0, start="*** ", end="\n"):
    print(start + " " * offset + line, end=end)


def flatten_dict(dct):
    flat = {}
    for key, value in dct.items():
        if isinstance(value, _SCALAR):
            flat[key] = value
        elif isinstance(value, dict):
            for k, v in flatten_dict(value).items():
                flat[key + "/" + k] = v
    return flat


def nest_dict(flat):
    out = {}
    for key, value in flat.items():
        parts = key.split("/")
        cur = out
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = value
    return out


def nested_update(org, upd):
    for key, value in upd.items():
        if isinstance(value, dict):
            if key in org:
                if not isinstance(org[key], dict):
                    raise ValueError("Mismatch between org_dict and upd_dict at node {}".format(key))
                nested_update(org[key], value)
            else:
                org[key] = value
        else:
            org[key] = value


def check_params(config, required_dict, optional_dict):
    """Same acceptance rules and error type as the reference's check_params (utils.py:403-429)."""
    if required_dict is None or optional_dict is None:
        return

    def _check(pm, vals):
        if vals is None or vals == [] or vals == ():
            return
        if isinstance(vals, list):
            if config[pm] not in vals:
                raise ValueError("{} has to be one of {}".format(pm, vals))
        elif not isinstance(config[pm], vals):
            raise ValueError("{} has to be of type {}".format(pm, vals))

    for pm, vals in required_dict.items():
        if pm not in config:
            raise ValueError("{} parameter has to be specified".format(pm))
        _check(pm, vals)
    for pm, vals in optional_dict.items():
        if pm in config:
            _check(pm, vals)
    for pm in config:
        if pm not in required_dict and pm not in optional_dict:
            raise ValueError("Unknown parameter: {}".format(pm))


def get_base_config(args):
    """Parse CLI + config file -> (args, base_config, base_model, config_module)."""
    parser = argparse.ArgumentParser(description="Experiment parameters")
    parser.add_argument("--config_file", required=True, help="Path to the configuration file")
    parser.add_argument("--mode", default="train",
                        help='Could be "train", "eval", "train_eval" or "infer"')
    parser.add_argument("--infer_output_file", default="infer-out.txt")
    parser.add_argument("--continue_learning", dest="continue_learning", action="store_true")
    parser.add_argument("--no_dir_check", dest="no_dir_check", action="store_true")
    parser.add_argument("--benchmark", dest="benchmark", action="store_true")
    parser.add_argument("--bench_steps", type=int, default=20)
    parser.add_argument("--bench_start", type=int)
    parser.add_argument("--debug_port", type=int)
    parser.add_argument("--enable_logs", dest="enable_logs", action="store_true")
    parser.add_argument("--use_xla_jit", dest="use_xla_jit", action="store_true")
    args, unknown = parser.parse_known_args(args)
    if args.mode not in ["train", "eval", "train_eval", "infer", "interactive_infer"]:
        raise ValueError("Mode has to be one of ['train', 'eval', 'train_eval', 'infer', 'interactive_infer']")
    config_module = runpy.run_path(args.config_file, init_globals={"tf": tf})
    base_config = config_module.get("base_params", None)
    if base_config is None:
        raise ValueError("base_config dictionary has to be defined in the config file")
    base_config["use_xla_jit"] = args.use_xla_jit or base_config.get("use_xla_jit", False)
    base_model = config_module.get("base_model", None)
    if base_model is None:
        raise ValueError("base_config class has to be defined in the config file")
    parser_unk = argparse.ArgumentParser()
    for pm, value in flatten_dict(base_config).items():
        if type(value) in (int, float) or isinstance(value, str):
            parser_unk.add_argument("--" + pm, default=value, type=type(value))
        elif type(value) == bool:
            parser_unk.add_argument("--" + pm, default=value, type=ast.literal_eval)
    config_update = parser_unk.parse_args(unknown)
    nested_update(base_config, nest_dict(vars(config_update)))
    return args, base_config, base_model, config_module


def check_logdir(args, base_config, restore_best_checkpoint=False):
    """Reference rules (utils.py:633-709): training into a non-empty logdir needs
    --continue_learning; eval/infer need an existing checkpoint.  Returns checkpoint path or None."""
    logdir = base_config.get("logdir")
    if logdir is None:
        return None
    ckpt_dir = os.path.join(logdir, "logs") if args.enable_logs else logdir
    from open_seq2seq.utils.checkpoint import latest_checkpoint
    try:
        if args.mode in ("train", "train_eval"):
            if os.path.isfile(logdir):
                raise IOError("There is a file with the same name as \"logdir\" parameter.")
            if os.path.isdir(logdir) and os.listdir(logdir) != []:
                if not args.continue_learning:
                    raise IOError("Log directory is not empty. If you want to continue learning, "
                                  "you should provide \"--continue_learning\" flag")
                ckpt = latest_checkpoint(ckpt_dir)
                if ckpt is None:
                    raise IOError("There is no valid checkpoint in the log directory. Can't restore variables.")
                return ckpt
            if args.continue_learning:
                raise IOError("The log directory is empty or does not exist. "
                              "You should probably not provide \"--continue_learning\" flag?")
            return None
        if args.mode in ("infer", "eval", "interactive_infer"):
            if os.path.isdir(logdir) and os.listdir(logdir) != []:
                best_dir = os.path.join(ckpt_dir, "best_models")
                if restore_best_checkpoint and os.path.isdir(best_dir):
                    deco_print("Restoring from the best checkpoint")
                    ckpt_dir = best_dir
                else:
                    deco_print("Restoring from the latest checkpoint")
                ckpt = latest_checkpoint(ckpt_dir)
                if ckpt is None:
                    raise IOError("There is no valid checkpoint in the {}. Can't load model".format(ckpt_dir))
                return ckpt
            raise IOError("{} does not exist or is empty, can't restore model".format(ckpt_dir))
    except IOError as e:
        if args.no_dir_check:
            print("Warning: {}".format(e))
            print("Resuming operation since no_dir_check argument was provided")
            return None
        raise
    return None


def create_logdir(args, base_config):
    logdir = base_config.get("logdir")
    if logdir and args.mode in ("train", "train_eval"):
        os.makedirs(logdir, exist_ok=True)
    return None, None


def adjust_for_benchmark(train_config, args):
    """--benchmark (reference utils.py:846-865): no samples / summaries / checkpoints, empty logdir,
    max_steps = bench_steps instead of num_epochs, bench_start default 10."""
    deco_print("Adjusting config for benchmarking")
    train_config["print_samples_steps"] = None
    train_config["print_loss_steps"] = 1
    train_config["save_summaries_steps"] = None
    train_config["save_checkpoint_steps"] = None
    train_config["logdir"] = str("")
    if "num_epochs" in train_config:
        del train_config["num_epochs"]
    train_config["max_steps"] = args.bench_steps
    if args.bench_start:
        train_config["bench_start"] = args.bench_start
    elif "bench_start" not in train_config:
        train_config["bench_start"] = 10
    train_config["data_layer_params"]["shuffle"] = False
    deco_print("New benchmarking config: max_steps={} shuffle=False".format(args.bench_steps))


def resolve_initializer(params, model_params, who):
    """Plugin params -> the engine's initializer name.  The reference passes `initializer(**initializer_params)`
    to tf.layers (encoder.py:60-72 / decoder.py: falls back to the model's `initializer`); without one tf.layers
    uses glorot_uniform, which is xavier_initializer(uniform=True).  Anything that is not Xavier raises."""
    import tensorflow as tf
    init = params.get("initializer", model_params.get("initializer"))
    ip = params.get("initializer_params", model_params.get("initializer_params")) or {}
    if "initializer" not in params and "initializer" in model_params:
        ip = model_params.get("initializer_params") or {}
    if init is None:
        return "xavier_uniform"
    if init is tf.contrib.layers.xavier_initializer:
        extra = set(ip) - {"uniform", "seed", "dtype"}
        if extra:
            raise NotImplementedError("%s: xavier_initializer arguments %r are not built" % (who, sorted(extra)))
        return "xavier_uniform" if ip.get("uniform", True) else "xavier_truncnorm"
    raise NotImplementedError("%s: initializer %r is not built (tf.contrib.layers.xavier_initializer is)" % (who, init))


def create_model(args, base_config, config_module, base_model, hvd=None, checkpoint=None):
    """utils.py:791-882: merge mode-specific params, apply --benchmark rewrites, build + compile."""
    train_config = copy.deepcopy(base_config)
    eval_config = copy.deepcopy(base_config)
    infer_config = copy.deepcopy(base_config)
    if args.mode in ("train", "train_eval"):
        if "train_params" in config_module:
            nested_update(train_config, copy.deepcopy(config_module["train_params"]))
    if args.mode in ("eval", "train_eval"):
        if "eval_params" in config_module:
            nested_update(eval_config, copy.deepcopy(config_module["eval_params"]))
    if args.mode in ("infer", "interactive_infer"):
        key = "infer_params" if args.mode == "infer" else "interactive_infer_params"
        if key in config_module:
            nested_update(infer_config, copy.deepcopy(config_module[key]))
    if args.benchmark:
        adjust_for_benchmark(train_config, args)
        args.mode = "train"
    if args.mode == "train_eval":
        train_model = base_model(params=train_config, mode="train", hvd=hvd)
        train_model.compile(checkpoint=checkpoint)   # --continue_learning: resume from the found checkpoint
        eval_model = base_model(params=eval_config, mode="eval", hvd=hvd)
        eval_model.compile(force_var_reuse=True, share_with=train_model)
        return [train_model, eval_model]
    if args.mode == "train":
        model = base_model(params=train_config, mode="train", hvd=hvd)
        model.compile(force_var_reuse=False, checkpoint=checkpoint)
    elif args.mode == "eval":
        # funcs.evaluate restores the checkpoint check_logdir found before it runs (utils/funcs.py:205-218)
        model = base_model(params=eval_config, mode="eval", hvd=hvd)
        model.compile(force_var_reuse=False, checkpoint=checkpoint)
    else:
        model = base_model(params=infer_config, mode=args.mode, hvd=hvd)
        model.compile(checkpoint=checkpoint)
    return model
