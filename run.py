"""Command-line entry with the reference's flags (run.py of NVIDIA/OpenSeq2Seq, :19-101):

  python run.py --config_file=<cfg.py> --mode=train|eval|train_eval|infer [--benchmark ...] [--a/b/c=v]

Multi-GPU: one process per GPU under `python -m torch.distributed.run`; torch.distributed (NCCL)
replaces Horovod/MPI for the three collectives the reference uses (gradient mean, initial parameter
broadcast, scalar gather)."""
from __future__ import print_function

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import openseq2seq_b200.compat as _compat  # noqa: E402

_compat.install()

from open_seq2seq.utils.utils import (check_logdir, create_logdir, create_model, deco_print,  # noqa: E402
                                      get_base_config)
from open_seq2seq.utils.funcs import evaluate, infer, train  # noqa: E402
from openseq2seq_b200.dist import TorchDistHvd  # noqa: E402


def main():
    args, base_config, base_model, config_module = get_base_config(sys.argv[1:])
    hvd = None
    if base_config.get("use_horovod", False) and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        hvd = TorchDistHvd.init()
    elif base_config.get("use_horovod", False):
        hvd = TorchDistHvd.single()
    restore_best = base_config.get("restore_best_checkpoint", False)
    checkpoint = check_logdir(args, base_config, restore_best)
    if args.enable_logs:
        create_logdir(args, base_config)
    if args.mode in ("train", "train_eval") and base_config.get("logdir"):
        os.makedirs(base_config["logdir"], exist_ok=True) if not args.benchmark else None
    if hvd is None or hvd.rank() == 0:
        deco_print("Running in mode {} on {} GPU(s)".format(args.mode, hvd.size() if hvd else 1))
    model = create_model(args, base_config, config_module, base_model, hvd, checkpoint)
    if args.mode == "train_eval":
        train(model[0], eval_model=model[1])
    elif args.mode == "train":
        train(model)   # create_model restored `checkpoint` (--continue_learning) before the first step
    elif args.mode == "eval":
        evaluate(model, checkpoint)
    elif args.mode == "infer":
        infer(model, checkpoint, args.infer_output_file)
    else:
        raise NotImplementedError("interactive_infer is outside the built path")


if __name__ == "__main__":
    main()
