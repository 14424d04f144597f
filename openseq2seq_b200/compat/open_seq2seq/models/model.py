"""Model base class: params schema, seeding, data-layer fan-out, step bookkeeping
(open_seq2seq/models/model.py:25-979), re-hosted on JasperEngine.

One process drives one GPU (the reference's Horovod mode, models/model.py:428-467); `hvd` may be
None (single GPU) or an object with size()/rank()/local_rank() (see run.py's TorchDistHvd, a
torch.distributed/NCCL stand-in for horovod.tensorflow).  The multi-tower single-process mode is
out of scope."""
import abc
import copy

import numpy as np
import tensorflow as tf

from open_seq2seq.utils.utils import check_params


class Model(metaclass=abc.ABCMeta):
    @staticmethod
    def get_required_params():
        return {"use_horovod": bool, "batch_size_per_gpu": int, "data_layer": None}

    @staticmethod
    def get_optional_params():
        return {
            "logdir": str, "num_gpus": int, "gpu_ids": list, "load_model": str,
            "save_summaries_steps": None, "print_loss_steps": None, "print_samples_steps": None,
            "print_bench_info_steps": None, "save_checkpoint_steps": None, "num_checkpoints": int,
            "restore_best_checkpoint": bool, "eval_steps": int, "finetune": bool,
            "eval_batch_size_per_gpu": int, "hooks": list, "random_seed": int, "num_epochs": int,
            "max_steps": int, "bench_start": int, "data_layer_params": dict, "optimizer": None,
            "optimizer_params": dict, "freeze_variables_regex": None, "initializer": None,
            "initializer_params": dict, "regularizer": None, "regularizer_params": dict,
            "dtype": [tf.float16, tf.float32, "mixed"], "lr_policy": None, "lr_policy_params": dict,
            "max_grad_norm": float, "larc_params": dict, "loss_scaling": None, "loss_scaling_params": dict,
            "summaries": list, "iter_size": int, "lm_vocab_file": str, "processed_data_folder": str,
            "use_trt": bool, "trt_precision_mode": str, "trt_max_workspace_size_bytes": int,
            "trt_minimum_segment_size": int, "trt_is_dynamic_op": bool, "trt_maximum_cached_engines": int,
            "use_xla_jit": bool,
        }

    def __init__(self, params, mode="train", hvd=None):
        check_params(params, self.get_required_params(), self.get_optional_params())
        self._params = copy.deepcopy(params)
        if self._params.get("max_grad_norm") is not None and self._params.get("larc_params") is not None:
            raise ValueError("LARC and gradient norm clipping should not be used together")
        if mode not in ("train", "infer", "eval", "interactive_infer"):
            raise ValueError("Mode has to be one of ['train', 'infer', 'eval', 'interactive_infer']")
        if "max_steps" in params and "num_epochs" in params:
            raise ValueError("You can't provide both max_steps and num_epochs. Please, remove one of them from the config.")
        if mode == "train":
            if "max_steps" not in params and "num_epochs" not in params:
                raise ValueError("For training mode either max_steps or num_epochs has to be provided")
        self._mode = mode
        self._interactive = mode == "interactive_infer"
        if self._interactive:
            self._mode = "infer"
        p = self._params
        for k in ("save_summaries_steps", "print_loss_steps", "print_samples_steps", "print_bench_info_steps",
                  "save_checkpoint_steps", "restore_best_checkpoint"):
            p.setdefault(k, None if k != "restore_best_checkpoint" else False)
        p["num_checkpoints"] = p.get("num_checkpoints", 5)
        p["finetune"] = p.get("finetune", False)
        p["load_model"] = p.get("load_model", None)
        p["iter_size"] = p.get("iter_size", 1)
        p["loss_scaling"] = p.get("loss_scaling", 1.0)
        p["loss_scaling_params"] = p.get("loss_scaling_params", None)
        p["summaries"] = p.get("summaries", [])
        if "dtype" not in p:
            p["dtype"] = tf.float32
        if int(p["iter_size"]) < 1:
            raise ValueError("iter_size must be >= 1")
        self._hvd = hvd if (p["use_horovod"] and hvd is not None) else None
        self.on_horovod = self._hvd is not None
        self._gpu_ids = [0]
        self.num_gpus = 1
        if self._mode == "eval" and "eval_batch_size_per_gpu" in p:
            p["batch_size_per_gpu"] = p["eval_batch_size_per_gpu"]
        # per-rank seed (models/model.py:309-313)
        rs = int(p.get("random_seed", int(np.random.randint(0, 2 ** 31 - 1))))
        self._seed = rs + (self._hvd.rank() if self.on_horovod else 0)
        np.random.seed(self._seed)
        dl_params = p.get("data_layer_params", {})
        dl_params = copy.deepcopy(dl_params)
        if "lm_vocab_file" in p:
            dl_params["lm_vocab_file"] = p["lm_vocab_file"]
        dl_params["batch_size"] = p["batch_size_per_gpu"]
        dl_params["mode"] = self._mode
        if self._interactive:
            dl_params["interactive"] = True
        workers = self._hvd.size() if self.on_horovod else 1
        wid = self._hvd.rank() if self.on_horovod else 0
        self._data_layer = p["data_layer"](params=dl_params, model=self, num_workers=workers, worker_id=wid)
        self._data_layers = [self._data_layer]
        if self._mode == "train":
            if "max_steps" in p:
                self._last_step = p["max_steps"]
                self._steps_in_epoch = None
            else:
                n = self._data_layer.get_size_in_samples()
                if n is None:
                    raise ValueError("num_epochs needs a data layer that implements get_size_in_samples()")
                self._steps_in_epoch = n // p["batch_size_per_gpu"] // workers // p["iter_size"]
                if self._steps_in_epoch == 0:
                    raise ValueError("Overall batch size is too big for this dataset.")
                self._last_step = p["num_epochs"] * self._steps_in_epoch
        self.engine = None
        self.loss = None
        self.train_op = None
        self.eval_losses = None
        self._outputs = [None]

    @abc.abstractmethod
    def compile(self, force_var_reuse=False, checkpoint=None, share_with=None):
        pass

    def get_data_layer(self, worker_id=0):
        return self._data_layers[worker_id]

    def get_output_tensors(self, worker_id=0):
        return self._outputs[worker_id]

    def get_num_objects_per_step(self, worker_id=0):
        return self._get_num_objects_per_step(worker_id)

    def _get_num_objects_per_step(self, worker_id=0):
        return None

    def evaluate(self, input_values, output_values):
        return []

    def finalize_evaluation(self, results_per_batch, training_step=None):
        return {}

    def infer(self, input_values, output_values):
        return []

    def finalize_inference(self, results_per_batch, output_file):
        pass

    def maybe_print_logs(self, input_values, output_values, training_step):
        return {}

    @property
    def params(self):
        return self._params

    @property
    def steps_in_epoch(self):
        return self._steps_in_epoch

    @property
    def last_step(self):
        return self._last_step

    @property
    def mode(self):
        return self._mode

    @property
    def hvd(self):
        return self._hvd
