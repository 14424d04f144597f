"""DataLayer base class (open_seq2seq/data/data_layer.py:16-157)."""
import abc
import copy

import tensorflow as tf

from open_seq2seq.utils.utils import check_params


class DataLayer(metaclass=abc.ABCMeta):
    @staticmethod
    def get_required_params():
        return {"mode": ["train", "eval", "infer"]}

    @staticmethod
    def get_optional_params():
        return {
            "batch_size": int, "shuffle": bool, "repeat": bool,
            "dtype": [tf.float32, tf.float16], "interactive": bool,
            "cache_features": bool, "cache_format": str, "cache_regenerate": bool,
        }

    def __init__(self, params, model, num_workers, worker_id):
        check_params(params, self.get_required_params(), self.get_optional_params())
        self._params = copy.deepcopy(params)
        self._model = model
        if "dtype" not in self._params:
            self._params["dtype"] = tf.float32
        if "shuffle" not in params:
            self._params["shuffle"] = (self._params["mode"] == "train")
        if self._params["mode"] != "train" and self._params["shuffle"]:
            raise ValueError("Shuffle should not be performed in %s mode" % self._params["mode"])
        self._num_workers = num_workers
        self._worker_id = worker_id

    @property
    def params(self):
        return self._params

    @abc.abstractmethod
    def build_graph(self):
        pass

    @property
    @abc.abstractmethod
    def iterator(self):
        pass

    @property
    @abc.abstractmethod
    def input_tensors(self):
        pass

    def get_size_in_samples(self):
        return None
