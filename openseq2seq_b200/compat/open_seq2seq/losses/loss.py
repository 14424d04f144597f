"""Loss base class (open_seq2seq/losses/loss.py:15-99)."""
import abc
import copy

import tensorflow as tf

from open_seq2seq.utils.utils import check_params


class Loss(metaclass=abc.ABCMeta):
    @staticmethod
    def get_required_params():
        return {}

    @staticmethod
    def get_optional_params():
        return {"dtype": [tf.float16, tf.float32]}

    def __init__(self, params, model, name="loss"):
        check_params(params, self.get_required_params(), self.get_optional_params())
        self._params = copy.deepcopy(params)
        self._model = model
        if "dtype" not in self._params:
            self._params["dtype"] = tf.float32
        self._name = name

    def compute_loss(self, input_dict):
        return self._compute_loss(input_dict)

    @abc.abstractmethod
    def _compute_loss(self, input_dict):
        pass

    @property
    def params(self):
        return self._params

    @property
    def name(self):
        return self._name
