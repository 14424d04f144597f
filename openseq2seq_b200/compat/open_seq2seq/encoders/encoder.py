"""Encoder base class: same constructor signature, params validation and `encode` wrapper as
open_seq2seq/encoders/encoder.py:16-150 (params deep-copied, check_params, dtype defaulting)."""
import abc
import copy

import tensorflow as tf

from open_seq2seq.utils.utils import check_params


class Encoder(metaclass=abc.ABCMeta):
    @staticmethod
    def get_required_params():
        return {}

    @staticmethod
    def get_optional_params():
        return {
            "regularizer": None, "regularizer_params": dict,
            "initializer": None, "initializer_params": dict,
            "dtype": [tf.float32, tf.float16, "mixed"],
        }

    def __init__(self, params, model, name="encoder", mode="train"):
        check_params(params, self.get_required_params(), self.get_optional_params())
        self._params = copy.deepcopy(params)
        self._model = model
        if "dtype" not in self._params:
            self._params["dtype"] = model.params["dtype"] if model else tf.float32
        self._name = name
        self._mode = mode
        self._compiled = False

    def encode(self, input_dict):
        """input_dict['source_tensors'] = [features bf16 [B,T,F] (device), lengths int32 [B]] ->
        {'outputs': [B,T',H], 'src_length': [B]} (open_seq2seq/encoders/encoder.py:95-138)."""
        return self._encode(input_dict)

    @abc.abstractmethod
    def _encode(self, input_dict):
        pass

    @property
    def params(self):
        return self._params

    @property
    def mode(self):
        return self._mode

    @property
    def name(self):
        return self._name
