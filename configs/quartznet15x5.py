"""QuartzNet 15x5 (separable convolutions, NovoGrad + cosine decay): the hyper-parameters of the reference's
example_configs/speech2text/quartznet15x5_LibriSpeech.py, with the block table generated instead of spelled
out and the dataset replaced by in-memory synthetic utterances (no dataset exists offline).
tests/test_compat_config.py checks, where the reference checkout is available, that this file and the
reference's config describe the same model."""
import os

import tensorflow as tf
from open_seq2seq.models import Speech2Text
from open_seq2seq.encoders import TDNNEncoder
from open_seq2seq.decoders import FullyConnectedCTCDecoder
from open_seq2seq.data.speech2text.speech2text import Speech2TextDataLayer
from open_seq2seq.losses import CTCLoss
from open_seq2seq.optimizers.lr_policies import cosine_decay
from open_seq2seq.optimizers.novograd import NovoGrad

_VOCAB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vocab.txt")
residual_dense = False

# (kernel, channels) of the fifteen 5-layer blocks (B1..B5, each repeated three times)
_BLOCKS = [(33, 256)] * 3 + [(39, 256)] * 3 + [(51, 512)] * 3 + [(63, 512)] * 3 + [(75, 512)] * 3


def _sep(k, c, repeat=1, stride=1, dilation=1, residual=False):
    d = {"type": "sep_conv1d", "repeat": repeat, "kernel_size": [k], "stride": [stride], "num_channels": c,
         "padding": "SAME", "dilation": [dilation]}
    if residual:
        d["residual"] = True
        d["residual_dense"] = residual_dense
    return d


convnet_layers = ([_sep(33, 256, stride=2)] +
                  [_sep(k, c, repeat=5, residual=True) for (k, c) in _BLOCKS] +
                  [_sep(87, 512, dilation=2, residual=True),
                   {"type": "conv1d", "repeat": 1, "kernel_size": [1], "stride": [1], "num_channels": 1024,
                    "padding": "SAME", "dilation": [1]}])

base_model = Speech2Text

base_params = {
    "random_seed": 0,
    "use_horovod": True,
    "num_epochs": 400,
    "num_gpus": 8,
    "batch_size_per_gpu": 32,
    "iter_size": 1,
    "save_summaries_steps": 100,
    "print_loss_steps": 10,
    "print_samples_steps": 2200,
    "eval_steps": 2200,
    "save_checkpoint_steps": 1100,
    "logdir": "jasper_log_folder",
    "num_checkpoints": 2,
    "optimizer": NovoGrad,
    "optimizer_params": {"beta1": 0.95, "beta2": 0.5, "epsilon": 1e-08, "weight_decay": 0.001,
                         "grad_averaging": False},
    "lr_policy": cosine_decay,
    "lr_policy_params": {"learning_rate": 0.01, "min_lr": 0.0, "warmup_steps": 1000},
    "dtype": tf.float32,
    "summaries": ["learning_rate", "variables", "gradients", "larc_summaries", "variable_norm",
                  "gradient_norm", "global_gradient_norm"],
    "encoder": TDNNEncoder,
    "encoder_params": {
        "convnet_layers": convnet_layers,
        "dropout_keep_prob": 1.0,
        "initializer": tf.contrib.layers.xavier_initializer,
        "initializer_params": {"uniform": False},
        "normalization": "batch_norm",
        "activation_fn": tf.nn.relu,
        "data_format": "channels_last",
        "use_conv_mask": True,
    },
    "decoder": FullyConnectedCTCDecoder,
    "decoder_params": {
        "initializer": tf.contrib.layers.xavier_initializer,
        "use_language_model": False,
        "infer_logits_to_pickle": False,
    },
    "loss": CTCLoss,
    "loss_params": {},
    "data_layer": Speech2TextDataLayer,
    "data_layer_params": {
        "num_audio_features": 64,
        "input_type": "logfbank",
        "vocab_file": _VOCAB,
        "norm_per_feature": True,
        "window": "hanning",
        "precompute_mel_basis": True,
        "sample_freq": 16000,
        "pad_to": 16,
        "dither": 1e-5,
        "backend": "librosa",
    },
}

train_params = {
    "data_layer": Speech2TextDataLayer,
    "data_layer_params": {
        "augmentation": {"n_freq_mask": 2, "n_time_mask": 2, "width_freq_mask": 6, "width_time_mask": 6},
        "dataset_files": ["synthetic:64:15.0:1234"],
        "max_duration": 16.7,
        "shuffle": True,
    },
}

eval_params = {
    "data_layer": Speech2TextDataLayer,
    "data_layer_params": {"dataset_files": ["synthetic:32:15.0:4321"], "shuffle": False},
}

infer_params = {
    "data_layer": Speech2TextDataLayer,
    "data_layer_params": {"dataset_files": ["synthetic:32:15.0:4321"], "shuffle": False},
}
