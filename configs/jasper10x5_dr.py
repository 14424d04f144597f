"""Jasper 10x5 Dense-Residual, NovoGrad + LARC, mixed precision: the hyper-parameters of the
reference's example_configs/speech2text/jasper10x5_LibriSpeech_nvgrad.py, generated from the block
table instead of spelled out, with the dataset replaced by in-memory
synthetic utterances (no dataset exists offline).  tests/test_compat_config.py checks, where the
reference checkout is available, that this file and the reference's config describe the same model.
"""
import tensorflow as tf
from open_seq2seq.models import Speech2Text
from open_seq2seq.encoders import TDNNEncoder
from open_seq2seq.decoders import FullyConnectedCTCDecoder
from open_seq2seq.data.speech2text.speech2text import Speech2TextDataLayer
from open_seq2seq.losses import CTCLoss
from open_seq2seq.optimizers.lr_policies import poly_decay
from open_seq2seq.optimizers.novograd import NovoGrad
import os

_VOCAB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vocab.txt")

# (kernel, channels, dropout keep) of the ten 5-layer dense-residual blocks
_BLOCKS = [(11, 256, 0.8), (11, 256, 0.8), (13, 384, 0.8), (13, 384, 0.8), (17, 512, 0.8),
           (17, 512, 0.8), (21, 640, 0.7), (21, 640, 0.7), (25, 768, 0.7), (25, 768, 0.7)]


def _layer(k, c, keep, repeat=1, stride=1, dilation=1, residual=False):
    d = {"type": "conv1d", "repeat": repeat, "kernel_size": [k], "stride": [stride], "num_channels": c,
         "padding": "SAME", "dilation": [dilation], "dropout_keep_prob": keep}
    if residual:
        d["residual"] = True
        d["residual_dense"] = True
    return d


convnet_layers = ([_layer(11, 256, 0.8, stride=2)] +
                  [_layer(k, c, keep, repeat=5, residual=True) for (k, c, keep) in _BLOCKS] +
                  [_layer(29, 896, 0.6, dilation=2), _layer(1, 1024, 0.6)])

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
    "optimizer_params": {"beta1": 0.95, "beta2": 0.98, "epsilon": 1e-08, "weight_decay": 0.001,
                         "grad_averaging": False},
    "lr_policy": poly_decay,
    "lr_policy_params": {"learning_rate": 0.02, "min_lr": 1e-5, "power": 2.0},
    "larc_params": {"larc_eta": 0.001},
    "dtype": "mixed",
    "loss_scaling": "Backoff",
    "summaries": ["learning_rate", "variables", "gradients", "larc_summaries", "variable_norm",
                  "gradient_norm", "global_gradient_norm"],
    "encoder": TDNNEncoder,
    "encoder_params": {
        "convnet_layers": convnet_layers,
        "dropout_keep_prob": 0.7,
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
        "augmentation": {
            "speed_perturbation_ratio": [0.9, 1., 1.1],
        },
        # 16 kHz / 15 s LibriSpeech-shaped synthetic utterances
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
