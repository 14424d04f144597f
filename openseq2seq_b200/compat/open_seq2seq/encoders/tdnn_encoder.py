"""TDNNEncoder (Jasper / Wave2Letter+ stack).  Params schema of
open_seq2seq/encoders/tdnn_encoder.py:19-46; the computation runs in JasperEngine."""
import tensorflow as tf

from .encoder import Encoder


class TDNNEncoder(Encoder):
    @staticmethod
    def get_required_params():
        return dict(Encoder.get_required_params(), **{
            "dropout_keep_prob": float,
            "convnet_layers": list,
            "activation_fn": None,
        })

    @staticmethod
    def get_optional_params():
        return dict(Encoder.get_optional_params(), **{
            "data_format": ["channels_first", "channels_last"],
            "normalization": [None, "batch_norm", "layer_norm", "instance_norm"],
            "bn_momentum": float,
            "bn_epsilon": float,
            "use_conv_mask": bool,
            "drop_block_prob": float,
            "drop_block_index": int,
        })

    def __init__(self, params, model, name="w2l_encoder", mode="train"):
        super(TDNNEncoder, self).__init__(params, model, name, mode)

    def engine_kwargs(self):
        """What JasperEngine needs from this plugin's params (tdnn_encoder.py:127-179)."""
        p = self.params
        if p.get("normalization", "batch_norm") != "batch_norm":
            raise NotImplementedError("TDNNEncoder: only normalization='batch_norm' has B200 kernels")
        if p.get("drop_block_prob", 0.0) > 0:
            raise NotImplementedError("TDNNEncoder: stochastic block drop is not built")
        apply_relu, clip = tf.resolve_activation(p["activation_fn"])
        if not apply_relu:
            raise NotImplementedError("TDNNEncoder: identity activation is not built")
        from open_seq2seq.utils.utils import resolve_initializer
        init = resolve_initializer(p, self._model.params if self._model is not None else {}, "TDNNEncoder")
        return dict(encoder_init=init, convnet_layers=p["convnet_layers"], bn_momentum=p.get("bn_momentum", 0.90),
                    bn_epsilon=p.get("bn_epsilon", 1e-3), use_conv_mask=p.get("use_conv_mask", False),
                    training=(self._mode == "train"), dropout_keep_default=p["dropout_keep_prob"],
                    relu_clip=clip)

    def _encode(self, input_dict):
        feats, lens = input_dict["source_tensors"]
        eng = self._model.engine
        out, out_len = eng.forward_encoder(feats, lens)
        return {"outputs": out, "src_length": out_len}
