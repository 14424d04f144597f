"""FullyConnectedTimeDecoder / FullyConnectedCTCDecoder (open_seq2seq/decoders/fc_decoders.py:73-253).
Projection and greedy CTC decoding run in libos2s_b200 (os2s_fc_fwd, os2s_ctc_greedy); the
language-model beam-search branch is the reference's CPU custom op and is out of scope."""
from .decoder import Decoder


class FullyConnectedTimeDecoder(Decoder):
    @staticmethod
    def get_required_params():
        return dict(Decoder.get_required_params(), **{"tgt_vocab_size": int})

    @staticmethod
    def get_optional_params():
        return dict(Decoder.get_optional_params(), **{
            "logits_to_outputs_func": None,
            "infer_logits_to_pickle": bool,
        })

    def __init__(self, params, model, name="fully_connected_time_decoder", mode="train"):
        super(FullyConnectedTimeDecoder, self).__init__(params, model, name, mode)

    def _decode(self, input_dict):
        """{'encoder_output': {'outputs': [B,T,H], 'src_length': [B]}} ->
        {'logits': time-major [T,B,V] view, 'outputs': [...], 'src_length': [B]} (fc_decoders.py:105-158)."""
        enc = input_dict["encoder_output"]
        eng = self._model.engine
        logits_btv = eng.forward_decoder()
        logits = logits_btv.transpose(0, 1)  # time-major view, no copy
        outputs = [logits]
        f = self.params.get("logits_to_outputs_func")
        if self._mode == "infer" and self.params.get("infer_logits_to_pickle"):
            # fc_decoders.py:146-147 + models/speech2text.py:214-217: the logits themselves, batch-major, are
            # the model outputs (dumped by Speech2Text.finalize_inference for scripts/decode.py)
            outputs = [logits_btv]
        elif f is not None:
            outputs = f(logits, input_dict)
        return {"outputs": outputs, "logits": logits, "src_length": enc["src_length"]}


class FullyConnectedCTCDecoder(FullyConnectedTimeDecoder):
    @staticmethod
    def get_required_params():
        return FullyConnectedTimeDecoder.get_required_params()

    @staticmethod
    def get_optional_params():
        return dict(FullyConnectedTimeDecoder.get_optional_params(), **{
            "use_language_model": bool, "decoder_library_path": str, "beam_width": int,
            "alpha": float, "beta": float, "trie_weight": float, "lm_path": str,
            "trie_path": str, "alphabet_config_path": str,
        })

    def __init__(self, params, model, name="fully_connected_ctc_decoder", mode="train"):
        super(FullyConnectedCTCDecoder, self).__init__(params, model, name, mode)
        if self.params.get("use_language_model", False):
            raise NotImplementedError(
                "FullyConnectedCTCDecoder: use_language_model=True needs the reference's CPU KenLM custom op "
                "(ctc_decoder_with_lm/), which is out of scope; use greedy decoding")
        self.params["logits_to_outputs_func"] = self._greedy

    def _greedy(self, logits, input_dict):
        """decode_without_lm (fc_decoders.py:244-251): sparse-like (tokens [B,T], lens [B])."""
        toks, lens = self._model.engine.greedy_decode()
        return [(toks, lens)]
