"""CTCLoss (open_seq2seq/losses/ctc_loss.py:19-89): fp32, ignore_longer_outputs_than_inputs,
NaN masking, batch mean.  The alpha/beta lattices and the gradient run on the GPU
(os2s_ctc_loss_fwd_bwd); in train mode the same call also enqueues the whole backward pass."""
import tensorflow as tf

from .loss import Loss


class CTCLoss(Loss):
    @staticmethod
    def get_optional_params():
        return dict(Loss.get_optional_params(), **{"mask_nan": bool})

    def __init__(self, params, model, name="ctc_loss"):
        super(CTCLoss, self).__init__(params, model, name)
        self._mask_nan = self.params.get("mask_nan", True)
        self.params["dtype"] = tf.float32

    def _compute_loss(self, input_dict):
        tgt, tgt_len = input_dict["target_tensors"]
        per_utt = self._model.engine.loss_and_backward(tgt, tgt_len)
        return per_utt.mean()
