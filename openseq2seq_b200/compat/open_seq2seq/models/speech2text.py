"""Speech2Text model: vocabulary plumbing into the decoder, WER evaluation, frame counter
(open_seq2seq/models/speech2text.py:51-360)."""
import numpy as np

from open_seq2seq.utils.utils import deco_print
from .encoder_decoder import EncoderDecoderModel


def levenshtein(a, b):
    """Edit distance between two sequences (models/speech2text.py:51-71 semantics)."""
    n, m = len(a), len(b)
    if n > m:
        a, b = b, a
        n, m = m, n
    current = list(range(n + 1))
    for i in range(1, m + 1):
        previous, current = current, [i] + [0] * n
        for j in range(1, n + 1):
            add, delete = previous[j] + 1, current[j - 1] + 1
            change = previous[j - 1]
            if a[j - 1] != b[i - 1]:
                change += 1
            current[j] = min(add, delete, change)
    return current[n]


class Speech2Text(EncoderDecoderModel):
    def _create_decoder(self):
        data_layer = self.get_data_layer()
        self.params["decoder_params"]["tgt_vocab_size"] = data_layer.params["tgt_vocab_size"]
        self.dump_outputs = self.params["decoder_params"].get("infer_logits_to_pickle", False)
        return super(Speech2Text, self)._create_decoder()

    def _get_num_objects_per_step(self, worker_id=0):
        """Number of input frames in the last batch (models/speech2text.py:356-360)."""
        dl = self.get_data_layer(worker_id)
        return dl.input_tensors["source_tensors"][1].sum()

    def _decode_batch(self, tokens, tok_lens):
        idx2char = self.get_data_layer().params["idx2char"]
        toks = tokens.cpu().numpy()
        lens = tok_lens.cpu().numpy()
        return ["".join(idx2char[int(c)] for c in toks[b, :lens[b]]) for b in range(toks.shape[0])]

    def maybe_print_logs(self, input_values, output_values, training_step):
        y, ylen = input_values["target_tensors"]
        toks, tl = output_values
        idx2char = self.get_data_layer().params["idx2char"]
        y0 = y[0].cpu().numpy()
        true_text = "".join(idx2char[int(c)] for c in y0[:int(ylen[0])])
        pred_text = self._decode_batch(toks[:1], tl[:1])[0]
        sample_wer = levenshtein(true_text.split(), pred_text.split()) / max(len(true_text.split()), 1)
        deco_print("Sample WER: {:.4f}".format(sample_wer), offset=4)
        deco_print("Sample target:     " + true_text, offset=4)
        deco_print("Sample prediction: " + pred_text, offset=4)
        return {"Sample WER": sample_wer}

    def evaluate(self, input_values, output_values):
        """Per batch: (word errors, word count) (models/speech2text.py:257-294)."""
        y, ylen = input_values["target_tensors"]
        toks, tl = output_values
        idx2char = self.get_data_layer().params["idx2char"]
        preds = self._decode_batch(toks, tl)
        yc, yl = y.cpu().numpy(), ylen.cpu().numpy()
        total_err, total_words = 0.0, 0.0
        for b, pred in enumerate(preds):
            truth = "".join(idx2char[int(c)] for c in yc[b, :yl[b]])
            total_err += levenshtein(truth.split(), pred.split())
            total_words += len(truth.split())
        return total_err, total_words

    def finalize_evaluation(self, results_per_batch, training_step=None):
        total_err = sum(r[0] for r in results_per_batch)
        total_words = sum(r[1] for r in results_per_batch)
        wer = total_err / max(total_words, 1.0)
        deco_print("Validation WER:  {:.4f}".format(wer), offset=4)
        return {"Eval WER": wer}

    def infer(self, input_values, output_values):
        if self.dump_outputs:
            # models/speech2text.py:300-304: one [T, V] logits array per utterance
            lg = output_values.float().cpu().numpy()
            return [lg[b] for b in range(lg.shape[0])], input_values["source_ids"]
        toks, tl = output_values
        return self._decode_batch(toks, tl), input_values["source_ids"]

    def finalize_inference(self, results_per_batch, output_file):
        import pandas as pd
        preds, ids = [], []
        for p, i in results_per_batch:
            preds += p
            ids += list(np.asarray(i[0]).reshape(-1))
        order = np.argsort(ids)
        files = [self.get_data_layer()._files[ids[k]][0] for k in order]
        if self.dump_outputs:
            # models/speech2text.py:325-343: {"logits": {wav: [T, V]}, "step_size": stride product x window_stride,
            # "vocab": idx2char}, the input of scripts/decode.py
            import pickle
            scale = 1
            for c in self.encoder.params.get("convnet_layers") or []:
                scale *= c["stride"][0]
            dump_out = {"logits": {f: preds[k] for f, k in zip(files, order)},
                        "step_size": scale * self.get_data_layer().params["window_stride"],
                        "vocab": self.get_data_layer().params["idx2char"]}
            with open(output_file, "wb") as f:
                pickle.dump(dump_out, f, protocol=pickle.HIGHEST_PROTOCOL)
            return
        pd.DataFrame({"wav_filename": files, "predicted_transcript": [preds[k] for k in order]},
                     columns=["wav_filename", "predicted_transcript"]).to_csv(output_file, index=False)
