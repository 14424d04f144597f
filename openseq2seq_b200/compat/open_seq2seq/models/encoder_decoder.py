"""EncoderDecoderModel: encoder -> decoder -> loss glue (open_seq2seq/models/encoder_decoder.py:10-190)."""

from open_seq2seq.optimizers.optimizers import optimizer_engine_kwargs
from .model import Model


class EncoderDecoderModel(Model):
    @staticmethod
    def get_required_params():
        return dict(Model.get_required_params(), **{"encoder": None, "decoder": None})

    @staticmethod
    def get_optional_params():
        return dict(Model.get_optional_params(), **{
            "encoder_params": dict, "decoder_params": dict, "loss": None, "loss_params": dict,
        })

    def __init__(self, params, mode="train", hvd=None):
        super(EncoderDecoderModel, self).__init__(params=params, mode=mode, hvd=hvd)
        if "encoder_params" not in self.params:
            self.params["encoder_params"] = {}
        if "decoder_params" not in self.params:
            self.params["decoder_params"] = {}
        if "loss_params" not in self.params:
            self.params["loss_params"] = {}
        self._encoder = self._create_encoder()
        self._decoder = self._create_decoder()
        if self.mode in ("train", "eval"):
            self._loss_computator = self._create_loss()
        else:
            self._loss_computator = None

    def _create_encoder(self):
        params = self.params["encoder_params"]
        return self.params["encoder"](params=params, mode=self.mode, model=self)

    def _create_decoder(self):
        params = self.params["decoder_params"]
        return self.params["decoder"](params=params, mode=self.mode, model=self)

    def _create_loss(self):
        return self.params["loss"](params=self.params["loss_params"], model=self)

    # ------------------------------------------------------------------ compile
    def compile(self, force_var_reuse=False, checkpoint=None, share_with=None):
        """Builds the device state.  `share_with` (an already compiled train model) makes an eval model
        run on the same parameters (the reference's force_var_reuse)."""
        from openseq2seq_b200.engine import JasperEngine
        self._data_layer.build_graph()
        if share_with is not None:
            self.engine = share_with.engine
            self._shared = True
            if hasattr(self._data_layer, "set_feature_dtype"):
                self._data_layer.set_feature_dtype(self.engine.act_dtype)
            return
        self._shared = False
        from open_seq2seq.utils.utils import resolve_initializer
        enc_kw = self._encoder.engine_kwargs()
        enc_kw["decoder_init"] = resolve_initializer(self._decoder.params, self.params, type(self._decoder).__name__)
        dl = self._data_layer.params
        world = self._hvd.size() if self.on_horovod else 1
        opt_kw = {}
        if self.mode == "train":
            opt_kw = optimizer_engine_kwargs(self.params, self.last_step)
        self.engine = JasperEngine(num_features=dl["num_audio_features"],
                                   vocab_size=self._decoder.params["tgt_vocab_size"], opt=opt_kw,
                                   world_size=world, seed=self._seed, **enc_kw)
        if hasattr(self._data_layer, "set_feature_dtype"):
            self._data_layer.set_feature_dtype(self.engine.act_dtype)   # features arrive in the engine's 16-bit format
        from open_seq2seq.utils import checkpoint as ckpt
        if checkpoint is not None:
            ckpt.restore(self.engine, checkpoint)
            self._restored_from = checkpoint
        elif self.mode == "train" and self.params.get("load_model"):
            # utils/funcs.py:117-144: no checkpoint in logdir -> initialise matching variables from load_model
            n = ckpt.restore_partial(self.engine, self.params["load_model"])
            from open_seq2seq.utils.utils import deco_print
            deco_print("load_model: restored %d variables from %s" % (n, self.params["load_model"]))
        if self.on_horovod:
            self._hvd.broadcast_parameters(self.engine)
            if self.mode == "train":
                self.engine.set_comm(self._hvd)

    # -------------------------------------------------------------- one step
    def _forward(self, batch):
        enc_out = self._encoder.encode({"source_tensors": batch["source_tensors"]})
        dec_in = {"encoder_output": enc_out}
        if "target_tensors" in batch:
            dec_in["target_tensors"] = batch["target_tensors"]
        dec_out = self._decoder.decode(dec_in)
        return enc_out, dec_out

    def train_step(self, batch):
        """sess.run(train_op): forward, loss + backward, gradient all-reduce, optimizer step -- issued
        through JasperEngine.train_step so the steady-state step is one CUDA-graph replay.  The plugin
        objects' encode / decode / compute_loss run the same kernels phase by phase (used by eval,
        infer and the parity tests).  Returns (mean loss, number of input frames) as device tensors."""
        feats, lens = batch["source_tensors"]
        y, ylen = batch["target_tensors"]
        per_utt = self.engine.train_step(feats, lens, y, ylen)
        self.loss = per_utt.mean()
        return self.loss, lens.sum()

    def eval_step(self, batch):
        enc_out, dec_out = self._forward(batch)
        loss = None
        if self._loss_computator is not None and "target_tensors" in batch:
            per = self.engine.loss_only(batch["target_tensors"][0], batch["target_tensors"][1])
            loss = per.mean()
        return loss, dec_out

    @property
    def encoder(self):
        return self._encoder

    @property
    def decoder(self):
        return self._decoder

    @property
    def loss_computator(self):
        return self._loss_computator
