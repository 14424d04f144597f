"""JasperEngine: the B200 training path behind the OpenSeq2Seq plugin classes.

It owns the device-resident state of one data-parallel replica -- fp32 master parameters, fp32
gradients, momentum, bf16 working copies (natural + transposed layouts), BN statistics, saved
activations -- and drives the hand-written sm_100a kernels of libos2s_b200 through the C ABI
(include/os2s.h).  PyTorch is used only as the device allocator / stream / NCCL host.

What it replaces in the reference (one training step):
  TDNNEncoder._encode                    open_seq2seq/encoders/tdnn_encoder.py:87-265
  conv_bn_actv / conv_bn_res_bn_actv     open_seq2seq/parts/cnns/conv_blocks.py:61-232
  FullyConnectedTimeDecoder._decode      open_seq2seq/decoders/fc_decoders.py:105-158
  CTCLoss._compute_loss                  open_seq2seq/losses/ctc_loss.py:44-89
  optimize_loss + MP wrapper + NovoGrad  open_seq2seq/optimizers/{optimizers,mp_wrapper,novograd}.py
  hvd.allreduce per gradient             open_seq2seq/optimizers/optimizers.py:77-104

All launches of a step are pre-bound once per (batch, time) shape into a flat "plan" of
(function, argument-list) pairs, so the per-step Python cost is one loop over ctypes calls and
the step never synchronises with the host.
"""
import collections
import ctypes
import math
import os

import torch

from . import _lib as L

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_ll = ctypes.c_longlong
_c_u64 = ctypes.c_uint64
_c_size_t = ctypes.c_size_t
_vp = ctypes.c_void_p


class OptHParams(ctypes.Structure):
    """Mirror of os2s_opt_hparams (include/os2s.h)."""
    _fields_ = [
        ("algo", _c_int), ("beta1", _c_float), ("beta2", _c_float), ("epsilon", _c_float),
        ("weight_decay", _c_float), ("momentum", _c_float), ("grad_averaging", _c_int),
        ("ema_persist", _c_int), ("larc_eta", _c_float), ("larc_eps", _c_float),
        ("larc_min_update", _c_float), ("larc_mode", _c_int), ("lr0", _c_float), ("min_lr", _c_float),
        ("power", _c_float), ("decay_steps", _c_ll), ("begin_decay_at", _c_ll), ("warmup_steps", _c_ll),
        ("use_loss_scaler", _c_int), ("scale_min", _c_float), ("scale_max", _c_float),
        ("step_factor", _c_float), ("step_window", _c_ll), ("world_size", _c_int),
        ("lr_policy", _c_int), ("decay_rate", _c_float), ("staircase", _c_int),
        ("max_grad_norm", _c_float), ("wb_f16", _c_int),
    ]


def _align(n, a=128):
    return (n + a - 1) // a * a


def same_padding(T_in, K, stride, dilation):
    """tf SAME padding: returns (T_out, pad_left, pad_right)."""
    T_out = -(-T_in // stride)
    total = max((T_out - 1) * stride + (K - 1) * dilation + 1 - T_in, 0)
    return T_out, total // 2, total - total // 2


class ConvLayer(object):
    """One conv+BN(+residual)+act layer of the TDNN stack, in kernel terms."""

    def __init__(self, name, K, stride, dil, c_in, c_out, keep, block, rep, block_end, res_sources,
                 dense):
        self.name, self.K, self.stride, self.dil = name, K, stride, dil
        self.c_in, self.c_out, self.keep = c_in, c_out, keep
        self.block, self.rep, self.block_end = block, rep, block_end
        self.res_sources = res_sources  # indices into the engine's block-input list
        self.dense = dense
        # kernel-level geometry (stride folded into channels), filled by the engine
        self.kK = K
        self.kC_in = c_in
        self.kpad = 0
        self.lead_taps = 0
        # sep_conv1d (QuartzNet / Jasper-Mini): "split" = depthwise kernel on the CUDA cores + 1x1 tcgen05 GEMM,
        # "compose" = equivalent dense kernel D[k,c] * P[c,o] formed on the fly (K = 1, stride 2, narrow inputs)
        self.sep = False
        self.sep_mode = None

    @property
    def wname(self):
        """Name of the 16-bit kernel the MAIN conv of this layer multiplies with."""
        if not self.sep:
            return self.name + "/kernel"
        return self.name + ("/kernel@composed" if self.sep_mode == "compose" else "/pointwise_kernel")


class JasperEngine(object):
    def __init__(self, convnet_layers, num_features, vocab_size, device="cuda", bn_momentum=0.9,
                 bn_epsilon=1e-3, use_conv_mask=True, training=True, dropout_keep_default=1.0,
                 opt=None, world_size=1, seed=0, relu_clip=0.0, act_dtype=None, conv_dtype=None,
                 encoder_init="xavier_truncnorm", decoder_init="xavier_uniform"):
        """act_dtype: format of every 16-bit tensor of the path (layer inputs / outputs, weight working
        copies, activation gradients): "bf16" (default) or "fp16" (the reference's "mixed" mode; one switch
        because tcgen05 kind::f16 needs both MMA operands in the same format); conv_dtype: storage format of
        the conv outputs (the BN inputs), "fp16" (default) or "fp32".  Environment defaults: OS2S_ACT_DTYPE,
        OS2S_CONV_DTYPE."""
        self.lib = L.load()
        for k in (encoder_init, decoder_init):
            if k not in ("xavier_truncnorm", "xavier_uniform"):
                raise ValueError("JasperEngine: initializer %r is not built (xavier_truncnorm | xavier_uniform)" % (k,))
        self.encoder_init, self.decoder_init = encoder_init, decoder_init
        act_dtype = act_dtype or os.environ.get("OS2S_ACT_DTYPE", "bf16")
        conv_dtype = conv_dtype or os.environ.get("OS2S_CONV_DTYPE", "fp16")
        if act_dtype not in ("bf16", "fp16") or conv_dtype not in ("fp16", "fp32"):
            raise ValueError("JasperEngine: act_dtype is 'bf16' | 'fp16', conv_dtype is 'fp16' | 'fp32'")
        self.act_dtype, self.conv_dtype = act_dtype, conv_dtype
        self.act_torch = torch.float16 if act_dtype == "fp16" else torch.bfloat16
        self.conv_torch = torch.float32 if conv_dtype == "fp32" else torch.float16
        # OS2S_HALF_F16 | OS2S_CONV_F32 (include/os2s.h)
        self.dtypes = (1 if act_dtype == "fp16" else 0) | (2 if conv_dtype == "fp32" else 0)
        self.conv_out_mode = 1 if conv_dtype == "fp32" else 3
        self.grad_out_mode = 4 if act_dtype == "fp16" else 0   # OS2S_OUT_F16_GRAD | OS2S_OUT_BF16
        self.device = torch.device(device)
        self.F = num_features
        self.V = vocab_size
        self.bn_momentum = float(bn_momentum)
        self.bn_eps = float(bn_epsilon)
        self.use_conv_mask = use_conv_mask
        self.training = training
        self.world_size = world_size
        self.seed = seed
        self.relu_clip = float(relu_clip)
        self.step_count = 0
        self._ws = collections.OrderedDict()
        self._profile = None
        # OS2S_NO_COMM=1: timing diagnosis only -- the gradient all-reduce is not issued (ranks diverge)
        self._suppress_comm = os.environ.get("OS2S_NO_COMM", "0") == "1"
        # replay the whole step as one CUDA graph after 2 eager steps (OS2S_CUDA_GRAPH=0 disables)
        self.use_cuda_graph = os.environ.get("OS2S_CUDA_GRAPH", "1") != "0"
        # with a communicator attached (N > 1) the step runs from the eager launch plan unless
        # OS2S_GRAPH_DIST=1 asks for NCCL collectives to be captured into the graph as well
        self.graph_with_comm = os.environ.get("OS2S_GRAPH_DIST", "0") == "1"
        # the peer-memory exchange is plain kernels + copies with device-resident flag counters, so the whole
        # multi-GPU step can replay from a CUDA graph (ranks may capture at different steps; the protocol is the
        # same).  Measured at N = 2 only (profiles/r02_peer_graph_n2.jsonl): on by default there, elsewhere with
        # OS2S_PEER_GRAPH=1 (set_comm decides)
        self.graph_with_peer = os.environ.get("OS2S_PEER_GRAPH", "") == "1"
        # weight-gradient kernels run on an auxiliary stream: wgrad(l) (tensor-bound, not on the critical
        # path) overlaps bn_bwd(l-1) (HBM-bound), which co-resides on the SMs (OS2S_OVERLAP_WGRAD=0 disables)
        self.overlap_wgrad = os.environ.get("OS2S_OVERLAP_WGRAD", "1") != "0"
        # BN-backward reductions of plain (single-branch) layers are accumulated in the epilogue of the
        # data-gradient kernel that produces their dA (OS2S_FUSE_BNRED=0 keeps the separate reduction pass)
        self.fuse_bn_reduce = os.environ.get("OS2S_FUSE_BNRED", "1") != "0"
        self.fuse_bn_min_channels = 384
        self._aux = None
        self.comm = None            # object with allreduce_(tensor) (openseq2seq_b200.dist.TorchDistHvd)
        self.bucket_bytes = int(float(os.environ.get("OS2S_BUCKET_MB", "128")) * (1 << 20))
        self.tail_bucket_bytes = int(float(os.environ.get("OS2S_TAIL_BUCKET_MB", "16")) * (1 << 20))
        self._side = None
        self.peer = None
        self._build_layers(convnet_layers, dropout_keep_default)
        self._alloc_params()
        self.set_optimizer(**(opt or {}))

    # ------------------------------------------------------------------ topology
    def _build_layers(self, cfg, keep_default):
        # Channel counts are padded to multiples of 128 (the tcgen05 tiles need C % 64 == 0, the weight
        # gradient C_in % 128 == 0): a layer's c_in / c_out are the PHYSICAL widths every kernel sees, lc_in /
        # lc_out the reference's.  Pad channels are exact zeros everywhere (zero kernel rows / columns, gamma =
        # beta = 0), so norms, gradients and statistics of the real channels are untouched.
        pad128 = lambda c: -(-c // 128) * 128
        layers = []
        self.block_inputs = []  # (physical channels, first consumer layer, logical channels) of every residual source
        first_stride = cfg[0]["stride"][0] if cfg else 1
        self.Fp = (-(-self.F // 64) * 64) if first_stride > 1 else pad128(self.F)
        c_in, lc_in = self.Fp, self.F
        res_list = []
        for bi, lc in enumerate(cfg):
            ltype = lc.get("type", "conv1d")
            if ltype not in ("conv1d", "sep_conv1d"):
                raise ValueError("JasperEngine: 'conv1d' and 'sep_conv1d' layers are built (got %r)" % (ltype,))
            if lc.get("padding", "SAME") != "SAME":
                raise ValueError("JasperEngine: only SAME padding is built")
            K = lc["kernel_size"][0]
            stride = lc["stride"][0]
            dil = lc["dilation"][0]
            lc_out = lc["num_channels"]
            c_out = pad128(lc_out)
            keep = lc.get("dropout_keep_prob", keep_default) if self.training else 1.0
            residual = lc.get("residual", False)
            dense = lc.get("residual_dense", False)
            sources = []
            if residual:
                self.block_inputs.append((c_in, len(layers), lc_in))
                idx = len(self.block_inputs) - 1
                if dense:
                    res_list.append(idx)
                    sources = list(res_list)
                else:
                    sources = [idx]
            if stride > 1 and (bi != 0 or lc["repeat"] != 1):
                raise ValueError("JasperEngine: stride > 1 is only built for the first layer")
            for ri in range(lc["repeat"]):
                end = residual and ri == lc["repeat"] - 1
                lyr = ConvLayer("conv%d%d" % (bi + 1, ri + 1), K, stride, dil, c_in, c_out, keep, bi, ri,
                                end, sources if end else [], dense)
                if ltype == "sep_conv1d":
                    lyr.sep = True
                    lyr.sep_mode = "compose" if (K == 1 or stride > 1 or c_in % 128 != 0 or K > 96) else "split"
                lyr.lc_in, lyr.lc_out = lc_in, lc_out
                layers.append(lyr)
                c_in, lc_in = c_out, lc_out
        self.layers = layers
        self.H, self.Hl = c_in, lc_in
        # which layers' OUTPUT is a residual source (needs an fp32 gradient accumulator)
        self.src_of_layer_output = {}
        for idx, (c, first_layer, _lc) in enumerate(self.block_inputs):
            if first_layer == 0:
                raise ValueError("JasperEngine: a residual block cannot be the first layer")
            self.src_of_layer_output[first_layer - 1] = idx
        # Dense-residual 1x1 convolutions are computed SOURCE-major: all consumers (block-end layers) of
        # source j share the input A_j, so their kernels W_{b,j} [C_j, C_b] are laid side by side and
        # Y_{.,j} = A_j @ [W_{b1,j} | W_{b2,j} | ...] is ONE GEMM with N = sum_b C_b (likewise one
        # data-gradient GEMM with K = sum_b C_b and one weight-gradient GEMM per source).
        # res_groups[j] = {"cj", "ntot", "consumers": [(layer index, branch position, C_b, first column)]}
        self.res_groups = []
        for j, (cj, first_layer, lcj) in enumerate(self.block_inputs):
            cons, col = [], 0
            for li, l in enumerate(layers):
                for n, jj in enumerate(l.res_sources):
                    if jj == j:
                        cons.append((li, n, l.c_out, col))
                        col += l.c_out
            self.res_groups.append({"cj": cj, "lcj": lcj, "ntot": col, "consumers": cons, "first_layer": first_layer})
        self.res_col = {}  # (layer index, branch position) -> (source, first column)
        for j, g in enumerate(self.res_groups):
            for (li, n, cb, col) in g["consumers"]:
                self.res_col[(li, n)] = (j, col)

    @staticmethod
    def var_scope_name(name):
        """The reference's full variable name of a parameter: encoder variables live
        under ForwardPass/w2l_encoder/, the decoder's dense layer under ForwardPass/fully_connected_ctc_decoder/."""
        if name.startswith("fc/"):
            return "ForwardPass/fully_connected_ctc_decoder/fully_connected/" + name[3:]
        return "ForwardPass/w2l_encoder/" + name

    @staticmethod
    def res_name(lyr, n):
        return (lyr.name + "/res_%d" % n) if lyr.dense else (lyr.name + "/res")

    @staticmethod
    def res_wname(lyr, n):
        """16-bit 1x1 kernel of residual branch n: a sep_conv1d block builds its residual convs with
        tf.layers.separable_conv1d(kernel_size=1) too (conv_blocks.py:79-85) -> composed per-channel scale x 1x1."""
        return JasperEngine.res_name(lyr, n) + ("/kernel@composed" if lyr.sep else "/kernel")

    @staticmethod
    def res_bn_name(lyr, n):
        return (lyr.name + "/res_bn_%d" % n) if lyr.dense else (lyr.name + "/res_bn")

    # ---------------------------------------------------------------- parameters
    def _alloc_params(self):
        """Flat fp32 master / grad / momentum buffers + bf16 copies; names follow SURVEY.md App. B."""
        specs = []  # (name, shape, kind, layer)

        def add(name, shape, kind, layer=None, store_shape=None):
            """shape: the reference's (logical) shape; store_shape: the zero-padded physical one."""
            specs.append({"name": name, "shape": tuple(shape), "kind": kind, "layer": layer,
                          "store_shape": tuple(store_shape or shape)})

        for lyr in self.layers:
            if lyr.stride > 1:
                # fold the stride into channels: x viewed [B, T/s, s*C]; taps regrouped (see engine docs)
                if lyr.stride != 2:
                    raise ValueError("JasperEngine: only stride 2 is built")
                lyr.fold = True
            else:
                lyr.fold = False
            ci, co, pci, pco = lyr.lc_in, lyr.lc_out, lyr.c_in, lyr.c_out
            if not lyr.sep:
                add(lyr.name + "/kernel", (lyr.K, ci, co), "conv", lyr, (lyr.K, pci, pco))
            else:
                # tf.layers.separable_conv1d variables: depthwise_kernel [K, C_in, 1], pointwise_kernel [1, C_in, C_out]
                add(lyr.name + "/depthwise_kernel", (lyr.K, ci, 1), "dw", lyr, (lyr.K, pci, 1))
                if lyr.sep_mode == "split":
                    add(lyr.name + "/pointwise_kernel", (1, ci, co), "conv", lyr, (1, pci, pco))
                else:
                    add(lyr.name + "/pointwise_kernel", (1, ci, co), "pw", lyr, (1, pci, pco))
                    # the composed dense kernel: a frozen pseudo-variable (16-bit slot + dense-gradient slot)
                    add(lyr.name + "/kernel@composed", (lyr.K, ci, co), "conv", lyr, (lyr.K, pci, pco))
            add(lyr.name + "/bn/gamma", (co,), "gamma", lyr, (pco,))
            add(lyr.name + "/bn/beta", (co,), "beta", lyr, (pco,))
            for n, j in enumerate(lyr.res_sources):
                bn = (lyr.name + "/res_bn_%d" % n) if lyr.dense else (lyr.name + "/res_bn")
                add(bn + "/gamma", (co,), "gamma", lyr, (pco,))
                add(bn + "/beta", (co,), "beta", lyr, (pco,))
            # the 1x1 residual kernels of every consumer of the source that feeds THIS layer: their
            # (merged) weight gradient is produced when backward reaches this layer, so they sit in its
            # region of the flat buffers (gradient buckets are cut by layer, last layer first)
            li = self.layers.index(lyr)
            for j, g in enumerate(self.res_groups):
                if g["first_layer"] != li:
                    continue
                for (lc, n, cb, col) in g["consumers"]:
                    cons = self.layers[lc]
                    rn = self.res_name(cons, n)
                    lcj, lcb = g["lcj"], cons.lc_out
                    if cons.sep:
                        add(rn + "/depthwise_kernel", (1, lcj, 1), "dw", cons, (1, g["cj"], 1))
                        add(rn + "/pointwise_kernel", (1, lcj, lcb), "pw", cons, (1, g["cj"], cb))
                        add(rn + "/kernel@composed", (1, lcj, lcb), "conv", cons, (1, g["cj"], cb))
                    else:
                        add(rn + "/kernel", (1, lcj, lcb), "conv", cons, (1, g["cj"], cb))
        add("fc/kernel", (self.Hl, self.V), "fc_w", None, (self.H, self.V))
        add("fc/bias", (self.V,), "fc_b")
        off = 0
        hoff = 0
        for s in specs:
            n = 1
            for d in s["shape"]:
                n *= d
            s["size"] = n
            n = 1
            for d in s["store_shape"]:
                n *= d
            s["store_size"] = n
            s["phys"] = s["store_shape"]          # (K, C_in, C_out) the kernels see (the folded layer keeps K here)
            s["view"] = tuple(slice(0, d) for d in s["shape"])
            s["offset"] = off
            if s["kind"] == "conv":
                s["half_offset"] = hoff
            off += _align(n)
            if s["kind"] == "conv":
                hoff += _align(n)
        self.specs = specs
        self.by_name = {s["name"]: s for s in specs}
        self._fix_folded_layout()
        total = self._total
        dev = self.device
        self.master = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(total, dtype=torch.float32, device=dev)
        # one bf16 working copy in the natural TF layout [K][Cin][Cout]: it is dgrad's K-major B operand
        # (reduction over C_out) and forward's MN-major B operand (reduction over C_in) at the same time
        self.wb = torch.zeros(self._half_total, dtype=self.act_torch, device=dev)
        # per residual source: bf16 kernels of all consumers side by side [C_j, Ntot_j] (gathered from wb
        # at the start of every forward) and the fp32 weight gradient of the merged GEMM (scattered back)
        self.wcat = [torch.zeros(g["cj"], g["ntot"], dtype=self.act_torch, device=dev) for g in self.res_groups]
        self.dwcat = [torch.zeros(g["cj"], g["ntot"], dtype=torch.float32, device=dev) for g in self.res_groups]
        # BN moving statistics [2][C] per BN instance (moving_mean = 0, moving_variance = 1)
        self.moving = {}
        for s in specs:
            if s["kind"] == "gamma":
                C = s["store_shape"][0]
                mv = torch.zeros(2, C, dtype=torch.float32, device=dev)
                mv[1].fill_(1.0)
                self.moving[s["name"][:-len("/gamma")]] = mv
        self.init_parameters(self.seed)

    def _fix_folded_layout(self):
        """Stride-2 layers: storage holds 2*K' taps (zero lead / trail taps) so that the same memory is
        a stride-1 kernel [K', 2*C_in, C_out].  Recompute offsets with the padded sizes."""
        off = 0
        hoff = 0
        for s in self.specs:
            lyr = s["layer"]
            if s["kind"] == "conv" and lyr is not None and lyr.fold and s["name"] == lyr.wname:
                K, pl_info = lyr.K, None
                # pad_left depends on T parity; pad_to guarantees even T (asserted at run time)
                pl = max((lyr.K - 1) * lyr.dil + 1 - lyr.stride, 0) // 2  # SAME pad_left for even T_in
                j_min = math.floor(-pl / 2.0)
                j_max = math.floor((K - 1 - pl) / 2.0)
                lyr.kK = j_max - j_min + 1
                lyr.kpad = -j_min
                lyr.lead_taps = 2 * j_min + pl  # <= 0 -> number of leading zero taps = -lead
                lyr.lead_taps = -lyr.lead_taps
                lyr.kC_in = 2 * lyr.c_in
                lyr.orig_pad_left = pl
                s["store_size"] = 2 * lyr.kK * lyr.c_in * lyr.c_out
                s["store_shape"] = (2 * lyr.kK, lyr.c_in, lyr.c_out)
                s["view"] = (slice(lyr.lead_taps, lyr.lead_taps + K), slice(0, lyr.lc_in), slice(0, lyr.lc_out))
            elif lyr is not None and s["name"] in (lyr.wname, lyr.name + "/depthwise_kernel") and not lyr.fold:
                lyr.kK, lyr.kC_in = lyr.K, lyr.c_in
                _, lyr.kpad, _ = same_padding(1 << 20, lyr.K, 1, lyr.dil)
                lyr.lead_taps = 0
            s["offset"] = off
            off += _align(s["store_size"])
            if s["kind"] == "conv":
                s["half_offset"] = hoff
                hoff += _align(s["store_size"])
        self._total = off
        self._half_total = hoff

    def _valid_slice(self, s):
        """(start, size) of the trainable part inside the stored tensor."""
        lyr = s["layer"]
        if s["kind"] == "conv" and lyr is not None and lyr.fold and s["name"] == lyr.wname:
            per_tap = lyr.c_in * lyr.c_out
            return lyr.lead_taps * per_tap, lyr.K * per_tap
        return 0, s["store_size"]

    def param_view(self, name, buf=None):
        """fp32 view (logical shape) of a parameter inside the flat master (or grad / mom) buffer."""
        s = self.by_name[name]
        buf = self.master if buf is None else buf
        full = buf[s["offset"]:s["offset"] + s["store_size"]].view(*s["store_shape"])
        return full[s["view"]]

    def named_parameters(self):
        """The model's variables (the composed pseudo-kernels of sep_conv1d layers are derived, not variables)."""
        return [(s["name"], self.param_view(s["name"])) for s in self.specs if not s["name"].endswith("@composed")]

    def init_parameters(self, seed=0):
        """tf.contrib.layers.xavier_initializer, n = (fan_in + fan_out) / 2: "xavier_truncnorm"
        (uniform=False: truncated normal, std sqrt(1.3 / n)) or "xavier_uniform" (uniform=True, also what
        tf.layers uses when no initializer is given: limit sqrt(3 / n)).  The Jasper config initialises the
        encoder kernels truncated-normal and the decoder uniform.  BN gamma = 1 / beta = 0, bias = 0."""
        gen = torch.Generator().manual_seed(seed)

        def xavier(shape, fan_in, fan_out, kind):
            n = (fan_in + fan_out) / 2.0
            if kind == "xavier_uniform":
                lim = math.sqrt(3.0 / n)
                return (torch.rand(shape, generator=gen) * 2 - 1) * lim
            std = math.sqrt(1.3 / n)
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
            return t

        for s in self.specs:
            v = self.param_view(s["name"])
            if s["name"].endswith("@composed"):
                v.zero_()
            elif s["kind"] in ("conv", "dw", "pw"):
                K, ci, co = s["shape"]
                v.copy_(xavier(s["shape"], K * ci, K * co, self.encoder_init))
            elif s["kind"] == "gamma":
                v.fill_(1.0)
            elif s["kind"] in ("beta", "fc_b"):
                v.zero_()
            elif s["kind"] == "fc_w":
                v.copy_(xavier(s["shape"], s["shape"][0], s["shape"][1], self.decoder_init))
        self.sync_half_copies()

    def load_parameters(self, params):
        """params: dict name -> tensor/ndarray with the reference's variable names (relative)."""
        for name, val in params.items():
            if name not in self.by_name:
                raise KeyError("unknown parameter %r" % name)
            self.param_view(name).copy_(torch.as_tensor(val, dtype=torch.float32))
        self.sync_half_copies()

    def sync_half_copies(self):
        """fp32 master -> bf16 natural + transposed copies for every conv kernel."""
        st = L.stream_ptr()
        for s in self.specs:
            if s["kind"] != "conv":
                continue
            if s["name"].endswith("@composed"):
                self._compose_call(s["name"])()
                continue
            K, R, C = self._kernel_geom(s)
            L.check(self.lib.os2s_weight_cast_transpose_p(
                _vp(self.master.data_ptr() + 4 * s["offset"]), _vp(self.wb.data_ptr() + 2 * s["half_offset"]),
                _vp(0), K, R, C, self.dtypes, st), "weight_cast_transpose")

    def _compose_call(self, cname, stream_handle=None):
        """Bound launch of os2s_sepconv_compose for one composed pseudo-kernel: D[k,c] * P[c,o] -> its 16-bit slot."""
        base = cname[:-len("/kernel@composed")]
        s = self.by_name[cname]
        sd, sp = self.by_name[base + "/depthwise_kernel"], self.by_name[base + "/pointwise_kernel"]
        K, (_, ci, co) = s["shape"][0], s["phys"]
        st0, _ = self._valid_slice(s)
        args = [_vp(self.master.data_ptr() + 4 * sd["offset"]), _vp(self.master.data_ptr() + 4 * sp["offset"]),
                _vp(self.wb.data_ptr() + 2 * (s["half_offset"] + st0)), K, ci, co, self.dtypes]
        fn = self.lib.os2s_sepconv_compose

        def call():
            st = stream_handle if stream_handle is not None else L.stream_ptr()
            return L.check(fn(*(args + [st])), "os2s_sepconv_compose")
        call.fn, call.args = fn, args
        return call

    def layer_first_offset(self, li):
        """Offset (elements) of the first variable of layer li in the flat parameter / gradient buffers."""
        l = self.layers[li]
        return self.by_name[(l.name + "/depthwise_kernel") if l.sep else (l.name + "/kernel")]["offset"]

    def _decompose_args(self, cname):
        """Arguments of os2s_sepconv_decompose_grad (without the stream) for one composed pseudo-kernel."""
        base = cname[:-len("/kernel@composed")]
        s = self.by_name[cname]
        sd, sp = self.by_name[base + "/depthwise_kernel"], self.by_name[base + "/pointwise_kernel"]
        K, (_, ci, co) = s["shape"][0], s["phys"]
        st0, _ = self._valid_slice(s)
        g = self.grad.data_ptr()
        m = self.master.data_ptr()
        return [_vp(g + 4 * (s["offset"] + st0)), _vp(m + 4 * sd["offset"]), _vp(m + 4 * sp["offset"]),
                _vp(g + 4 * sd["offset"]), _vp(g + 4 * sp["offset"]), K, ci, co]

    def _kernel_geom(self, s):
        lyr = s["layer"]
        if lyr is not None and s["name"] == lyr.wname and not (lyr.sep and lyr.sep_mode == "split"):
            return lyr.kK, lyr.kC_in, lyr.c_out      # (the folded stride-2 layer: K' taps over 2 C_in channels)
        return s["phys"]                             # pointwise / residual kernels: K = 1

    # ----------------------------------------------------------------- optimizer
    def set_optimizer(self, algo="novograd", beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.0,
                      momentum=0.9, grad_averaging=False, ema_persist=False, larc_eta=0.0, larc_eps=1e-7,
                      larc_min_update=1e-7, larc_mode="clip", learning_rate=0.01, min_lr=0.0, power=1.0,
                      decay_steps=0, begin_decay_at=0, warmup_steps=0, loss_scaling=True, scale_min=1.0,
                      scale_max=2.0 ** 14, step_factor=2.0, step_window=2000, initial_scale=None,
                      lr_policy="poly_decay", decay_rate=1.0, use_staircase_decay=False, iter_size=1,
                      l2_regularizer_scale=0.0, max_grad_norm=0.0, freeze_variables_regex=None):
        """algo: "novograd" | "momentum" | "adam"; lr_policy: "poly_decay" | "cosine_decay" | "exp_decay" |
        "fixed_lr" (lr_policies.py); l2_regularizer_scale: tf.contrib.layers.l2_regularizer(scale) on the
        variables the reference builds with it (conv / dense kernels and BN gammas, conv_blocks.py:203,219;
        fc_decoders.py:138); iter_size > 1 accumulates g / iter_size over that many calls of
        train_step and updates on the last one (optimizers.py:212-259); max_grad_norm > 0: global-norm
        clipping (optimizers.py:408-433, exclusive with LARC :161-164); freeze_variables_regex: variables whose
        name matches (re.match on the reference's variable name, models/model.py:502-507) are not updated."""
        hp = OptHParams()
        if algo not in ("novograd", "momentum", "adam"):
            raise ValueError("JasperEngine: optimizer %r has no fused step (novograd / momentum / adam)" % (algo,))
        policies = {"poly_decay": 0, "cosine_decay": 1, "exp_decay": 2, "fixed_lr": 3}
        if lr_policy not in policies:
            raise ValueError("JasperEngine: lr_policy %r is not built (%s)" % (lr_policy, ", ".join(sorted(policies))))
        hp.algo = {"novograd": 0, "momentum": 1, "adam": 2}[algo]
        hp.lr_policy, hp.decay_rate, hp.staircase = policies[lr_policy], float(decay_rate), int(bool(use_staircase_decay))
        self.iter_size = int(iter_size)
        if self.iter_size < 1:
            raise ValueError("JasperEngine: iter_size must be >= 1")
        self._micro = 0
        hp.beta1, hp.beta2, hp.epsilon = beta1, beta2, epsilon
        hp.weight_decay, hp.momentum = weight_decay, momentum
        hp.grad_averaging, hp.ema_persist = int(grad_averaging), int(ema_persist)
        hp.larc_eta, hp.larc_eps, hp.larc_min_update = larc_eta, larc_eps, larc_min_update
        hp.larc_mode = 0 if larc_mode == "clip" else 1
        hp.lr0, hp.min_lr, hp.power = learning_rate, min_lr, power
        hp.decay_steps, hp.begin_decay_at, hp.warmup_steps = int(decay_steps), int(begin_decay_at), int(warmup_steps)
        hp.use_loss_scaler = int(bool(loss_scaling))
        hp.scale_min, hp.scale_max, hp.step_factor = scale_min, scale_max, step_factor
        hp.step_window = int(step_window)
        hp.world_size = self.world_size
        if max_grad_norm and larc_eta > 0:
            raise AttributeError("LARC and gradient norm clipping should not be used together")
        hp.max_grad_norm = float(max_grad_norm or 0.0)
        hp.wb_f16 = 1 if self.act_dtype == "fp16" else 0
        self.hp = hp
        dev = self.device
        n = len(self.specs)
        chunk = self.lib.os2s_opt_chunk_elems()
        ptr = lambda buf, s, esz: buf.data_ptr() + esz * s["offset"]
        w = [ptr(self.master, s, 4) for s in self.specs]
        g = [ptr(self.grad, s, 4) for s in self.specs]
        m = [ptr(self.mom, s, 4) for s in self.specs]
        wb = [(self.wb.data_ptr() + 2 * s["half_offset"]) if (s["kind"] == "conv" and not s["name"].endswith("@composed"))
              else 0 for s in self.specs]
        sizes = [s["store_size"] for s in self.specs]
        ct, co = [], []
        for i, sz in enumerate(sizes):
            for o in range(0, sz, chunk):
                ct.append(i)
                co.append(o)
        i64 = lambda x: torch.tensor(x, dtype=torch.int64, device=dev)
        # Adam: second moments, laid out like the momentum buffer; gradient accumulator for iter_size > 1
        self.mom2 = torch.zeros(self._total, dtype=torch.float32, device=dev) if algo == "adam" else None
        self.grad_acc = torch.zeros(self._total, dtype=torch.float32, device=dev) if self.iter_size > 1 else None
        v = [ptr(self.mom2, s, 4) for s in self.specs] if algo == "adam" else [0] * n
        # the reference builds the 1x1 residual kernels WITHOUT a regularizer (conv_blocks.py:80-86)
        def regularized(s):
            n = s["name"]
            if n.endswith("@composed") or s["kind"] not in ("conv", "dw", "pw", "gamma", "fc_w"):
                return False
            # residual convs are built without a regularizer (their BN gammas have one)
            return not (s["kind"] in ("conv", "dw", "pw") and ("/res_" in n or "/res/" in n))
        reg = [float(l2_regularizer_scale) if regularized(s) else 0.0 for s in self.specs]
        self._reg = torch.tensor(reg, dtype=torch.float32, device=dev) if l2_regularizer_scale else None
        self._frozen = None
        self.frozen_names = []
        flags = [1 if s["name"].endswith("@composed") else 0 for s in self.specs]   # derived, never updated
        if freeze_variables_regex:
            import re
            pat = re.compile(freeze_variables_regex)
            for i, s in enumerate(self.specs):
                if not s["name"].endswith("@composed") and pat.match(self.var_scope_name(s["name"])):
                    flags[i] = 1
                    self.frozen_names.append(s["name"])
        if any(flags):
            self._frozen = torch.tensor(flags, dtype=torch.int32, device=dev)
        self._opt = {
            "w": i64(w), "g": i64(g), "m": i64(m), "v": i64(v), "wb": i64(wb), "sizes": i64(sizes),
            "ct": torch.tensor(ct, dtype=torch.int32, device=dev), "co": i64(co),
            "norms": torch.zeros(2 * n, dtype=torch.float32, device=dev),
            "nonfinite": torch.zeros(1, dtype=torch.int32, device=dev),
            "coef": torch.zeros(n, dtype=torch.float32, device=dev),
            "ema": torch.zeros(n, dtype=torch.float32, device=dev),
            "n": n, "n_chunks": len(ct),
        }
        scale0 = (initial_scale if initial_scale is not None else (scale_max if loss_scaling else 1.0))
        self.fstate = torch.zeros(8, dtype=torch.float32, device=dev)
        self.fstate[0] = scale0
        self.istate = torch.zeros(8, dtype=torch.int64, device=dev)
        self.istate[1] = -1
        self._ws = collections.OrderedDict()

    # ----------------------------------------------------------------- workspace
    def _workspace(self, B, T):
        """Per-shape workspaces (activations, launch plan, captured graph) are cached least-recently-used
        under a byte budget (OS2S_WS_BUDGET_GB, default 60% of the device memory): with real variable-length
        data T takes ~100 values, and a B = 32 x 15 s workspace alone is 4.5 GB."""
        key = (B, T, self.training)
        ws = self._ws.pop(key, None)
        if ws is None:
            budget = self._ws_budget_bytes()
            need = _Workspace.estimate_bytes(self, B, T)
            while self._ws and sum(w.nbytes for w in self._ws.values()) + need > budget:
                old_key = next(iter(self._ws))
                old = self._ws.pop(old_key)
                if getattr(self, "_last_ws", None) is old:
                    self._last_ws = None
                old.release()
            ws = _Workspace(self, B, T)
        self._ws[key] = ws          # (re)insert as most recently used
        return ws

    def clear_workspaces(self):
        """Forget every cached per-shape workspace (after changing dropout / topology options by hand)."""
        for w in self._ws.values():
            w.release()
        self._ws = collections.OrderedDict()
        self._last_ws = None

    def _ws_budget_bytes(self):
        b = os.environ.get("OS2S_WS_BUDGET_GB")
        if b:
            return int(float(b) * (1 << 30))
        if self.device.type == "cuda":
            return int(0.6 * torch.cuda.get_device_properties(self.device).total_memory)
        return 1 << 62

    # ---------------------------------------------------------------- public API
    def forward_encoder(self, feats, feat_lens):
        """TDNNEncoder._encode: feats bf16 [B,T,F] (zero padded), lens int32 [B] -> (bf16 [B,T',H], lens)."""
        B, T, F = feats.shape
        self._check_feats(feats)
        ws = self._workspace(B, T)
        ws.run_forward(feats, feat_lens)
        return ws.A[-1], ws.lens_out

    def forward_decoder(self):
        """FullyConnectedTimeDecoder._decode on the last encoder output -> logits fp32 [B,T',V]."""
        ws = self._last_ws
        ws.run_decoder()
        return ws.logits

    def forward(self, feats, feat_lens):
        """encoder + decoder: (logits fp32 [B,T',V] batch-major, out_lens int32 [B])."""
        _, out_lens = self.forward_encoder(feats, feat_lens)
        return self.forward_decoder(), out_lens

    def set_comm(self, comm, bucket_bytes=None):
        """Data-parallel gradient exchange (reference: hvd.allreduce per gradient,
        optimizers/optimizers.py:77-104).  Gradients live in one flat fp32 buffer whose tail is
        produced first by the backward pass; contiguous buckets are all-reduced (SUM) on a side stream
        as soon as the layers that write them have been enqueued, overlapping the exchange with the remaining
        backward kernels; the optimizer waits for the side stream."""
        self.comm = comm if (comm is not None and comm.size() > 1) else None
        if bucket_bytes:
            self.bucket_bytes = int(bucket_bytes)
        if self.comm is not None and self._side is None:
            self._side = torch.cuda.Stream()
        # one node, NCCL backend: the buckets are summed over NVLink peer memory by the copy engines
        # (csrc/peer.cu) instead of NCCL's SM-resident ring; None = NCCL all-reduce
        self.peer = None
        if self.comm is not None and hasattr(self.comm, "make_peer_exchange"):
            self.peer = self.comm.make_peer_exchange(self.grad, self.grad_buckets())
            if os.environ.get("OS2S_PEER_GRAPH", "") == "":
                self.graph_with_peer = self.comm.size() == 2
        self._ws = collections.OrderedDict()

    def grad_buckets(self):
        """[(start, end)] (elements of the flat gradient buffer) in the order backward completes them: cut at
        layer boundaries, last layer first, as soon as a bucket holds bucket_bytes."""
        out = []
        end = self._total
        tail_cut = False
        for li in range(len(self.layers) - 1, -1, -1):
            start = self.layer_first_offset(li)
            # the exchange of the LAST bucket is not hidden by any backward work: keep it short
            taper = not tail_cut and 0 < start * 4 <= self.tail_bucket_bytes and end > start
            if (end - start) * 4 >= self.bucket_bytes or li == 0 or taper:
                out.append((start, end))
                end = start
                tail_cut = tail_cut or taper
        return out

    def set_training(self, flag):
        """train mode: batch statistics + dropout; eval mode: moving statistics, no dropout
        (tdnn_encoder.py:127-128, tf.layers.batch_normalization(training=...))."""
        self.training = bool(flag)

    def loss_only(self, labels, label_lens):
        """Per-utterance CTC loss of the last forward (no backward pass) -- eval mode."""
        ws = self._last_ws
        return ws.run_loss_only(labels, label_lens)

    def backward_from_dlogits(self, dlogits):
        """Backward pass from a given gradient wrt the logits (fp32 [B,T',V]); used by tests and by
        losses other than CTC.  Gradients land in self.grad (scaled by whatever scale dlogits carries)."""
        ws = self._last_ws
        ws.run_backward(None, None, dlogits=dlogits)

    def loss_and_backward(self, labels, label_lens):
        """labels int32 [B,Lmax]; returns per-utterance loss (device tensor) after enqueueing backward."""
        ws = self._last_ws
        ws.run_backward(labels, label_lens)
        return ws.loss

    def optimizer_step(self):
        """LARC + loss-scaler + NovoGrad on the (already summed over ranks) gradients."""
        self._launch_optimizer()

    def _launch_optimizer(self):
        o = self._opt
        st = L.stream_ptr()
        if self._profile is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        L.check(self.lib.os2s_opt_step3(
            L.ptr(o["w"]), L.ptr(o["g"]), L.ptr(o["m"]), L.ptr(o["v"]) if self.mom2 is not None else _vp(0),
            L.ptr(o["wb"]), L.ptr(self._reg) if self._reg is not None else _vp(0),
            L.ptr(self._frozen) if self._frozen is not None else _vp(0), L.ptr(o["sizes"]), L.ptr(o["ct"]),
            L.ptr(o["co"]), o["n"], o["n_chunks"], ctypes.byref(self.hp), L.ptr(o["norms"]),
            L.ptr(o["nonfinite"]), L.ptr(self.fstate), L.ptr(self.istate), L.ptr(o["coef"]), L.ptr(o["ema"]),
            st), "os2s_opt_step")
        if self._profile is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            # per parameter: g, w read twice (norms + update), m read + written, w written, 2-byte copy of conv kernels
            self._profile.append(("hbm:optimizer", float(self._total) * 28.0 + float(self._half_total) * 2.0, e0, e1))
        self.step_count += 1

    def train_step(self, feats, feat_lens, labels, label_lens):
        """forward, loss + backward (with the overlapped gradient all-reduce when set_comm was called),
        optimizer -- one call, replayed as a CUDA graph in steady state.  Returns the per-utterance
        loss tensor (device, static buffer)."""
        B, T, F = feats.shape
        self._check_feats(feats)
        return self._workspace(B, T).run_train_step(feats, feat_lens, labels, label_lens)

    def _check_feats(self, feats):
        """Features arrive in the activation storage format; a bf16 / fp16 / fp32 tensor of the other kind is
        converted here (one small elementwise kernel), anything else is an error."""
        if feats.dim() != 3 or feats.shape[2] != self.F or not feats.is_contiguous():  # F: the reference's feature count
            raise ValueError("JasperEngine: features must be contiguous [B,T,%d]" % self.F)
        if feats.dtype not in (torch.bfloat16, torch.float16, torch.float32):
            raise ValueError("JasperEngine: features must be bf16 / fp16 / fp32")

    def greedy_decode(self):
        """tf.nn.ctc_greedy_decoder on the last forward's logits -> (tokens [B,T'], lens [B])."""
        ws = self._last_ws
        st = L.stream_ptr()
        L.check(self.lib.os2s_ctc_greedy(L.ptr(ws.logits), L.ptr(ws.lens_out), L.ptr(ws.tokens), L.ptr(ws.tok_lens),
                                         L.ptr(ws.neg_sum), ws.B, ws.T2, self.V, _c_ll(ws.T2 * self.V),
                                         _c_ll(self.V), 1, st), "os2s_ctc_greedy")
        return ws.tokens, ws.tok_lens

    def kernel_launches_per_step(self):
        """Kernels of libos2s_b200 launched by one training step (conv + BN stats/apply + FC + CTC(3) +
        BN bwd (2 each) + optimizer (3) + transpose (1) + featurizer (3)); memsets are not counted."""
        ws = self._last_ws
        n = 0
        for entry in ws._fwd_plan:
            n += 1
        n += 1  # fc fwd
        for entry in (ws._bwd_plan or []):
            name = entry[0].__name__
            n += {"os2s_ctc_loss_fwd_bwd": 3, "os2s_fc_bwd_p": 2, "os2s_bn_bwd_p": 2, "os2s_bn_bwd_apply_p": 1, "zero_slices": 0,
                  "os2s_sepconv_decompose_grad": 2,
                  "bucket_allreduce": 4 if self.peer is not None else 0,   # signal, wait, slice sum, signal
                  "stream_record": 0, "stream_wait": 0}.get(name, 1)
        return n + 3 + 3 + (1 if self.peer is not None else 0)              # + os2s_peer_finish's wait

    def profile_conv_launches(self, step_fn, steps=2):
        """Time every tensor-core conv launch of `steps` instrumented steps with CUDA events.
        Returns achieved TFLOP/s over all conv launches (algorithmic FLOPs / summed kernel time)."""
        self._profile = []
        try:
            for _ in range(steps):
                step_fn()
            torch.cuda.synchronize()
            rec = self._profile
        finally:
            self._profile = None
        by, hb = {}, {}
        for kind, work, e0, e1 in rec:
            tgt = hb if kind.startswith("hbm:") else by
            d = tgt.setdefault(kind, [0.0, 0.0, 0])
            d[0] += work
            d[1] += e0.elapsed_time(e1)
            d[2] += 1
        tot_f = sum(d[0] for d in by.values())
        tot_ms = sum(d[1] for d in by.values())
        tot_b = sum(d[0] for d in hb.values())
        tot_bms = sum(d[1] for d in hb.values())
        return {"tflops": round(tot_f / tot_ms / 1e9, 1) if tot_ms > 0 else 0.0,
                "ms": round(tot_ms / steps, 3), "tflop": round(tot_f / steps / 1e12, 3),
                "launches": int(sum(d[2] for d in by.values()) / steps),
                "by_kind": {k: {"tflops": round(d[0] / d[1] / 1e9, 1), "ms_per_step": round(d[1] / steps, 3),
                                "launches_per_step": d[2] // steps} for k, d in by.items()},
                # HBM-bound family (BN forward / backward, optimizer): algorithmic bytes / CUDA-event time
                "hbm": {"gbs": round(tot_b / tot_bms / 1e6, 1) if tot_bms > 0 else 0.0,
                        "ms": round(tot_bms / steps, 3), "gbytes": round(tot_b / steps / 1e9, 3),
                        "by_kind": {k[4:]: {"gbs": round(d[0] / d[1] / 1e6, 1), "ms_per_step": round(d[1] / steps, 3),
                                            "launches_per_step": d[2] // steps} for k, d in hb.items()}}}


class _ZeroSlices(object):
    """Plan entry: zero a few gradient slices (structural zeros of the folded first-layer kernel), on
    the stream that produced them (the weight-gradient stream)."""
    __name__ = "zero_slices"

    def __init__(self, ws, slices):
        self.ws, self.slices = ws, slices

    def __call__(self):
        with torch.cuda.stream(self.ws.aux_stream()):
            for z in self.slices:
                z.zero_()
        return 0


class _ZeroMain(object):
    """Plan entry: zero a scratch tensor on the main stream."""
    __name__ = "zero_slices"

    def __init__(self, t):
        self.t = t

    def __call__(self):
        self.t.zero_()
        return 0


class _StreamRecord(object):
    """Plan entry: record an event on the main or the aux stream of the workspace."""
    __name__ = "stream_record"

    def __init__(self, ws, which, event):
        self.ws, self.which, self.event = ws, which, event

    def __call__(self):
        self.event.record(self.ws.aux_stream() if self.which == "aux" else torch.cuda.current_stream())
        return 0


class _StreamWait(object):
    """Plan entry: make the main / aux stream wait for an event (no-op when both are the same stream)."""
    __name__ = "stream_wait"

    def __init__(self, ws, which, event):
        self.ws, self.which, self.event = ws, which, event

    def __call__(self):
        (self.ws.aux_stream() if self.which == "aux" else torch.cuda.current_stream()).wait_event(self.event)
        return 0


class _BucketAllReduce(object):
    """Plan entry: all-reduce (SUM) one contiguous gradient bucket on the side stream once everything
    enqueued so far on the compute stream has finished."""
    __name__ = "bucket_allreduce"

    def __init__(self, eng, index, flat_slice):
        self.eng, self.index, self.slice = eng, index, flat_slice
        self.event = torch.cuda.Event()

    def __call__(self):
        eng = self.eng
        if eng._suppress_comm:
            return 0
        self.event.record()
        if eng.peer is not None:
            eng._side.wait_event(self.event)
            eng.peer.exchange_bucket(self.index, eng._side)
            return 0
        with torch.cuda.stream(eng._side):
            eng._side.wait_event(self.event)
            eng.comm.allreduce_(self.slice)
        return 0


class _Workspace(object):
    """Per-(B,T) activations, gradients and the pre-bound launch plan."""

    @staticmethod
    def estimate_bytes(eng, B, T):
        T2 = T // 2 if eng.layers[0].fold else T
        csz = 4 if eng.conv_dtype == "fp32" else 2
        rows = B * T2
        n = sum(l.c_out for l in eng.layers) * (csz + 2)
        n += sum(g["ntot"] for g in eng.res_groups) * (csz + 2)
        n += sum(b[0] for b in eng.block_inputs) * 4
        n += 3 * max(l.c_out for l in eng.layers) * 2
        return rows * n + B * T * eng.Fp * 2 + (64 << 20)

    def release(self):
        """Drop the captured graph and every tensor so the caching allocator can reuse the memory."""
        self.graph = None
        self._fwd_plan = self._bwd_plan = None
        self._keep = []
        for k in ("Y", "A", "YRcat", "dYRcat", "dA", "dY2", "dY", "dres", "red", "stats_all", "stats", "stats_cat",
                  "mean_invstd", "red_all", "fused_red", "logits", "dlogits", "feats", "_ctc_ws", "tokens"):
            if hasattr(self, k):
                setattr(self, k, None)

    def __init__(self, eng, B, T):
        self.eng = eng
        self.B, self.T = B, T
        lib = eng.lib
        dev = eng.device
        # ONE mutable stream handle shared by every pre-bound call: updating its .value retargets the
        # whole plan (needed for CUDA-graph capture, which runs on torch's capture stream)
        st = _vp(torch.cuda.current_stream().cuda_stream)
        self._st = st
        first = eng.layers[0]
        if first.fold:
            if T % 2 != 0:
                raise ValueError("JasperEngine: input length must be even for the stride-2 layer (pad_to)")
            T2 = T // 2
        else:
            T2 = T
        self.T2 = T2
        M = B * T2
        self.M = M
        bf = lambda *shape: torch.empty(*shape, dtype=eng.act_torch, device=dev)   # gradients: the 16-bit format
        f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        layers = eng.layers
        nl = len(layers)
        # conv outputs (BN inputs) are fp16 (never a tensor-core operand, 3 more mantissa bits than bf16) or
        # fp32 (conv_dtype); layer outputs are bf16 or fp16 (act_dtype)
        # (zeros: a skipped forward tile leaves its rows untouched, and the BN-backward sums multiply those rows by
        # exact zeros -- they must never hold the NaN bit patterns of uninitialised memory)
        f16 = lambda *shape: torch.zeros(*shape, dtype=eng.conv_torch, device=dev)
        act = lambda *shape: torch.empty(*shape, dtype=eng.act_torch, device=dev)
        self.Y = [f16(B, T2, l.c_out) for l in layers]
        self.A = [act(B, T2, l.c_out) for l in layers]
        # residual-branch conv outputs / their gradients, one matrix per SOURCE: [B, T2, sum_b C_b]
        # (consumer b's branch is the column slice starting at its first column)
        self.YRcat = [f16(B, T2, g["ntot"]) for g in eng.res_groups]
        self.dYRcat = [bf(B, T2, g["ntot"]) for g in eng.res_groups]
        cmax = max(l.c_out for l in layers)
        self.dA = bf(B, T2, cmax)
        nres_max = max([len(l.res_sources) for l in layers] + [0])
        # conv-output gradients are ping-ponged by layer parity so that wgrad(l) (aux stream) can still
        # read dY of layer l while bn_bwd(l-1) already writes the other buffer
        self.dY2 = [bf(B, T2, cmax), bf(B, T2, cmax)]
        self.dY = self.dY2[0]
        # split sep_conv1d layers: depthwise outputs (saved for the pointwise weight gradient) and their gradients
        self.Z = {li: act(B, T2, l.c_in) for li, l in enumerate(layers) if l.sep and l.sep_mode == "split"}
        cin_max = max([l.c_in for l in layers if l.sep and l.sep_mode == "split"] + [0])
        self.dZ2 = [bf(B, T2, cin_max), bf(B, T2, cin_max)] if cin_max else None
        self._st_aux = _vp(0)
        self.dres = [f32(B, T2, b[0]) for b in eng.block_inputs]
        self.red = f32((2 + nres_max) * cmax)
        self.lens_in = torch.zeros(B, dtype=torch.int32, device=dev)
        self.lens_out = torch.zeros(B, dtype=torch.int32, device=dev)
        # BN bookkeeping: stats arena (zeroed every step) and saved mean/invstd
        n_bn = sum(1 + len(l.res_sources) for l in layers)
        # one flat arena (a single memset per step): [2][cmax] per main-path BN, then [2][Ntot_j] per source
        n_main = len(layers)
        cat_sizes = [2 * g["ntot"] for g in eng.res_groups]
        self.stats_all = torch.zeros(n_main * 2 * cmax + sum(cat_sizes), dtype=torch.float32, device=dev)
        self.stats = self.stats_all[:n_main * 2 * cmax].view(n_main, 2, cmax)
        self.stats_cat, o = [], n_main * 2 * cmax
        for sz in cat_sizes:
            self.stats_cat.append(self.stats_all[o:o + sz])
            o += sz
        self.mean_invstd = torch.zeros(n_bn, 2, cmax, dtype=torch.float32, device=dev)
        # [2][C] BN-backward sums of every plain layer whose dA comes from a conv data-gradient kernel
        self.fused_red = {}
        offs, o = {}, 0
        for li, l in enumerate(layers):
            plain = (not l.res_sources) and (li not in eng.src_of_layer_output) and li + 1 < len(layers)
            # (the fused sums ride in the tcgen05 data-gradient kernel of the NEXT layer: not when that is a
            # depthwise + pointwise pair, whose dA comes out of the CUDA-core depthwise kernel)
            plain = plain and not (li + 1 < len(layers) and layers[li + 1].sep and layers[li + 1].sep_mode == "split")
            if eng.fuse_bn_reduce and plain and l.c_out >= eng.fuse_bn_min_channels and l.c_out % 64 == 0:
                offs[li] = o
                o += 2 * l.c_out
        self.red_all = torch.zeros(max(o, 1), dtype=torch.float32, device=dev)
        for li, off in offs.items():
            self.fused_red[li] = self.red_all[off:off + 2 * layers[li].c_out]
        self.logits = f32(B, T2, eng.V)
        self.dlogits = f32(B, T2, eng.V)
        self.loss = f32(B)
        self.tokens = torch.zeros(B, T2, dtype=torch.int32, device=dev)
        self.tok_lens = torch.zeros(B, dtype=torch.int32, device=dev)
        self.neg_sum = f32(B)
        # static input buffer, Fp >= F channels wide (pad channels stay zero)
        self.feats = torch.zeros(B, T, eng.Fp, dtype=eng.act_torch, device=dev)
        self.graph = None
        self.graph_L = -1
        self._eager_steps = 0
        self._ctc_ws = None
        self._ctc_L = -1
        self._build_forward_plan(st)
        self._bwd_plan = None
        self._bwd_L = -1
        self.nbytes = _Workspace.estimate_bytes(eng, B, T)

    # -- helpers
    def _p(self, t, off_elems=0, esz=None):
        esz = t.element_size() if esz is None else esz
        return _vp(t.data_ptr() + off_elems * esz)

    def _param_ptr(self, buf, name, esz=4):
        s = self.eng.by_name[name]
        return _vp(buf.data_ptr() + esz * s["offset"])

    def _half_ptr(self, buf, name):
        s = self.eng.by_name[name]
        return _vp(buf.data_ptr() + 2 * s["half_offset"])

    def _build_forward_plan(self, st):
        eng, lib = self.eng, self.eng.lib
        B, T2, M = self.B, self.T2, self.M
        plan = []
        self._seed_slots = []
        self.bn_slot = {}
        self.x_of_layer = []
        self._keep = []
        vparr = lambda ptrs: (ctypes.c_void_p * len(ptrs))(*[p.value if isinstance(p, _vp) else p for p in ptrs])
        iarr = lambda xs: (ctypes.c_int * len(xs))(*xs)
        llarr = lambda xs: (ctypes.c_longlong * len(xs))(*xs)
        # length-aware tile skipping (include/os2s.h, os2s_conv1d_fwd_p): only with the conv mask, which is what
        # makes every layer input zero past the utterance's length (OS2S_SKIP_TILES=0 computes every tile)
        self.skip_tiles = eng.use_conv_mask and os.environ.get("OS2S_SKIP_TILES", "1") != "0"
        rl = self._p(self.lens_out) if self.skip_tiles else _vp(0)
        nl_ = len(eng.layers)
        # (0) sep_conv1d: form the composed dense kernels D[k,c] * P[c,o] from the fp32 masters
        for sp_ in eng.specs:
            if sp_["name"].endswith("@composed"):
                cc = eng._compose_call(sp_["name"])
                plan.append([cc.fn, cc.args + [st]])
        # (1) lay the 1x1 residual kernels of every source side by side (bf16, from the working copy)
        src, dst, rows, rbytes, sp, dp = [], [], [], [], [], []
        for j, g in enumerate(eng.res_groups):
            for (lc, n, cb, col) in g["consumers"]:
                rn = eng.res_wname(eng.layers[lc], n)
                src.append(self._half_ptr(eng.wb, rn))
                dst.append(self._p(eng.wcat[j], col))
                rows.append(g["cj"])
                rbytes.append(cb * 2)
                sp.append(cb * 2)
                dp.append(g["ntot"] * 2)
        for i in range(0, len(src), 64):
            sl = slice(i, i + 64)
            a = (vparr(src[sl]), vparr(dst[sl]), iarr(rows[sl]), iarr(rbytes[sl]), llarr(sp[sl]), llarr(dp[sl]))
            self._keep.append(a)
            plan.append([lib.os2s_multi_copy_2d, [len(src[sl])] + list(a) + [st]])
        bn_idx = 0
        for li, l in enumerate(eng.layers):
            x_ptr = self._p(self.feats) if li == 0 else self._p(self.A[li - 1])
            self.x_of_layer.append(self.A[li - 1] if li > 0 else None)
            # (2) the input of this layer is a residual source: ONE GEMM for all its consumers' branches
            for j, g in enumerate(eng.res_groups):
                if g["first_layer"] == li:
                    stats_ptr = self._p(self.stats_cat[j]) if eng.training else _vp(0)
                    plan.append([lib.os2s_conv1d_fwd_p, [x_ptr, self._p(eng.wcat[j]), self._p(self.YRcat[j]), B, T2,
                                                         g["cj"], g["ntot"], 1, 1, 0, eng.conv_out_mode, stats_ptr,
                                                         rl, eng.dtypes, st],
                                 ("fwd", 2.0 * B * T2 * g["cj"] * g["ntot"])])
            flops = 2.0 * B * T2 * l.K * l.c_in * l.c_out  # algorithmic (un-folded) FLOPs
            slot_main = bn_idx
            bn_idx += 1
            # training: the conv epilogue accumulates the BN statistics of its (rounded) output
            stats_ptr = self._p(self.stats[li]) if eng.training else _vp(0)
            if l.sep and l.sep_mode == "split":
                # depthwise K-tap stage on the CUDA cores, then the pointwise stage as a 1x1 tcgen05 GEMM
                plan.append([lib.os2s_depthwise_conv1d,
                             [x_ptr, self._param_ptr(eng.master, l.name + "/depthwise_kernel"), self._p(self.Z[li]), B, T2,
                              l.c_in, l.K, -l.kpad, l.dil, eng.grad_out_mode, eng.dtypes, st],
                             ("hbm:depthwise", float(M) * l.c_in * 4)])
                call = [lib.os2s_conv1d_fwd_p, [self._p(self.Z[li]), self._half_ptr(eng.wb, l.wname), self._p(self.Y[li]),
                                                B, T2, l.c_in, l.c_out, 1, 1, 0, eng.conv_out_mode, stats_ptr,
                                                _vp(0), eng.dtypes, st],
                        ("fwd", 2.0 * B * T2 * l.c_in * l.c_out)]
            else:
                call = [lib.os2s_conv1d_fwd_p, [x_ptr, self._half_ptr(eng.wb, l.wname), self._p(self.Y[li]),
                                                B, T2, l.kC_in, l.c_out, l.kK, l.dil, l.kpad, eng.conv_out_mode, stats_ptr,
                                                # the last layer's output is not masked: every row of it is defined
                                                rl if li < nl_ - 1 else _vp(0), eng.dtypes, st],
                        ("fwd", flops)]
            plan.append(call)
            ys, lds = [self._p(self.Y[li])], [l.c_out]
            sts, st_lds = [self._p(self.stats[li])], [l.c_out]
            names = [l.name + "/bn"]
            slots = [slot_main]
            for n, j in enumerate(l.res_sources):
                jj, col = eng.res_col[(li, n)]
                ntot = eng.res_groups[jj]["ntot"]
                ys.append(self._p(self.YRcat[jj], col))
                lds.append(ntot)
                sts.append(self._p(self.stats_cat[jj], col))
                st_lds.append(ntot)
                names.append(eng.res_bn_name(l, n))
                slots.append(bn_idx)
                bn_idx += 1
            nb = len(ys)
            y_h, st_h = vparr(ys), vparr(sts)
            ld_h, stld_h = iarr(lds), iarr(st_lds)
            g_h = vparr([self._param_ptr(eng.master, nm + "/gamma") for nm in names])
            b_h = vparr([self._param_ptr(eng.master, nm + "/beta") for nm in names])
            mi_h = vparr([self._p(self.mean_invstd[s]) for s in slots])
            mv_h = vparr([self._p(eng.moving[nm]) for nm in names])
            self.bn_slot[li] = (slots, names, (y_h, ld_h, st_h, g_h, b_h, mi_h, mv_h))
            last = li == len(eng.layers) - 1
            lens_ptr = self._p(self.lens_out) if (eng.use_conv_mask and not last) else _vp(0)
            args = [nb, y_h, ld_h, st_h, stld_h, g_h, b_h, mi_h, mv_h, self._p(self.A[li]), lens_ptr, B, T2, l.c_out,
                    _c_float(eng.bn_eps), _c_float(eng.bn_momentum), _c_float(l.keep if eng.training else 1.0),
                    _c_u64((eng.seed * 1000003 + 4099 * li + 17) & 0xFFFFFFFFFFFFFFFF), 1, _c_float(eng.relu_clip),
                    0 if eng.training else 1, _vp(eng.istate.data_ptr() + 5 * 8), eng.dtypes, st]
            self._keep.append(stld_h)
            ysz = 4 if eng.conv_dtype == "fp32" else 2
            plan.append([lib.os2s_bn_apply_fwd_p, args, ("hbm:bn_fwd", float(M) * l.c_out * (nb * ysz + 2))])
        self._fc_call = [lib.os2s_fc_fwd_p, [self._p(self.A[-1]), self._param_ptr(eng.master, "fc/kernel"),
                                             self._param_ptr(eng.master, "fc/bias"), self._p(self.logits), M, eng.H,
                                             eng.V, eng.dtypes, st]]
        self._fwd_plan = plan
        self.n_launch_fwd = len(plan) + 2

    def set_inputs(self, feats, feat_lens):
        """Copy one batch into the static input buffers (the plans / graphs read only these)."""
        if feats.data_ptr() != self.feats.data_ptr():
            if self.eng.Fp == self.eng.F:
                self.feats.copy_(feats, non_blocking=True)
            else:
                self.feats[:, :, :self.eng.F].copy_(feats, non_blocking=True)
        self.lens_in.copy_(feat_lens.to(torch.int32), non_blocking=True)

    def set_targets(self, labels, label_lens):
        L_max = int(labels.shape[1])
        L_cap = max(32, -(-L_max // 32) * 32)  # bucket label capacity so plans / graphs are reused
        if self._bwd_plan is None or L_cap > self._bwd_L:
            # the capacity only grows (label_lens carries the real lengths): batches whose longest transcript
            # differs do not rebuild the plan / re-capture the graph
            self._build_backward_plan(L_cap)
            self.graph = None
            self._eager_steps = 0
        elif L_max < self._bwd_L:
            self.labels[:, L_max:].zero_()
        self.labels[:, :L_max].copy_(labels.to(torch.int32), non_blocking=True)
        self.label_lens.copy_(label_lens.to(torch.int32), non_blocking=True)

    def _body_forward(self):
        eng = self.eng
        self._st.value = torch.cuda.current_stream().cuda_stream
        s = eng.layers[0].stride
        torch.div(self.lens_in + (s - 1), s, rounding_mode="floor", out=self.lens_out)
        if eng.training:
            self.stats_all.zero_()
        self._exec(self._fwd_plan)

    def aux_stream(self):
        """Stream of the weight-gradient kernels: a dedicated one, or the current stream when the
        overlap is disabled or conv launches are being timed with events."""
        eng = self.eng
        if eng.overlap_wgrad and eng._profile is None:
            if eng._aux is None:
                eng._aux = torch.cuda.Stream()
            return eng._aux
        return torch.cuda.current_stream()

    def _body_backward(self, plan):
        eng = self.eng
        self._st.value = torch.cuda.current_stream().cuda_stream
        self._st_aux.value = self.aux_stream().cuda_stream
        self._exec(plan)
        if eng.comm is not None:
            if eng.peer is not None and not eng._suppress_comm:
                eng.peer.finish(eng._side)
            torch.cuda.current_stream().wait_stream(eng._side)

    def run_forward(self, feats, feat_lens):
        self.set_inputs(feats, feat_lens)
        self._body_forward()
        self.eng._last_ws = self

    def _exec(self, plan):
        """Run a launch plan; when the engine is in profiling mode, conv launches are bracketed by
        CUDA events on the launching stream (bench.py's roofline measurement)."""
        prof = self.eng._profile
        lib_check = L.check
        if prof is None:
            for entry in plan:
                rc = entry[0](*entry[1])
                if rc != 0:
                    lib_check(rc, entry[0].__name__)
            return
        for entry in plan:
            meta = entry[2] if len(entry) > 2 else None
            if meta is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = entry[0](*entry[1])
            if rc != 0:
                lib_check(rc, entry[0].__name__)
            if meta is not None:
                e1.record()
                prof.append((meta[0], meta[1], e0, e1))

    def run_decoder(self):
        self._st.value = torch.cuda.current_stream().cuda_stream
        fn, args = self._fc_call[0], self._fc_call[1]
        L.check(fn(*args), "os2s_fc_fwd_p")

    def _build_backward_plan(self, L_max):
        eng, lib = self.eng, self.eng.lib
        B, T2, M, st = self.B, self.T2, self.M, self._st
        dev = eng.device
        need = lib.os2s_ctc_workspace_bytes(B, T2, L_max)
        if self._ctc_ws is None or self._ctc_ws.numel() < need:
            self._ctc_ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
        self.labels = torch.zeros(B, L_max, dtype=torch.int32, device=dev)
        self.label_lens = torch.zeros(B, dtype=torch.int32, device=dev)
        plan = []
        plan.append([lib.os2s_ctc_loss_fwd_bwd,
                     [self._p(self.logits), self._p(self.labels), self._p(self.label_lens), self._p(self.lens_out),
                      self._p(self.dlogits), self._p(self.loss), self._p(self._ctc_ws), _c_size_t(int(need)),
                      self._p(eng.fstate), B, T2, eng.V, L_max, _c_ll(T2 * eng.V), _c_ll(eng.V), st]])
        plan.append([lib.os2s_fc_bwd_p, [self._p(self.A[-1]), self._p(self.dlogits),
                                         self._param_ptr(eng.master, "fc/kernel"), self._p(self.dA),
                                         self._param_ptr(eng.grad, "fc/kernel"), self._param_ptr(eng.grad, "fc/bias"),
                                         M, eng.H, eng.V, eng.dtypes, st]])
        nl = len(eng.layers)
        rl = self._p(self.lens_out) if self.skip_tiles else _vp(0)
        if self.fused_red:
            plan.append([_ZeroMain(self.red_all), []])
        buckets = eng.grad_buckets() if eng.comm is not None else []   # final once their first layer is enqueued
        n_bucket = 0
        sa = self._st_aux                # aux-stream handle (== main stream when overlap is off / profiling)
        ev_bn = [torch.cuda.Event() for _ in range(nl)]
        ev_wg = [torch.cuda.Event() for _ in range(nl)]
        vparr = lambda ptrs: (ctypes.c_void_p * len(ptrs))(*[p.value if isinstance(p, _vp) else p for p in ptrs])
        iarr = lambda xs: (ctypes.c_int * len(xs))(*xs)
        llarr = lambda xs: (ctypes.c_longlong * len(xs))(*xs)
        for li in range(nl - 1, -1, -1):
            l = eng.layers[li]
            par = li & 1
            dY = self.dY2[par]
            slots, names, (y_h, ld_h, st_h, g_h, b_h, mi_h, mv_h) = self.bn_slot[li]
            nb = len(slots)
            dg_h = vparr([self._param_ptr(eng.grad, nm + "/gamma") for nm in names])
            db_h = vparr([self._param_ptr(eng.grad, nm + "/beta") for nm in names])
            # gradient of branch n goes into its column slice of the source's dYRcat (row stride = ld_h[n])
            dys = [self._p(dY)]
            for n, j in enumerate(l.res_sources):
                jj, col = eng.res_col[(li, n)]
                dys.append(self._p(self.dYRcat[jj], col))
            dy_h = vparr(dys)
            if li in eng.src_of_layer_output:
                dA_ptr, dA_f32 = self._p(self.dres[eng.src_of_layer_output[li]]), 1
            else:
                dA_ptr, dA_f32 = self._p(self.dA), 0
            if li + 2 < nl:
                # this layer's dY buffer was last read by wgrad(li + 2) on the aux stream
                plan.append([_StreamWait(self, "main", ev_wg[li + 2]), []])
            if li in self.fused_red:
                # the two reductions were accumulated by dgrad(li + 1) below (enqueued earlier)
                nm = names[0]
                plan.append([lib.os2s_bn_bwd_apply_p, [self._p(self.Y[li]), self._p(self.mean_invstd[slots[0]]),
                                                       self._param_ptr(eng.master, nm + "/gamma"),
                                                       self._param_ptr(eng.grad, nm + "/gamma"),
                                                       self._param_ptr(eng.grad, nm + "/beta"), self._p(dY), dA_ptr,
                                                       self._p(self.A[li]), self._p(self.fused_red[li]), M, l.c_out,
                                                       _c_float(l.keep), eng.dtypes, st],
                             # apply pass only: dA + a + y in, dy out
                             ("hbm:bn_bwd", float(M) * l.c_out * (2 + 2 + (4 if eng.conv_dtype == "fp32" else 2) + 2))])
            else:
                plan.append([lib.os2s_bn_bwd_p, [nb, y_h, ld_h, mi_h, g_h, dg_h, db_h, dy_h, dA_ptr, dA_f32,
                                                 self._p(self.A[li]), self._p(self.red), M, l.c_out, _c_float(l.keep),
                                                 1, eng.dtypes, st],
                             # reduce pass (dA, a, y_j) + apply pass (dA, a, y_j in, dy_j out)
                             ("hbm:bn_bwd", float(M) * l.c_out * (2 * ((4 if dA_f32 else 2) + 2)
                                                                  + nb * (2 * (4 if eng.conv_dtype == "fp32" else 2) + 2)))])
            self._keep += [dg_h, db_h, dy_h]
            plan.append([_StreamRecord(self, "main", ev_bn[li]), []])
            # the input of this layer is residual source j: every consumer block has written its slice of
            # dYRcat[j] by now (they are all later layers)
            src_j = eng.src_of_layer_output.get(li - 1) if li > 0 else None
            # ---- main stream (critical path) first: data gradients, so that bn_bwd of the next layer can
            # start as soon as they finish while this layer's weight gradients are still running
            if src_j is not None:
                g = eng.res_groups[src_j]
                # all residual branches that read source j: ONE GEMM, reduction over sum_b C_b (first writer
                # of the fp32 accumulator; the main-path gradient below adds to it)
                plan.append([lib.os2s_conv1d_dgrad_p, [self._p(self.dYRcat[src_j]), self._p(eng.wcat[src_j]),
                                                       self._p(self.dres[src_j]), B, T2, g["cj"], g["ntot"], 1, 1, 0, 1,
                                                       rl, eng.dtypes, st], ("dgrad", 2.0 * B * T2 * g["cj"] * g["ntot"])])
            split = l.sep and l.sep_mode == "split"
            dZ = self.dZ2[par] if split else None
            ev_dz = None
            if split:
                # pointwise data gradient dZ = dY . P^T (tcgen05), then the depthwise data gradient (flipped taps)
                plan.append([lib.os2s_conv1d_dgrad_p, [self._p(dY), self._half_ptr(eng.wb, l.wname), self._p(dZ), B, T2,
                                                       l.c_in, l.c_out, 1, 1, 0, eng.grad_out_mode, _vp(0), eng.dtypes, st],
                             ("dgrad", 2.0 * B * T2 * l.c_in * l.c_out)])
                ev_dz = torch.cuda.Event()
                plan.append([_StreamRecord(self, "main", ev_dz), []])   # the depthwise weight gradient (aux) reads dZ
                if li > 0:
                    if src_j is not None:
                        mode, out_ptr = 2, self._p(self.dres[src_j])
                    else:
                        mode, out_ptr = eng.grad_out_mode, self._p(self.dA)
                    plan.append([lib.os2s_depthwise_conv1d,
                                 [self._p(dZ), self._param_ptr(eng.master, l.name + "/depthwise_kernel"), out_ptr, B, T2,
                                  l.c_in, l.K, l.kpad, -l.dil, mode, eng.dtypes, st],
                                 ("hbm:depthwise", float(M) * l.c_in * (2 + (8 if mode == 2 else 2)))])
            elif li > 0:
                if src_j is not None:
                    mode, out_ptr = 2, self._p(self.dres[src_j])
                else:
                    mode, out_ptr = eng.grad_out_mode, self._p(self.dA)
                if (li - 1) in self.fused_red:
                    # dA of the plain layer below + its BN-backward sums in the same kernel
                    lp = eng.layers[li - 1]
                    plan.append([lib.os2s_conv1d_dgrad_bnred_p,
                                 [self._p(dY), self._half_ptr(eng.wb, l.wname), out_ptr, B, T2, l.kC_in,
                                  l.c_out, l.kK, l.dil, l.kpad, self._p(self.A[li - 1]), self._p(self.Y[li - 1]),
                                  _c_float(lp.keep), self._p(self.fused_red[li - 1]), rl, eng.dtypes, st],
                                 ("dgrad", 2.0 * B * T2 * l.K * l.c_in * l.c_out)])
                else:
                    plan.append([lib.os2s_conv1d_dgrad_p, [self._p(dY), self._half_ptr(eng.wb, l.wname),
                                                           out_ptr, B, T2, l.kC_in, l.c_out, l.kK, l.dil, l.kpad, mode,
                                                           rl, eng.dtypes, st],
                                 ("dgrad", 2.0 * B * T2 * l.K * l.c_in * l.c_out)])
            # ---- aux stream: weight gradients of this layer (enqueued after the critical-path kernels)
            plan.append([_StreamWait(self, "aux", ev_bn[li]), []])
            # main conv wgrad (stored layout == kernel layout, also for the folded stride-2 layer)
            x_ptr = self._p(self.A[li - 1]) if li > 0 else self._p(self.feats)
            if split:
                plan.append([lib.os2s_conv1d_wgrad_p, [self._p(self.Z[li]), self._p(dY), self._param_ptr(eng.grad, l.wname),
                                                       B, T2, l.c_in, l.c_out, 1, 1, 0, _vp(0), eng.dtypes, sa],
                             ("wgrad", 2.0 * B * T2 * l.c_in * l.c_out)])
                plan.append([_StreamWait(self, "aux", ev_dz), []])
                plan.append([lib.os2s_depthwise_conv1d_wgrad,
                             [x_ptr, self._p(dZ), self._param_ptr(eng.grad, l.name + "/depthwise_kernel"), B, T2,
                              l.c_in, l.K, l.dil, l.kpad, eng.dtypes, sa]])
            else:
                plan.append([lib.os2s_conv1d_wgrad_p, [x_ptr, self._p(dY), self._param_ptr(eng.grad, l.wname),
                                                       B, T2, l.kC_in, l.c_out, l.kK, l.dil, l.kpad, rl, eng.dtypes, sa],
                             ("wgrad", 2.0 * B * T2 * l.K * l.c_in * l.c_out)])
            if l.fold:
                # structurally-zero taps of the folded stride-2 kernel get no gradient
                s_ = eng.by_name[l.wname]
                st0, n_ = eng._valid_slice(s_)
                zs = []
                if st0 > 0:
                    zs.append(eng.grad[s_["offset"]:s_["offset"] + st0])
                if st0 + n_ < s_["store_size"]:
                    zs.append(eng.grad[s_["offset"] + st0 + n_:s_["offset"] + s_["store_size"]])
                if zs:
                    plan.append([_ZeroSlices(self, zs), []])
            if l.sep and l.sep_mode == "compose":
                # fold the dense weight gradient back onto the depthwise / pointwise variables
                plan.append([lib.os2s_sepconv_decompose_grad, eng._decompose_args(l.wname) + [sa]])
            if src_j is not None:
                # weight gradients of all 1x1 kernels that read source j: one GEMM into [C_j, Ntot_j], then
                # scattered to the per-variable gradient buffers (which live in this layer's region)
                g = eng.res_groups[src_j]
                plan.append([lib.os2s_conv1d_wgrad_p, [x_ptr, self._p(self.dYRcat[src_j]), self._p(eng.dwcat[src_j]),
                                                       B, T2, g["cj"], g["ntot"], 1, 1, 0, rl, eng.dtypes, sa],
                             ("wgrad", 2.0 * B * T2 * g["cj"] * g["ntot"])])
                src, dst, rows, rbytes, sp, dp = [], [], [], [], [], []
                for (lc, n, cb, col) in g["consumers"]:
                    rn = eng.res_wname(eng.layers[lc], n)
                    src.append(self._p(eng.dwcat[src_j], col))
                    dst.append(self._param_ptr(eng.grad, rn))
                    rows.append(g["cj"])
                    rbytes.append(cb * 4)
                    sp.append(g["ntot"] * 4)
                    dp.append(cb * 4)
                for i in range(0, len(src), 64):
                    sl = slice(i, i + 64)
                    a = (vparr(src[sl]), vparr(dst[sl]), iarr(rows[sl]), iarr(rbytes[sl]), llarr(sp[sl]), llarr(dp[sl]))
                    self._keep.append(a)
                    plan.append([lib.os2s_multi_copy_2d, [len(src[sl])] + list(a) + [sa]])
                for (lc, n, cb, col) in g["consumers"]:
                    if eng.layers[lc].sep:
                        plan.append([lib.os2s_sepconv_decompose_grad,
                                     eng._decompose_args(eng.res_wname(eng.layers[lc], n)) + [sa]])
            plan.append([_StreamRecord(self, "aux", ev_wg[li]), []])
            if n_bucket < len(buckets) and buckets[n_bucket][0] == eng.layer_first_offset(li):
                # the bucket also holds weight gradients produced on the aux stream
                b0, b1 = buckets[n_bucket]
                plan.append([_StreamWait(self, "main", ev_wg[li]), []])
                plan.append([_BucketAllReduce(eng, n_bucket, eng.grad[b0:b1]), []])
                n_bucket += 1
        # join: the optimizer needs every weight gradient
        plan.append([_StreamWait(self, "main", ev_wg[0]), []])
        if nl > 1:
            plan.append([_StreamWait(self, "main", ev_wg[1]), []])
        self._bwd_plan = plan
        self._bwd_L = L_max
        self.n_launch_bwd = len(plan) + 4

    def run_loss_only(self, labels, label_lens):
        self.set_targets(labels, label_lens)
        self._st.value = torch.cuda.current_stream().cuda_stream
        fn, args = self._bwd_plan[0][0], self._bwd_plan[0][1]
        L.check(fn(*args), "os2s_ctc_loss_fwd_bwd")
        return self.loss

    def run_backward(self, labels, label_lens, dlogits=None):
        if dlogits is None:
            self.set_targets(labels, label_lens)
            plan = self._bwd_plan
        else:
            if self._bwd_plan is None:
                self._build_backward_plan(32)
            self.dlogits.copy_(dlogits)
            plan = self._bwd_plan[1:]  # skip the CTC launch
        self._body_backward(plan)

    def run_train_step(self, feats, feat_lens, labels, label_lens):
        """Whole training step.  After two eager steps of a given shape the step body (forward, CTC,
        backward incl. bucketed all-reduce, optimizer) is captured into a CUDA graph and replayed."""
        eng = self.eng
        self.set_inputs(feats, feat_lens)
        self.set_targets(labels, label_lens)
        eng._last_ws = self
        if eng.peer is not None and eng.step_count % 128 == 127 and eng.peer.timed_out():
            raise RuntimeError("peer-memory gradient exchange: a rank did not signal within the time-out "
                               "(OS2S_PEER_TIMEOUT_S); the gradients of the last steps are not rank sums")
        if eng.iter_size > 1:
            return self._accumulating_step()
        if eng.use_cuda_graph and eng._profile is None and (eng.comm is None or eng.graph_with_comm
                                                            or (eng.peer is not None and eng.graph_with_peer)):
            if self.graph is not None:
                self.graph.replay()
                eng.step_count += 1
                return self.loss
            if self._eager_steps >= 2:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                # thread_local: the data layer's producer thread keeps allocating / synchronising events on its
                # side stream while this thread captures (the default "global" mode would invalidate the capture)
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._step_body()
                self.graph = g
                g.replay()
                eng.step_count += 1
                return self.loss
        self._step_body()
        self._eager_steps += 1
        return self.loss

    def _accumulating_step(self):
        """iter_size > 1 (optimizers.py:212-259): every call adds g / iter_size to the accumulator; the
        iter_size-th call reduces the ACCUMULATED gradient over the ranks, runs LARC + the optimizer on it
        and clears the accumulator.  The loss scale is constant inside a window (it only changes in the
        optimizer step), so scaled gradients are accumulated and unscaled once.  Runs from the eager plan."""
        eng = self.eng
        eng._suppress_comm = True   # no per-bucket all-reduce of the micro-step gradients
        try:
            self._body_forward()
            self.run_decoder()
            self._body_backward(self._bwd_plan)
        finally:
            eng._suppress_comm = False
        eng.grad_acc.add_(eng.grad, alpha=1.0 / eng.iter_size)
        eng._micro += 1
        if eng._micro < eng.iter_size:
            eng.istate[5] += 1      # attempted-step counter: a fresh dropout stream for the next micro-step
            return self.loss
        eng._micro = 0
        eng.grad.copy_(eng.grad_acc)
        eng.grad_acc.zero_()
        if eng.comm is not None:
            eng.comm.allreduce_(eng.grad)
        eng._launch_optimizer()
        return self.loss

    def _step_body(self):
        self._body_forward()
        self.run_decoder()
        self._body_backward(self._bwd_plan)
        self.eng._launch_optimizer()
