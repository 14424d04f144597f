"""ORACLE (test infrastructure only -- never imported by the product path).

PyTorch-CPU fp32 "fast twin" of the NumPy oracle: the same graph as oracle/encoder.py +
oracle/ctc.py + oracle/optimizer.py, expressed with torch CPU ops so that (a) autograd supplies the
backward pass the reference gets from TF autodiff, and (b) it is fast enough to be timed as the
CPU baseline in bench.py (`cpu_baseline.kind = "port"`; TF1 cannot be installed offline).

Follows the same reference lines as the NumPy oracle:
  open_seq2seq/encoders/tdnn_encoder.py:87-265, open_seq2seq/parts/cnns/conv_blocks.py:61-232,
  open_seq2seq/decoders/fc_decoders.py:105-158, open_seq2seq/losses/ctc_loss.py:44-89,
  open_seq2seq/optimizers/{optimizers.py:289-378,novograd.py:93-126,mp_wrapper.py:44-122}.
It is validated against the NumPy restatement in tests/test_oracle.py, and -- in the build container, where
/root/reference exists -- against the reference's OWN encoder / decoder / loss builders executed over an eager
stand-in for their TensorFlow symbols (tests/test_reference_encoder_executed_cpu.py: same logits to 1e-10, same
lengths, same variable set).
"""
import math

import torch
import torch.nn.functional as F

from .encoder import same_padding


def conv1d_same(x, w, stride=1, dilation=1):
    """x [B,T,Cin], w [K,Cin,Cout] (TF layout) -> [B,T_out,Cout]."""
    B, T, C = x.shape
    K = w.shape[0]
    _, pl, pr = same_padding(T, K, stride, dilation)
    xin = F.pad(x.transpose(1, 2), (pl, pr))
    y = F.conv1d(xin, w.permute(2, 1, 0), stride=stride, dilation=dilation)
    return y.transpose(1, 2)


def sep_conv1d_same(x, depthwise, pointwise, stride=1, dilation=1):
    """tf.layers.separable_conv1d(use_bias=False, padding='SAME', depth_multiplier=1) as conv_blocks.py:27-40 calls
    it: x [B,T,C], depthwise_kernel [K,C,1], pointwise_kernel [1,C,Cout] -> [B,T_out,Cout].  The depthwise stage
    is a grouped cross-correlation with SAME padding computed from (K, stride, dilation) like conv1d."""
    B, T, C = x.shape
    K = depthwise.shape[0]
    _, pl, pr = same_padding(T, K, stride, dilation)
    xin = F.pad(x.transpose(1, 2), (pl, pr))
    z = F.conv1d(xin, depthwise.permute(1, 2, 0), stride=stride, dilation=dilation, groups=C)   # [B,C,T_out]
    return z.transpose(1, 2) @ pointwise[0]


def layer_conv(x, params, name, layer, stride, dilation):
    """Main / residual convolution of a block: conv1d or sep_conv1d by the layer's type (the residual convs of
    a sep_conv1d block are separable too, conv_blocks.py:66,79-85)."""
    if layer.get("type", "conv1d") == "sep_conv1d":
        return sep_conv1d_same(x, params[name + "/depthwise_kernel"], params[name + "/pointwise_kernel"], stride, dilation)
    return conv1d_same(x, params[name + "/kernel"], stride, dilation)


def batch_norm_train(x, gamma, beta, eps=1e-3):
    mean = x.mean(dim=(0, 1))
    var = x.var(dim=(0, 1), unbiased=False)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta, mean, var


def sequence_mask(lengths, maxlen, dtype):
    return (torch.arange(maxlen)[None, :] < lengths[:, None]).to(dtype)[:, :, None]


def round_like_device(kind):
    """Storage-rounding emulation for the 'what if the reference stored activations like the B200
    path does' comparison: conv outputs fp16, layer outputs bf16 (straight-through gradient)."""
    dt = torch.float16 if kind == "conv" else torch.bfloat16
    return lambda t: t.to(dt).to(t.dtype)


def tdnn_encode(x, src_len, layers, params, training=True, bn_eps=1e-3, use_conv_mask=True,
                dropout_masks=None, collect=None, stats=None, emulate_storage=False):
    """Same contract as oracle.encoder.tdnn_encode, on torch tensors (autograd-capable).
    emulate_storage=True rounds conv outputs to fp16 and layer outputs to bf16 where the device
    path stores them (arithmetic stays in the tensor's dtype)."""
    rc_ = round_like_device("conv") if emulate_storage else (lambda t: t)
    ra_ = round_like_device("act") if emulate_storage else (lambda t: t)
    src_len = src_len.clone()
    max_len = x.shape[1]
    mask = sequence_mask(src_len, max_len, x.dtype) if use_conv_mask else None
    feats = x
    res_agg = []
    for bi, layer in enumerate(layers):
        K = layer["kernel_size"][0]
        stride = layer["stride"][0]
        dil = layer["dilation"][0]
        residual = layer.get("residual", False)
        dense = layer.get("residual_dense", False)
        if use_conv_mask:
            feats = feats * mask
        layer_res = None
        if residual:
            layer_res = feats
            if dense:
                res_agg.append(layer_res)
                layer_res = list(res_agg)
            else:
                layer_res = [layer_res]
        for ri in range(layer["repeat"]):
            name = "conv%d%d" % (bi + 1, ri + 1)
            src_len = (src_len + stride - 1) // stride
            max_len = (max_len + stride - 1) // stride
            if ri > 0 and use_conv_mask:
                feats = feats * mask
            if use_conv_mask and stride > 1:
                mask = sequence_mask(src_len, max_len, x.dtype)
            conv = rc_(layer_conv(feats, params, name, layer, stride, dil))
            if collect is not None:
                collect[name + "/conv"] = conv
            bn, mean, var = batch_norm_train(conv, params[name + "/bn/gamma"], params[name + "/bn/beta"], bn_eps)
            if stats is not None:
                stats[name + "/bn"] = (mean.detach(), var.detach(), conv.shape[0] * conv.shape[1])
            if residual and ri == layer["repeat"] - 1:
                for j, res in enumerate(layer_res):
                    rname = (name + "/res_%d" % j) if dense else (name + "/res")
                    bname = (name + "/res_bn_%d" % j) if dense else (name + "/res_bn")
                    rc = rc_(layer_conv(res, params, rname, layer, 1, 1))
                    rb, mean, var = batch_norm_train(rc, params[bname + "/gamma"], params[bname + "/beta"], bn_eps)
                    if stats is not None:
                        stats[bname] = (mean.detach(), var.detach(), rc.shape[0] * rc.shape[1])
                    bn = bn + rb
            out = torch.relu(bn)
            if training and dropout_masks is not None and name in dropout_masks:
                m, keep = dropout_masks[name]
                out = out * m / keep
            if use_conv_mask and not (bi == len(layers) - 1 and ri == layer["repeat"] - 1):
                # the mask the NEXT layer applies to its input (tdnn_encoder.py:185-186,204-205); applying
                # it here is the same function and makes `feats` the tensor the device path stores
                out = out * mask
            feats = ra_(out)
            if collect is not None:
                collect[name + "/out"] = feats
    return feats, src_len


def fc_decode(enc_out, kernel, bias):
    B, T, H = enc_out.shape
    logits = enc_out.reshape(B * T, H) @ kernel + bias
    return logits.reshape(B, T, -1).transpose(0, 1)  # [T,B,V]


def ctc_loss_mean(logits_tbv, labels, label_lens, input_lens):
    """CTCLoss._compute_loss: mean over batch of per-utterance NLL, blank = V-1,
    ignore_longer_outputs_than_inputs (zero_infinity) and NaN masking."""
    V = logits_tbv.shape[2]
    lp = F.log_softmax(logits_tbv.float(), dim=2)
    per = F.ctc_loss(lp, labels, input_lens, label_lens, blank=V - 1, reduction="none", zero_infinity=True)
    per = torch.where(torch.isnan(per), torch.zeros_like(per), per)
    return per.mean()


def xavier_init(shape, uniform, gen):
    """tf.contrib.layers.xavier_initializer: n = (fan_in + fan_out)/2."""
    if len(shape) == 3:
        fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
    else:
        fan_in, fan_out = shape
    n = (fan_in + fan_out) / 2.0
    if uniform:
        lim = math.sqrt(3.0 / n)
        return (torch.rand(shape, generator=gen) * 2 - 1) * lim
    std = math.sqrt(1.3 / n)
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t


def init_params(layers, num_features, vocab, seed=0):
    """Parameter dict with the reference's variable names."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    c_in = num_features
    block_inputs = []
    for bi, layer in enumerate(layers):
        K = layer["kernel_size"][0]
        c_out = layer["num_channels"]
        residual = layer.get("residual", False)
        dense = layer.get("residual_dense", False)
        if residual:
            if dense:
                block_inputs.append(c_in)
                res_ch = list(block_inputs)
            else:
                res_ch = [c_in]
        for ri in range(layer["repeat"]):
            name = "conv%d%d" % (bi + 1, ri + 1)
            sep = layer.get("type", "conv1d") == "sep_conv1d"
            if sep:
                p[name + "/depthwise_kernel"] = xavier_init((K, c_in, 1), False, gen)
                p[name + "/pointwise_kernel"] = xavier_init((1, c_in, c_out), False, gen)
            else:
                p[name + "/kernel"] = xavier_init((K, c_in, c_out), False, gen)
            p[name + "/bn/gamma"] = torch.ones(c_out)
            p[name + "/bn/beta"] = torch.zeros(c_out)
            if residual and ri == layer["repeat"] - 1:
                for j, rc in enumerate(res_ch):
                    rname = (name + "/res_%d" % j) if dense else (name + "/res")
                    bname = (name + "/res_bn_%d" % j) if dense else (name + "/res_bn")
                    if sep:
                        p[rname + "/depthwise_kernel"] = xavier_init((1, rc, 1), False, gen)
                        p[rname + "/pointwise_kernel"] = xavier_init((1, rc, c_out), False, gen)
                    else:
                        p[rname + "/kernel"] = xavier_init((1, rc, c_out), False, gen)
                    p[bname + "/gamma"] = torch.ones(c_out)
                    p[bname + "/beta"] = torch.zeros(c_out)
            c_in = c_out
    p["fc/kernel"] = xavier_init((c_in, vocab), True, gen)
    p["fc/bias"] = torch.zeros(vocab)
    return p


def forward_loss(params, layers, feats, feat_len, labels, label_lens, training=True, dropout_masks=None,
                 collect=None, emulate_storage=False):
    enc, out_len = tdnn_encode(feats, feat_len, layers, params, training=training,
                               dropout_masks=dropout_masks, collect=collect, emulate_storage=emulate_storage)
    logits = fc_decode(enc, params["fc/kernel"], params["fc/bias"])
    loss = ctc_loss_mean(logits, labels, label_lens, out_len)
    return loss, logits, out_len


def larc_novograd_step(params, grads, momentum, lr, beta1=0.95, beta2=0.98, epsilon=1e-8,
                       weight_decay=0.001, larc_eta=0.001, min_update=1e-7, larc_eps=1e-7):
    """optimizers.py:333-377 (clip mode) then novograd.py:93-126 (reference-as-written EMA)."""
    with torch.no_grad():
        for name, w in params.items():
            g = grads[name]
            v_norm = w.norm()
            g_norm = g.norm()
            r = torch.clamp(larc_eta * v_norm / (lr * (g_norm + larc_eps)), min=min_update, max=1.0)
            g = g * r
            g2 = (g * g).sum()
            g = g / torch.sqrt(g2 + epsilon)
            g = g + weight_decay * w
            m = momentum.get(name)
            m = g if m is None else beta1 * m + g
            momentum[name] = m
            w -= lr * m


def backward_with_saved_forward(params, layers, feats, feat_len, saved_conv, saved_out, dlogits_btv,
                                bn_eps=1e-3, use_conv_mask=True):
    """Reference backward pass (what TF autodiff computes) evaluated AT A GIVEN FORWARD STATE.

    saved_conv: name -> conv output Y (main: 'conv21', residual branch n: 'conv25/res_0'), saved_out:
    name -> layer output A.  ReLU gates and BN statistics are recomputed from these saved tensors
    with the oracle's own op definitions (conv1d_same, batch_norm_train), layer by layer with local
    autograd, so the result isolates the backward arithmetic from the forward pass's sensitivity
    (a ReLU network's gradient is discontinuous in its forward values).  Returns name -> gradient."""
    grads = {}
    p = {k: v.detach().double() for k, v in params.items()}
    # ---- topology walk (forward order)
    plan = []
    src_len = feat_len.clone()
    max_len = feats.shape[1]
    masks = {}
    res_agg = []
    prev = None
    for bi, layer in enumerate(layers):
        residual = layer.get("residual", False)
        dense = layer.get("residual_dense", False)
        stride = layer["stride"][0]
        layer_res = None
        if residual:
            if dense:
                res_agg.append(prev)
                layer_res = list(res_agg)
            else:
                layer_res = [prev]
        for ri in range(layer["repeat"]):
            name = "conv%d%d" % (bi + 1, ri + 1)
            src_len = (src_len + stride - 1) // stride
            max_len = (max_len + stride - 1) // stride
            last = bi == len(layers) - 1 and ri == layer["repeat"] - 1
            m = sequence_mask(src_len, max_len, torch.float64) if (use_conv_mask and not last) else None
            end = residual and ri == layer["repeat"] - 1
            plan.append({"name": name, "in": prev, "K": layer["kernel_size"][0], "stride": stride,
                         "dil": layer["dilation"][0], "mask": m, "res": layer_res if end else [], "dense": dense,
                         "layer": layer})
            prev = name
    # ---- FC
    enc = saved_out[plan[-1]["name"]].detach().double().requires_grad_(True)
    wk = p["fc/kernel"].clone().requires_grad_(True)
    bk = p["fc/bias"].clone().requires_grad_(True)
    logits = fc_decode(enc, wk, bk).transpose(0, 1)
    logits.backward(dlogits_btv.double())
    grads["fc/kernel"], grads["fc/bias"] = wk.grad, bk.grad
    dA = {plan[-1]["name"]: enc.grad}
    # ---- layers in reverse
    for node in reversed(plan):
        name = node["name"]
        d_out = dA.pop(name)
        ys = [saved_conv[name].detach().double().requires_grad_(True)]
        gs = [p[name + "/bn/gamma"].clone().requires_grad_(True)]
        bs = [p[name + "/bn/beta"].clone().requires_grad_(True)]
        bnames = [name + "/bn"]
        for n, src in enumerate(node["res"]):
            bn_name = (name + "/res_bn_%d" % n) if node["dense"] else (name + "/res_bn")
            cn = (name + "/res_%d" % n) if node["dense"] else (name + "/res")
            ys.append(saved_conv[cn].detach().double().requires_grad_(True))
            gs.append(p[bn_name + "/gamma"].clone().requires_grad_(True))
            bs.append(p[bn_name + "/beta"].clone().requires_grad_(True))
            bnames.append(bn_name)
        tot = 0
        for y, g, b in zip(ys, gs, bs):
            tot = tot + batch_norm_train(y, g, b, bn_eps)[0]
        out = torch.relu(tot)
        if node["mask"] is not None:
            out = out * node["mask"]
        out.backward(d_out)
        for bn_name, g, b in zip(bnames, gs, bs):
            grads[bn_name + "/gamma"], grads[bn_name + "/beta"] = g.grad, b.grad
        # main conv
        x_src = feats if node["in"] is None else saved_out[node["in"]]
        x = x_src.detach().double().requires_grad_(node["in"] is not None)
        wnames = [k for k in (name + "/kernel", name + "/depthwise_kernel", name + "/pointwise_kernel") if k in p]
        wl = {k: p[k].clone().requires_grad_(True) for k in wnames}
        layer_conv(x, wl, name, node["layer"], node["stride"], node["dil"]).backward(ys[0].grad)
        for k in wnames:
            grads[k] = wl[k].grad
        if node["in"] is not None:
            dA[node["in"]] = dA.get(node["in"], 0) + x.grad
        # residual 1x1 convs
        for n, src in enumerate(node["res"]):
            cn = (name + "/res_%d" % n) if node["dense"] else (name + "/res")
            xr = saved_out[src].detach().double().requires_grad_(True)
            rnames = [k for k in (cn + "/kernel", cn + "/depthwise_kernel", cn + "/pointwise_kernel") if k in p]
            wr = {k: p[k].clone().requires_grad_(True) for k in rnames}
            layer_conv(xr, wr, cn, node["layer"], 1, 1).backward(ys[1 + n].grad)
            for k in rnames:
                grads[k] = wr[k].grad
            dA[src] = dA.get(src, 0) + xr.grad
    return grads
