"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy (float64) restatement of the reference's TDNNEncoder forward pass and FC decoder:

  open_seq2seq/encoders/tdnn_encoder.py:87-265      _encode: masks, length bookkeeping, block loop,
                                                    dense residual aggregation, dropout
  open_seq2seq/parts/cnns/conv_blocks.py:170-232    conv_bn_actv
  open_seq2seq/parts/cnns/conv_blocks.py:61-168     conv_bn_res_bn_actv (dense residual)
  open_seq2seq/decoders/fc_decoders.py:105-158      FullyConnectedTimeDecoder._decode

The ops themselves are TensorFlow 1.x (not vendored / not installed); restated semantics
(SURVEY.md Appendix A, A1-A6):
  conv1d SAME: T_out = ceil(T/s); pad_total = max((T_out-1)*s + (K-1)*d + 1 - T, 0); left = total//2
               (extra on the right); cross-correlation; kernel [K, C_in, C_out]; no bias.
  batch_normalization(training=True, momentum, eps): batch mean / biased variance over all B*T
               positions (including masked ones); moving stats <- m*moving + (1-m)*batch, the variance
               fed to the moving average carries Bessel's correction N/(N-1) (fused BN).
  dropout: identity when keep_prob == 1 (parity runs); otherwise x * mask / keep with an injected mask.
Parity status: unpinned by the reference's own tests (no golden vectors for conv/BN); triangulated
against torch.nn.functional.{conv1d,batch_norm} in tests/test_oracle_encoder.py.
"""
import numpy as np


def same_padding(T_in, K, stride, dilation):
    T_out = -(-T_in // stride)
    k_eff = (K - 1) * dilation + 1
    total = max((T_out - 1) * stride + k_eff - T_in, 0)
    left = total // 2
    return T_out, left, total - left


def conv1d_same(x, w, stride=1, dilation=1):
    """x [B,T,C_in], w [K,C_in,C_out] -> [B,T_out,C_out] (tf.layers.conv1d, SAME, no bias)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    B, T, C = x.shape
    K, Ci, Co = w.shape
    assert Ci == C
    T_out, pl, pr = same_padding(T, K, stride, dilation)
    xp = np.zeros((B, T + pl + pr, C))
    xp[:, pl:pl + T] = x
    y = np.zeros((B, T_out, Co))
    for k in range(K):
        sl = xp[:, k * dilation:k * dilation + (T_out - 1) * stride + 1:stride]
        y += sl @ w[k]
    return y


def batch_norm_train(x, gamma, beta, eps=1e-3):
    """Returns (y, mean, biased_var) over axes (B,T)."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=(0, 1))
    var = x.var(axis=(0, 1))
    y = (x - mean) / np.sqrt(var + eps) * gamma + beta
    return y, mean, var


def update_moving(moving_mean, moving_var, mean, var, n, momentum=0.9):
    unbiased = var * (n / max(n - 1.0, 1.0))
    return (moving_mean * momentum + mean * (1 - momentum),
            moving_var * momentum + unbiased * (1 - momentum))


def sequence_mask(lengths, maxlen):
    return (np.arange(maxlen)[None, :] < np.asarray(lengths)[:, None]).astype(np.float64)[:, :, None]


def relu(x):
    return np.maximum(x, 0.0)


def tdnn_encode(x, src_len, layers, params, training=True, activation=relu, bn_eps=1e-3,
                use_conv_mask=True, dropout_masks=None, collect=None):
    """TDNNEncoder._encode forward (tdnn_encoder.py:109-265), channels_last, batch_norm.

    x [B,T,F] already padded (pad_to applied by the data layer); src_len [B].
    layers : the config's convnet_layers list.
    params : dict name -> array with TF variable names relative to the encoder scope, e.g.
             'conv11/kernel', 'conv11/bn/gamma', 'conv25/res_0/kernel', 'conv25/res_bn_0/beta'.
    dropout_masks : optional dict 'conv{b}{r}' -> {0,1} mask [B,T,C] and keep prob applied as
             x*mask/keep (tf.nn.dropout with an injected mask); None -> keep_prob = 1.
    collect : optional dict filled with intermediate tensors (conv outputs, activations).
    Returns (outputs [B,T',C], out_len [B]).
    """
    x = np.asarray(x, dtype=np.float64)
    src_len = np.asarray(src_len).astype(np.int64)
    max_len = x.shape[1]
    mask = sequence_mask(src_len, max_len) if use_conv_mask else None
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
                mask = sequence_mask(src_len, max_len)
            conv = conv1d_same(feats, params[name + "/kernel"], stride, dil)
            if collect is not None:
                collect[name + "/conv"] = conv
            bn, _, _ = batch_norm_train(conv, params[name + "/bn/gamma"], params[name + "/bn/beta"], bn_eps)
            if residual and ri == layer["repeat"] - 1:
                total = 0.0
                for j, res in enumerate(layer_res):
                    rname = (name + "/res_%d" % j) if dense else (name + "/res")
                    bname = (name + "/res_bn_%d" % j) if dense else (name + "/res_bn")
                    rc = conv1d_same(res, params[rname + "/kernel"], 1, 1)
                    rb, _, _ = batch_norm_train(rc, params[bname + "/gamma"], params[bname + "/beta"], bn_eps)
                    total = total + rb
                bn = bn + total
            out = activation(bn)
            if training and dropout_masks is not None and name in dropout_masks:
                m, keep = dropout_masks[name]
                out = out * m / keep
            feats = out
            if collect is not None:
                collect[name + "/out"] = feats
    return feats, src_len


def fc_decode(enc_out, kernel, bias):
    """FullyConnectedTimeDecoder._decode: [B,T,H] -> time-major logits [T,B,V] (fc_decoders.py:126-148)."""
    B, T, H = enc_out.shape
    logits = enc_out.reshape(B * T, H) @ np.asarray(kernel, dtype=np.float64) + np.asarray(bias, dtype=np.float64)
    return logits.reshape(B, T, -1).transpose(1, 0, 2)
