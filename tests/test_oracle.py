"""CPU tests: the oracle is pinned against the reference's own golden vectors / known answers and
triangulated against independent implementations (torch, torchaudio)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc as OC
from oracle import encoder as OE
from oracle import featurizer as FZ
from oracle import optimizer as OO
from oracle import torch_twin as TT


def test_ctc_oracle_reproduces_reference_golden_vector(golden_dir):
    # known answers: /root/reference ctc_decoder_with_lm/ctc-test.py:64-67 (greedy), :73 (log prob)
    g = np.load(os.path.join(golden_dir, "ctc_test_logits.npz"))
    lg, vocab = g["logits"], list(g["vocab"])
    toks, score = OC.ctc_greedy_decode(lg, [lg.shape[0]])
    assert "".join(vocab[c] for c in toks[0]) == "then seconds"
    assert abs(-score[0] - 7079.117) < 1e-3
    lab = np.array([[vocab.index(c) for c in "then seconds"]])
    loss, _ = OC.ctc_loss_and_grad(lg, lab, [lab.shape[1]], [lg.shape[0]])
    assert abs(loss[0] - 1.1842575) < 1e-3


def test_ctc_oracle_matches_torch_ctc_including_edge_cases():
    rng = np.random.default_rng(0)
    T, B, V, Lmax = 40, 4, 29, 12
    lg = rng.standard_normal((T, B, V))
    labels = rng.integers(0, V - 1, size=(B, Lmax))
    labels[0, :4] = [2, 2, 2, 9]
    label_lens = [12, 5, 0, 12]
    in_lens = [40, 23, 9, 8]  # last is infeasible (12 labels in 8 frames)
    loss, grad = OC.ctc_loss_and_grad(lg, labels, label_lens, in_lens)
    x = torch.tensor(lg, requires_grad=True)
    ref = torch.nn.functional.ctc_loss(x.log_softmax(2), torch.tensor(labels), torch.tensor(in_lens),
                                       torch.tensor(label_lens), blank=V - 1, reduction="none", zero_infinity=True)
    ref.sum().backward()
    assert loss[3] == 0.0
    assert np.abs(loss - ref.detach().numpy()).max() < 1e-9
    assert np.abs(grad - x.grad.numpy()).max() < 1e-9
    mean, gm = OC.ctc_loss_mean(lg, labels, label_lens, in_lens)
    assert abs(mean - loss.mean()) < 1e-12 and np.allclose(gm, grad / B)


def test_mel_filterbank_and_stft_match_independent_implementations():
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 64, 16000, norm="slaney", mel_scale="slaney").T
    assert np.abs(fb.numpy() - FZ.mel_filterbank()).max() < 1e-6
    rng = np.random.default_rng(0)
    s = rng.standard_normal(5000)
    win = torch.hann_window(320, periodic=False, dtype=torch.float64)
    S = torch.stft(torch.tensor(s), n_fft=512, hop_length=160, win_length=320, window=win, center=True,
                   pad_mode="reflect", return_complex=True).abs() ** 2
    S2 = FZ.stft_power(s)
    assert S2.shape == (257, 1 + 5000 // 160)
    assert np.abs(S.numpy() - S2).max() < 1e-5 * S2.max()


def test_featurizer_shape_and_normalisation_like_reference_test():
    # mirrors data/speech2text/speech_utils_test.py:45-85 (shape, mean 0 / std 1) for the librosa backend
    rng = np.random.default_rng(1)
    sig = np.clip(3000 * rng.standard_normal(43200), -32768, 32767).astype(np.int16)
    f, dur = FZ.logfbank_features(sig)
    assert f.shape == (1 + 43200 // 160, 64) and abs(dur - 2.7) < 1e-9
    assert np.abs(f.mean(0)).max() < 1e-6 and np.abs(f.std(0) - 1).max() < 1e-6
    batch, lens = FZ.batch_features([sig, sig[:16000]], pad_to=16)
    assert batch.shape[1] % 16 == 0 and lens.tolist() == [271, 101]
    assert np.all(batch[1, 101:] == 0)


def _mini_layers():
    return [
        {"type": "conv1d", "repeat": 1, "kernel_size": [11], "stride": [2], "num_channels": 16, "padding": "SAME",
         "dilation": [1]},
        {"type": "conv1d", "repeat": 2, "kernel_size": [5], "stride": [1], "num_channels": 16, "padding": "SAME",
         "dilation": [1], "residual": True, "residual_dense": True},
        {"type": "conv1d", "repeat": 2, "kernel_size": [7], "stride": [1], "num_channels": 24, "padding": "SAME",
         "dilation": [1], "residual": True, "residual_dense": True},
        {"type": "conv1d", "repeat": 1, "kernel_size": [9], "stride": [1], "num_channels": 32, "padding": "SAME",
         "dilation": [2]},
    ]


def test_same_padding_rule_matches_tf_examples():
    # conv11 K11 s2 on even T=1504 -> 4/5; stride 1 symmetric; K29 d2 -> 28/28
    assert OE.same_padding(1504, 11, 2, 1) == (752, 4, 5)
    assert OE.same_padding(1503, 11, 2, 1) == (752, 5, 5)
    assert OE.same_padding(752, 29, 1, 2) == (752, 28, 28)
    assert OE.same_padding(752, 1, 1, 1) == (752, 0, 0)


def test_numpy_encoder_matches_torch_twin_and_torch_conv():
    layers = _mini_layers()
    p = TT.init_params(layers, 8, 29, seed=1)
    torch.manual_seed(0)
    x = torch.randn(3, 48, 8, dtype=torch.float64)
    lens = torch.tensor([48, 30, 17])
    x = x * TT.sequence_mask(lens, 48, x.dtype)
    y_t, l_t = TT.tdnn_encode(x, lens, layers, {k: v.double() for k, v in p.items()})
    y_n, l_n = OE.tdnn_encode(x.numpy(), lens.numpy(), layers, {k: v.numpy() for k, v in p.items()})
    assert l_t.tolist() == l_n.tolist() == [24, 15, 9]
    assert np.abs(y_t.numpy() - y_n).max() < 1e-10
    # batch norm against torch's own implementation (biased variance, eps inside the sqrt)
    c = torch.randn(4, 10, 6, dtype=torch.float64)
    g, b = torch.rand(6, dtype=torch.float64) + 0.5, torch.randn(6, dtype=torch.float64)
    ref = torch.nn.functional.batch_norm(c.reshape(-1, 6), None, None, g, b, training=True, eps=1e-3).reshape(4, 10, 6)
    got, mean, var = OE.batch_norm_train(c.numpy(), g.numpy(), b.numpy(), 1e-3)
    assert np.abs(got - ref.numpy()).max() < 1e-12
    mm, mv = OE.update_moving(np.zeros(6), np.ones(6), mean, var, 40, 0.9)
    assert np.allclose(mv, 0.9 + 0.1 * c.reshape(-1, 6).var(0, unbiased=True).numpy())


def test_poly_decay_and_backoff_scaler_trajectories():
    assert OO.poly_decay(0, 0.02, 100, power=2.0, min_lr=1e-5) == pytest.approx(0.02)
    assert OO.poly_decay(50, 0.02, 100, power=2.0, min_lr=1e-5) == pytest.approx((0.02 - 1e-5) * 0.25 + 1e-5)
    assert OO.poly_decay(500, 0.02, 100, power=2.0, min_lr=1e-5) == pytest.approx(1e-5)
    assert OO.poly_decay(5, 0.02, 100, warmup_steps=10, begin_decay_at=20) == pytest.approx(0.01)
    s = OO.BackoffScaler(step_window=4)
    assert s.scale == 2.0 ** 14
    assert s.update(False, 1.0) is False and s.scale == 2.0 ** 14     # iteration 0: since = 1
    assert s.update(True, 1.0) is True and s.scale == 2.0 ** 13        # overflow at iteration 1
    for _ in range(3):
        s.update(False, 1.0)
    assert s.scale == 2.0 ** 13
    s.update(False, 1.0)                                               # iteration 5: since = 4 -> grow
    assert s.scale == 2.0 ** 14
    assert s.update(False, np.inf) is True


def test_mp_wrapper_unscale_semantics_like_reference_test():
    # optimizers/mp_wrapper_test.py:61-95: a gradient of 1e-8 survives because it is unscaled in fp32
    w = [np.ones(4, dtype=np.float32)]
    scaler = OO.BackoffScaler()
    g_scaled = [np.full(4, 1e-8 * scaler.scale, dtype=np.float32)]
    st = OO.NovoGradState(1)
    before = w[0].copy()
    skipped, lr, step = OO.train_step(w, [g_scaled], st, scaler, 0, lambda s: 0.1,
                                      dict(beta1=0.0, beta2=0.5, epsilon=1e-8, weight_decay=0.0))
    assert not skipped and step == 1
    # the 1e-8 gradient is not flushed to zero: update = lr * g / sqrt(||g||^2 + eps)
    expect = 0.1 * 1e-8 / np.sqrt(4e-16 + 1e-8)
    assert np.allclose(before - w[0], expect, rtol=2e-2) and expect > 0


def test_novograd_as_written_vs_corrected_ema():
    rng = np.random.default_rng(0)
    w1 = [rng.standard_normal(8).astype(np.float32)]
    w2 = [w1[0].copy()]
    s1, s2 = OO.NovoGradState(1), OO.NovoGradState(1)
    for i in range(3):
        g = [(rng.standard_normal(8) * (i + 1)).astype(np.float32)]
        OO.novograd_step(w1, g, s1, 0.01, ema_persist=False)
        OO.novograd_step(w2, g, s2, 0.01, ema_persist=True)
    assert s1.ema[0] == 0.0 and s2.ema[0] > 0.0
    assert not np.allclose(w1[0], w2[0])


def test_lr_policies_and_adam_against_closed_forms():
    """cosine / exponential policies as tf.train.{cosine,exponential}_decay define them (the reference
    passes min_lr as cosine_decay's alpha) and Adam against torch.optim.Adam (eps -> 0 limit, where the
    TF and torch placements of epsilon coincide)."""
    import torch
    from oracle import optimizer as OO
    assert OO.cosine_decay(0, 1.0, 10) == 1.0
    assert abs(OO.cosine_decay(5, 1.0, 10, min_lr=0.2) - (0.8 * 0.5 + 0.2)) < 1e-12
    assert abs(OO.cosine_decay(50, 1.0, 10, min_lr=0.2) - 0.2) < 1e-12
    assert abs(OO.cosine_decay(1, 1.0, 10, warmup_steps=4) - 0.25 * (0.5 * (1 + np.cos(np.pi * 0.1)))) < 1e-12
    assert OO.exp_decay(0, 0.1, 10, 0.5, True, begin_decay_at=3) == 0.1
    assert abs(OO.exp_decay(13, 0.1, 10, 0.5, True, begin_decay_at=3) - 0.05) < 1e-12
    assert abs(OO.exp_decay(8, 0.1, 10, 0.5, False, begin_decay_at=3) - 0.1 * 0.5 ** 0.5) < 1e-12
    assert OO.exp_decay(1000, 0.1, 10, 0.5, True, min_lr=1e-3) == 1e-3
    rng = np.random.default_rng(0)
    w = [rng.standard_normal((5, 7)).astype(np.float32), rng.standard_normal(11).astype(np.float32)]
    tw = [torch.tensor(x.copy(), requires_grad=True) for x in w]
    opt = torch.optim.Adam(tw, lr=0.01, betas=(0.9, 0.999), eps=1e-12)
    st = OO.AdamState(2)
    for _ in range(4):
        g = [rng.standard_normal(x.shape).astype(np.float32) for x in w]
        for t, gg in zip(tw, g):
            t.grad = torch.tensor(gg)
        opt.step()
        OO.adam_step(w, g, st, 0.01, epsilon=1e-12)
    for a, b in zip(w, tw):
        assert np.abs(a - b.detach().numpy()).max() < 1e-5


def test_psf_backend_shape_and_normalisation_pins(golden_dir):
    """speech_utils_test.py:45-73 for features_type='logfbank': shape (ceil8(1 + (n - win)//stride), F)
    and global mean 0 / std 1 to 6 places, on the reference's own toy wavs; plus the filterbank against
    an independent construction (torchaudio's HTK triangles differ only by psf's bin flooring, so the
    check is structural: partition of unity between the first and last centre bins)."""
    import scipy.io.wavfile as wave
    from oracle import featurizer as FZ
    wav_dir = os.path.join(golden_dir, "toy_speech_data", "wav_files")
    names = ['46gc040q.wav', '206o0103.wav', '48rc041b.wav']   # the three files the reference test uses
    for name in names:
        sr, sig = wave.read(os.path.join(wav_dir, name))
        for num_features in (64, 40):
            for stride in (10e-3, 5e-3, 40e-3):
                for win in (20e-3, 30e-3):
                    n_win, n_hop = int(sr * win), int(sr * stride)
                    length = 1 + (sig.shape[0] - n_win) // n_hop
                    if length % 8:
                        length += 8 - length % 8
                    for fn in (FZ.psf_logfbank_features, FZ.psf_spectrogram_features, FZ.psf_mfcc_features):
                        f, dur = fn(sig, sr, num_features, win, stride, pad_to=8)
                        assert f.shape == (length, num_features)
                        assert abs(np.mean(f)) < 1e-6 and abs(np.std(f) - 1.0) < 1e-6
                        assert abs(dur - len(sig) / sr) < 1e-12
                # the reference's assertion on too many spectrogram bins (speech_utils_test.py:74-85)
                with pytest.raises(AssertionError):
                    FZ.psf_spectrogram_features(sig, sr, int(sr * win) // 2 + 2, win, stride, pad_to=8)
    fb = FZ.psf_mel_filterbank(64, 512, 16000, 0.0, 8000.0)
    assert fb.shape == (64, 257) and fb.min() >= 0.0 and fb.max() <= 1.0
    peaks = fb.argmax(1)
    assert np.all(np.diff(peaks) >= 0)
    inner = fb[:, peaks[0]:peaks[-1] + 1].sum(0)
    assert np.allclose(inner, 1.0, atol=1e-12)


def test_augmentation_length_bounds_pin(golden_dir):
    """The reference's own test of augment_audio_signal (speech_utils_test.py:20-43): 100 draws with
    speed_perturbation_ratio 0.2 / 0.5 (+ noise) keep the length within [1 - r, 1 + r] x the input length."""
    import scipy.io.wavfile as wave
    from oracle import augment as AU
    sr, signal = wave.read(os.path.join(golden_dir, "toy_speech_data", "wav_files", "46gc040q.wav"))
    signal = signal.astype(np.float32)[:16000]     # 1 s of it: the bound is length-independent
    rng = np.random.RandomState(0)
    for r in (0.2, 0.5):
        aug = {"speed_perturbation_ratio": r, "noise_level_min": -90, "noise_level_max": -46}
        for _ in range(100 if r == 0.2 else 20):
            sr_new, amp = AU.draw_augmentation(len(signal), sr, aug, rng)
            n_out = AU.resample_out_len(len(signal), sr, sr_new)
            assert signal.shape[0] * (1 - r) <= n_out <= signal.shape[0] * (1 + r)
            assert 10 ** (-90 / 20.0) <= amp < 10 ** (-46 / 20.0)
        out = AU.augment_audio_signal(signal, sr, aug, rng)
        assert signal.shape[0] * (1 - r) <= out.shape[0] <= signal.shape[0] * (1 + r)
    # a list of ratios is a uniform choice among them (speech_utils.py:247-248)
    aug = {"speed_perturbation_ratio": [0.9, 1.0, 1.1]}
    seen = {AU.draw_augmentation(16000, 16000, aug, rng)[0] for _ in range(60)}
    assert seen == {14400, 16000, 17600}


def test_resampler_restatement_against_an_independent_polyphase_resampler():
    """resampy is not installed: the restated 'kaiser_best' band-limited interpolation must agree with
    scipy.signal.resample_poly (an independent Kaiser-windowed polyphase resampler) on a band-limited signal,
    and with the closed form of a resampled sinusoid."""
    import scipy.signal as ss
    from oracle import augment as AU
    sr, n = 16000, 8000
    t = np.arange(n) / sr
    x = np.sin(2 * np.pi * 440 * t) + 0.5 * np.sin(2 * np.pi * 3000 * t + 1) + 0.25 * np.sin(2 * np.pi * 6500 * t)
    for sr_new, up, down in ((17600, 11, 10), (14400, 9, 10)):
        y = AU.resample(x, sr, sr_new)
        assert len(y) == AU.resample_out_len(n, sr, sr_new) == n * up // down
        tt = np.arange(len(y)) / sr_new
        # 6.5 kHz survives both rates' Nyquist (7.2 kHz at 14.4 kHz) but sits in the filter's transition band
        exact = np.sin(2 * np.pi * 440 * tt) + 0.5 * np.sin(2 * np.pi * 3000 * tt + 1)
        ref = ss.resample_poly(x - 0.25 * np.sin(2 * np.pi * 6500 * t), up, down)
        y2 = AU.resample(x - 0.25 * np.sin(2 * np.pi * 6500 * t), sr, sr_new)
        core = slice(300, len(y2) - 300)
        assert np.abs(y2[core] - exact[core]).max() < 5e-3
        assert np.abs(y2[core] - ref[core]).max() < 5e-3


def test_spec_augment_masks_restatement():
    """speech_utils.py:419-433: n_freq_mask bands of width <= width_freq_mask over the features, n_time_mask
    bands of width <= width_time_mask over the frames, zeros written into the normalised features."""
    from oracle import augment as AU
    rng = np.random.RandomState(3)
    aug = {"n_freq_mask": 2, "n_time_mask": 2, "width_freq_mask": 6, "width_time_mask": 6}
    f = np.ones((120, 64))
    masks = AU.draw_spec_masks(120, 64, aug, rng)
    assert len(masks) == 4 and [m[0] for m in masks] == [0, 0, 1, 1]
    out = AU.apply_spec_masks(f, masks)
    for kind, base, width in masks:
        assert 0 <= width <= 6
        if kind == 0:
            assert (out[:, base:base + width] == 0).all() and base + width <= 64
        else:
            assert (out[base:base + width] == 0).all() and base + width <= 120
    assert out.sum() >= 120 * 64 - 2 * 6 * 120 - 2 * 6 * 64
    # a time band that does not fit is dropped (features.shape[0] - time_band > 0 guard)
    assert all(m[0] == 0 for m in AU.draw_spec_masks(3, 64, dict(aug, width_time_mask=50), np.random.RandomState(0))
               if m[2] >= 3)


def test_separable_conv_restatement():
    """tf.layers.separable_conv1d(depth_multiplier=1, use_bias=False, SAME) restated in torch_twin.sep_conv1d_same:
    equal to the dense convolution with W[k,c,o] = D[k,c] P[c,o] (the definition of a separable kernel) for the
    stride / dilation combinations of the QuartzNet config, and to an explicit per-channel loop."""
    torch.manual_seed(0)
    x = torch.randn(2, 37, 8, dtype=torch.float64)
    D = torch.randn(5, 8, 1, dtype=torch.float64)
    P = torch.randn(1, 8, 6, dtype=torch.float64)
    for stride, dil in ((1, 1), (2, 1), (1, 2)):
        a = TT.sep_conv1d_same(x, D, P, stride, dil)
        b = TT.conv1d_same(x, D * P, stride, dil)
        assert float((a - b).abs().max()) < 1e-12
    # explicit loops, stride 1: z[t,c] = sum_k D[k,c] x[t - pad + k, c]
    K, pad = 5, 2
    z = torch.zeros(2, 37, 8, dtype=torch.float64)
    for t in range(37):
        for k in range(K):
            s = t - pad + k
            if 0 <= s < 37:
                z[:, t] += D[k, :, 0] * x[:, s]
    assert float((z @ P[0] - TT.sep_conv1d_same(x, D, P, 1, 1)).abs().max()) < 1e-12


def test_augmentation_draws_reproduce_the_executed_reference_fixture(golden_dir):
    """tests/golden/reference_augmentation_draws.json holds what the reference's own augment_audio_signal did for
    seeded np.random streams (output length, drawn noise level; written by tools/make_golden_reference_draws.py in
    the build container).  The oracle's draw function and, through it, the GPU data layer's host-side draws have to
    reproduce them from the same seeds."""
    import json
    from oracle import augment as AU
    fx = json.load(open(os.path.join(golden_dir, "reference_augmentation_draws.json")))
    n, sr = fx["n_samples"], fx["sample_freq"]
    assert len(fx["cases"]) >= 32
    for case in fx["cases"]:
        sr_new, amp = AU.draw_augmentation(n, sr, case["augmentation"], np.random.RandomState(case["seed"]))
        n_out = AU.resample_out_len(n, sr, sr_new) if sr_new > 0 else n
        assert n_out == case["n_out"], case
        if case["noise_level_db"] is None:
            assert amp == 0
        else:
            assert abs(amp - 10.0 ** (case["noise_level_db"] / 20.0)) < 1e-12, case
