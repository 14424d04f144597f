"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: the data-parallel gradient exchange is a
SUM all-reduce over one flat fp32 buffer whose 1/N lives in the optimizer's unscale factor, which
must equal the reference's Horovod mean (optimizers/optimizers.py:77-104); rank-0 broadcast of the
state; scalar gather; per-rank data seeds; contiguous eval sharding (speech2text.py:200-210)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from openseq2seq_b200.dist import TorchDistHvd
    from oracle import optimizer as OO
    hvd = TorchDistHvd.init(backend="gloo")
    assert hvd.size() == world and hvd.rank() == rank
    rng = np.random.default_rng(100 + rank)
    sizes = [7, 130, 33]
    offs = np.cumsum([0] + sizes)
    flat = torch.tensor(rng.standard_normal(offs[-1]).astype(np.float32))
    mine = flat.clone()
    hvd.allreduce_(flat)                                   # SUM over ranks, in place, one buffer
    # every rank rebuilds both ranks' gradients and checks sum == what the all-reduce produced
    both = [np.random.default_rng(100 + r).standard_normal(offs[-1]).astype(np.float32) for r in range(world)]
    assert np.allclose(flat.numpy(), both[0] + both[1], atol=1e-6)
    # the reference semantics: mean over ranks (hvd.allreduce average=True), then the optimizer.
    w_sum = [np.ones(n, dtype=np.float32) * (i + 1) for i, n in enumerate(sizes)]
    w_ref = [w.copy() for w in w_sum]
    st1, st2 = OO.NovoGradState(3), OO.NovoGradState(3)
    sc1, sc2 = OO.BackoffScaler(), OO.BackoffScaler()
    opt = dict(beta1=0.95, beta2=0.98, epsilon=1e-8, weight_decay=0.001)
    # (a) oracle fed per-rank gradients (it averages them)
    per_rank = [[b[offs[i]:offs[i + 1]] * sc1.scale for i in range(3)] for b in both]
    OO.train_step(w_ref, per_rank, st1, sc1, 0, lambda s: 0.01, opt, larc_params=dict(larc_eta=0.001))
    # (b) the summed buffer + unscale by 1/(scale * world): what os2s_opt_step computes
    summed = [(flat.numpy()[offs[i]:offs[i + 1]] * sc2.scale / world) for i in range(3)]
    OO.train_step(w_sum, [summed], st2, sc2, 0, lambda s: 0.01, opt, larc_params=dict(larc_eta=0.001))
    for a, b in zip(w_ref, w_sum):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-7)
    assert abs(hvd.sum_scalar(rank + 1.5) - (1.5 + 2.5)) < 1e-9
    t = torch.full((4,), float(rank))
    torch.distributed.broadcast(t, src=0)
    assert torch.all(t == 0)
    hvd.barrier()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("1")


def test_two_rank_gradient_sum_equals_horovod_mean_and_collectives(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_data_layer_sharding_and_seeding_without_gpu():
    import openseq2seq_b200.compat as compat
    compat.install()
    import tensorflow as tf
    from open_seq2seq.data.speech2text.speech2text import Speech2TextDataLayer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    class _M(object):
        params = {"random_seed": 3}

    base = {"num_audio_features": 64, "input_type": "logfbank", "vocab_file": os.path.join(root, "configs", "vocab.txt"),
            "dataset_files": ["synthetic:10:1.0:5"], "backend": "librosa", "norm_per_feature": True, "pad_to": 16,
            "batch_size": 2, "dtype": tf.float32}
    ev = [Speech2TextDataLayer(dict(base, mode="eval", shuffle=False), _M(), 2, r) for r in range(2)]
    assert ev[0].get_size_in_samples() == 5 and ev[1].get_size_in_samples() == 5
    assert [f[0] for f in ev[0]._files] + [f[0] for f in ev[1]._files] == ["synthetic:%d:1:5" % i for i in range(10)]
    tr = [Speech2TextDataLayer(dict(base, mode="train"), _M(), 2, r) for r in range(2)]
    assert tr[0].get_size_in_samples() == tr[1].get_size_in_samples() == 10   # train: no sharding
    assert tr[0].params["tgt_vocab_size"] == 29 and tr[0].params["char2idx"]["a"] == 1
    with pytest.raises(ValueError, match="Shuffle should not be performed"):
        Speech2TextDataLayer(dict(base, mode="eval", shuffle=True), _M(), 1, 0)
    with pytest.raises(ValueError, match="Unknown parameter"):
        Speech2TextDataLayer(dict(base, mode="train", bogus=1), _M(), 1, 0)
