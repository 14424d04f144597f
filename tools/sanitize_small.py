"""One small launch of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

  compute-sanitizer --tool memcheck  --log-file profiles/r02_sanitizer_memcheck.log  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck --log-file profiles/r02_sanitizer_racecheck.log python tools/sanitize_small.py
  compute-sanitizer --tool synccheck --log-file profiles/r02_sanitizer_synccheck.log python tools/sanitize_small.py

Shapes are small (the tools slow kernels down 10-100x) but reach every code path of the hand-rolled
mbarrier / cluster protocols: single-CTA tiles, CTA pairs (cta_group::2), pairs + halo tile, the pair weight
gradient with an odd block count, the fused BN-statistics / BN-backward epilogues, both 16-bit formats."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseq2seq_b200.engine import JasperEngine  # noqa: E402

LAYERS = [
    {"type": "conv1d", "repeat": 1, "kernel_size": [11], "stride": [2], "num_channels": 256, "padding": "SAME",
     "dilation": [1], "dropout_keep_prob": 0.9},
    {"type": "conv1d", "repeat": 2, "kernel_size": [5], "stride": [1], "num_channels": 384, "padding": "SAME",
     "dilation": [1], "dropout_keep_prob": 0.9, "residual": True, "residual_dense": True},
    {"type": "sep_conv1d", "repeat": 2, "kernel_size": [13], "stride": [1], "num_channels": 384, "padding": "SAME",
     "dilation": [1], "dropout_keep_prob": 0.9, "residual": True, "residual_dense": False},
    {"type": "conv1d", "repeat": 1, "kernel_size": [3], "stride": [1], "num_channels": 256, "padding": "SAME",
     "dilation": [2], "dropout_keep_prob": 0.9},
]


def main():
    torch.manual_seed(0)
    for half, conv in (("bf16", "fp16"), ("fp16", "fp32")):
        eng = JasperEngine(LAYERS, 64, 29, training=True, act_dtype=half, conv_dtype=conv,
                           opt=dict(algo="novograd", weight_decay=0.001, larc_eta=0.001, learning_rate=0.02,
                                    decay_steps=100, power=2.0, loss_scaling=True))
        eng.use_cuda_graph = False
        B, T = 2, 288
        lens = torch.tensor([288, 170], dtype=torch.int32).cuda()
        x = torch.randn(B, T, 64).cuda()
        labels = torch.randint(0, 28, (B, 20), dtype=torch.int32).cuda()
        ll = torch.tensor([20, 11], dtype=torch.int32).cuda()
        for _ in range(2):
            loss = eng.train_step(x, lens, labels, ll)
        eng.greedy_decode()
        torch.cuda.synchronize()
        print("sanitize_small: %s/%s loss %.3f step %d" % (half, conv, float(loss.mean()), int(eng.istate[2])))


if __name__ == "__main__":
    main()
