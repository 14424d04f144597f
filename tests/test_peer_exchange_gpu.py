"""Peer-memory gradient exchange (csrc/peer.cu) against NCCL's all-reduce; needs two GPUs on the box (the
single-GPU `-m gpu` run skips it; tools/r2_peer_n2.sh runs the same check under `gpurun --gpus 2 / 4`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchange_matches_nccl_and_is_bit_identical_on_all_ranks():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "peer_check.py"), "--floats", "20000000", "--iters", "4"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"] and out["identical_bits_on_all_ranks"] and out["max_rel_err_vs_nccl"] < 1e-6
