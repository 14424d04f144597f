"""One Jasper 10x5 training step bracketed by cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:tapgemm_kmajor -s 70 -c 2 -o gpurun_out/conv_full python tools/profile_step.py
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openseq2seq_b200.compat as compat  # noqa: E402

compat.install()
from open_seq2seq.utils.utils import get_base_config, nested_update  # noqa: E402
import bench  # noqa: E402

B = int(os.environ.get("PROFILE_BATCH", "32"))
_, cfg, model_cls, module = get_base_config(["--config_file=" + os.path.join(ROOT, "configs", "jasper10x5_dr.py")])
cfg = copy.deepcopy(cfg)
nested_update(cfg, copy.deepcopy(module["train_params"]))
cfg.pop("num_epochs", None)
cfg["max_steps"] = 1000
cfg["batch_size_per_gpu"] = B
model = model_cls(params=cfg, mode="train", hvd=None)
model.compile()
dl = model.get_data_layer()
waves = bench.synth_waveforms(0, B, 15.0)
y, ylen = bench.synth_labels(0, B)
yd, yl = torch.tensor(y).cuda(), torch.tensor(ylen).cuda()


def step():
    feats, flens = dl.featurize(waves, seed=model.engine.step_count)
    model.train_step({"source_tensors": [feats, flens], "target_tensors": [yd, yl]})


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step; loss", float(model.loss))
