"""configs/jasper10x5_dr.py with a FILE-backed training set: dataset_files = [$OS2S_DATASET_CSV]
(tools/make_wav_dataset.py writes one).  Everything else -- model, optimizer, augmentation -- is the
headline configuration."""
import copy
import os
import runpy

_m = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "jasper10x5_dr.py"))
base_model = _m["base_model"]
base_params = copy.deepcopy(_m["base_params"])
train_params = copy.deepcopy(_m["train_params"])
eval_params = copy.deepcopy(_m["eval_params"])
infer_params = copy.deepcopy(_m["infer_params"])
_csv = os.environ.get("OS2S_DATASET_CSV")
if _csv:
    train_params["data_layer_params"]["dataset_files"] = [_csv]
