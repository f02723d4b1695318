"""A few launches of the batched GPU get_spec on the BASELINE config-4 corpus (for ncu)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

win = importlib.import_module("autoencoded-vocal-analysis_b200.models.window_vae_dataset")
rng = np.random.default_rng(0)
fs = bench.FINCH_P['fs']
audio = [(3000 * rng.standard_normal(int(60.0 * fs), dtype=np.float32)).astype(np.int16) for _ in range(4)]
rois = [np.array([[1.0, 25.0], [30.0, 58.0]]) for _ in range(4)]
ds = win.FixedWindowDataset(["f%d.wav" % k for k in range(4)], None, dict(bench.FINCH_P), audio=audio, fs=fs, rois=rois)
for _ in range(4):
    x = ds.sample_batch(1024)
torch.cuda.synchronize()
print("ok", tuple(x.shape))
