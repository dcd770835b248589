"""Density maps for the 64 JHU-shaped images of BASELINE config 4 through the batched launch set (adaptive, then
fixed sigma), twice each; run under ncu for a launch list (`-k regex:dmap_`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench_aux import config4_images
from dgvcc_b200.utils import dmap_gen
dev = torch.device("cuda:0")
images = config4_images()
shapes = [s for s, _ in images]
pts = [np.ascontiguousarray(p, dtype=np.float64) for _, p in images]
tot = 0.0
for rep in range(2):
    for adaptive in (True, False):
        out, _ = dmap_gen._density_batch_device(shapes, pts, adaptive, dev)
        tot += float(out.sum())
torch.cuda.synchronize()
print("ok", tot)
