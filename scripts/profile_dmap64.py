"""Adaptive density maps for the 64 JHU-shaped images of BASELINE config 4 (device-resident loop), for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench_aux import config4_images
from dgvcc_b200.utils import dmap_gen
dev = torch.device("cuda:0")
tot = 0.0
for (h, w), pts in config4_images():
    p = np.ascontiguousarray(pts, dtype=np.float64)
    tot += float(dmap_gen._density_device(h, w, p, True, dev).sum()) if len(p) else 0.0
torch.cuda.synchronize()
print("ok", tot)
