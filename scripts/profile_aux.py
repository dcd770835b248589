"""One warm-up + one measured pass of the ISW covariance (C=256) and an adaptive density map (2048^2, 25k heads);
run under ncu with -k regex:'isw_gram_tc|isw_cov_finish|dmap_' to capture the secondary kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dgvcc_b200 import _native, synthetic
from dgvcc_b200.utils import dmap_gen

dev = torch.device("cuda:0")
lib = _native.lib()
b, c, hw = 8, 256, 6400
x = torch.randn(b, c, hw, device=dev)
eye = torch.eye(c, device=dev)
n = lib.dgvcc_isw_workspace_bytes(b, c, hw)
ws = torch.empty(n, dtype=torch.uint8, device=dev)
fc = torch.empty(b, c, c, device=dev)
pts = synthetic.crowd_points(np.random.default_rng(4004), 25000, 2048, 2048, dtype=np.float64)
for _ in range(2):
    lib.dgvcc_isw_covariance(_native.ptr(x), _native.ptr(eye), b, c, hw, 1, _native.ptr(ws), n, _native.ptr(fc), _native.stream_ptr(dev))
    d = dmap_gen.gaussian_filter_density(np.empty((2048, 2048, 0)), pts)
torch.cuda.synchronize()
print("ok", float(fc[0, 0, 0]), float(d.sum()))
