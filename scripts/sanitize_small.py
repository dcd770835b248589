"""Small end-to-end calls of every kernel family in one process: a quick "does everything launch and finish" check,
and the workload to hand to compute-sanitizer (memcheck / racecheck) where that tool is available -- it is closed
on this GPU pool.  Its substitute is the checked build: `DGVCC_BOUNDS_CHECK=1 python scripts/sanitize_small.py` runs the
same calls through libdgvcc_b200_chk.so, whose kernels assert every table-derived / data-dependent index
(csrc/common.cuh); the whole GPU suite runs that way with `DGVCC_BOUNDS_CHECK=1 python -m pytest tests -m gpu`.
Sizes are tiny on purpose; parity is the job of tests/, not of this script."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dgvcc_b200 import synthetic
from dgvcc_b200.losses.bl import BL
from dgvcc_b200.losses.lw import lw_loss
from dgvcc_b200.losses.ortho import ortho_loss
from dgvcc_b200.models.ISW import CovMatrix_ISW, InstanceWhitening, instance_whitening_loss, variance_of_covariance
from dgvcc_b200.utils import dmap_gen
from dgvcc_b200.datasets import bay_targets, den_targets

dev = torch.device("cuda:0")
from dgvcc_b200 import _native
print(f"library: {_native.lib()._name} (device-side index checks: {'ON' if _native.lib().dgvcc_bounds_checked() else 'off'})")
rng = np.random.default_rng(1)
# Bayesian loss: ragged batch incl. an empty image and a multi-chunk one
counts = [40, 0, 3, 1500]
pts = [torch.from_numpy(synthetic.crowd_points(rng, n, 256, 192, dtype=np.float32)).to(dev) for n in counts]
tgt = [torch.rand(n, device=dev) for n in counts]
dens = torch.rand(4, 1, 24, 32, device=dev, requires_grad=True)
st = torch.full((4,), 192.0, device=dev)
BL(8.0, 256, 8, 1.0, True, dev)(pts, st, tgt, dens).backward()
# density maps: ragged list, adaptive and fixed
shapes = [(100, 140), (64, 64), (257, 300)]
plist = [synthetic.crowd_points(rng, n, s[1], s[0], dtype=np.float64) for n, s in zip((60, 0, 2300), shapes)]
for fixed in (False, True):
    maps = dmap_gen.gaussian_filter_density_batch(shapes, plist, fixed=fixed)
full = [torch.from_numpy(m).to(dev) for m in maps]
den_targets.train_density_targets([full[0], full[2]], [(0, 0, 10, 20, 1), (0, 0, 100, 30, 0)], (64, 64), 2)
d = bay_targets.cal_dists(plist[0])
bay_targets.crop_targets(plist[0], d, 10, 20, 64, 64)
# ISW: tensor-core shapes and a SIMT fallback shape
for (b, c, h, w) in [(2, 64, 16, 16), (2, 128, 12, 12), (1, 40, 7, 5)]:
    x = torch.randn(b, c, h, w, device=dev, requires_grad=True)
    cm = CovMatrix_ISW(dim=c, relax_denom=2.0)
    eye, rev = cm.get_eye_matrix()
    _, wt = InstanceWhitening(c)(x)
    cm.set_variance_of_covariance(variance_of_covariance(wt.detach()[:2] if b >= 2 else wt.detach().repeat(2, 1, 1, 1), eye, rev))
    _, mask, margin, num = cm.get_mask_matrix()
    instance_whitening_loss(wt, eye, mask, margin, num).backward()
    lw_loss(x, torch.rand(b, 1, h, w, device=dev)).backward()
a, bb = torch.randn(64, 256, device=dev, requires_grad=True), torch.randn(64, 256, device=dev, requires_grad=True)
ortho_loss(a, bb).backward()
torch.cuda.synchronize()
print("ok")
