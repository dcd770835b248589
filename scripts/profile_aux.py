"""One warm-up + one measured pass of the secondary kernels: the ISW module (InstanceWhitening + covariance loss,
forward + backward, B=8 C=256 HW=6400) and one density map (2048^2, 25 000 heads), adaptive then fixed sigma.
Run under ncu with -k regex:'isw_|dmap_' -s 23 -c 23 (23 matching launches per pass: 7 ISW, 9 adaptive, 7 fixed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dgvcc_b200 import synthetic
from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss
from dgvcc_b200.utils import dmap_gen

dev = torch.device("cuda:0")
b, c, h, w = 8, 256, 80, 80
xin = torch.randn(b, c, h, w, device=dev, requires_grad=True)
eye = torch.eye(c, device=dev)
mask = torch.triu(torch.ones(c, c, device=dev), 1)
num = mask.sum()
margin = torch.zeros((), device=dev)
iw = InstanceWhitening(c)
pts = synthetic.crowd_points(np.random.default_rng(4004), 25000, 2048, 2048, dtype=np.float64)
for _ in range(2):
    xin.grad = None
    y, wt = iw(xin)
    loss = instance_whitening_loss(wt, eye, mask, margin, num)
    loss.backward()
    da = dmap_gen._density_device(2048, 2048, pts, True, dev)
    df = dmap_gen._density_device(2048, 2048, pts, False, dev)
torch.cuda.synchronize()
print("ok", float(loss), float(da.sum()), float(df.sum()))
