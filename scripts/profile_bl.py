"""Two fused-BL steps on BASELINE config 3 (first = warm-up), dense sweep (what bench.py grades); run under ncu with
-k regex:bl_ -s 7 -c 7 (7 launches per step: grid build, minima, z, counts, row reduction, selection, grad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200 import synthetic
from dgvcc_b200.losses.bl import BL

dev = torch.device("cuda:0")
cfg = int(os.environ.get("CFG", "3"))
counts = synthetic.config_counts(cfg)
w, h = synthetic.CONFIG_SHAPES[cfg]
pts, tgt, dens, st = synthetic.bl_batch(cfg, counts, w, h)
pts = [torch.from_numpy(p).to(dev) for p in pts]
tgt = [torch.from_numpy(t).to(dev) for t in tgt]
dens = torch.from_numpy(dens).to(dev).requires_grad_(True)
st = torch.from_numpy(st).to(dev)
mod = BL(8.0, max(w, h), 8, 1.0, True, dev)
mod.exact_cull = False
for _ in range(2):
    dens.grad = None
    loss = mod(pts, st, tgt, dens)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
