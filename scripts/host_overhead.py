"""Where does the host time of one end-to-end BL step go? (perf_counter around phases, with syncs)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200 import synthetic
from dgvcc_b200.losses import bl as blmod

dev = torch.device("cuda:0")
counts = synthetic.config_counts(3)
w, h = synthetic.CONFIG_SHAPES[3]
pts, tgt, dens, st = synthetic.bl_batch(3, counts, w, h)
pts_h = [torch.from_numpy(p).pin_memory() for p in pts]
tgt_h = [torch.from_numpy(t).pin_memory() for t in tgt]
dens_h = torch.from_numpy(dens).pin_memory()
st_h = torch.from_numpy(st).pin_memory()
grad_h = torch.empty_like(dens_h).pin_memory()
mod = blmod.BL(8.0, 2048, 8, 1.0, True, dev)
sync = torch.cuda.synchronize

def phase(name, fn, acc):
    sync(); t0 = time.perf_counter(); r = fn(); sync(); acc[name] = acc.get(name, 0) + time.perf_counter() - t0; return r

for it in range(6):
    acc = {}
    d = phase("h2d density", lambda: dens_h.to(dev, non_blocking=True).requires_grad_(True), acc)
    s = phase("h2d st", lambda: st_h.to(dev, non_blocking=True), acc)
    packed = phase("pack points (host cat + pin + h2d + meta)", lambda: blmod._Packed(pts_h, True, dev), acc)
    targets = phase("pack targets", lambda: blmod._pack_targets(tgt_h, packed, dev), acc)
    t0 = time.perf_counter()
    loss = blmod._FusedBL.apply(d, packed, targets, s, 8.0, 8.0, 1.0, True, 1.0 / 16, None, False)
    acc["forward launch (host, async)"] = time.perf_counter() - t0
    sync(); acc["forward total"] = time.perf_counter() - t0
    t0 = time.perf_counter(); loss.backward(); acc["backward launch (host, async)"] = time.perf_counter() - t0
    sync(); acc["backward total"] = time.perf_counter() - t0
    phase("d2h grad + loss", lambda: (grad_h.copy_(d.grad, non_blocking=True), float(loss.detach())), acc)
    if it >= 4:
        print({k: round(v * 1e3, 3) for k, v in acc.items()})
