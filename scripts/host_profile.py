"""cProfile of the end-to-end BL step (host lists in, loss + gradient out) to see where the host time goes."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200 import synthetic
from dgvcc_b200.losses.bl import BL

dev = torch.device("cuda:0")
counts = synthetic.config_counts(3)
w, h = synthetic.CONFIG_SHAPES[3]
pts, tgt, dens, st = synthetic.bl_batch(3, counts, w, h)
pts_h = [torch.from_numpy(p).pin_memory() for p in pts]
tgt_h = [torch.from_numpy(t).pin_memory() for t in tgt]
dens_h = torch.from_numpy(dens).pin_memory()
st_h = torch.from_numpy(st).pin_memory()
grad_h = torch.empty_like(dens_h).pin_memory()
mod = BL(8.0, 2048, 8, 1.0, True, dev)

def step():
    d = dens_h.to(dev, non_blocking=True).requires_grad_(True)
    loss = mod(pts_h, st_h.to(dev, non_blocking=True), tgt_h, d)
    loss.backward()
    grad_h.copy_(d.grad, non_blocking=True)
    return float(loss.detach())

for _ in range(5):
    step()
# host time until everything is enqueued vs total
t0 = time.perf_counter(); n = 50
for _ in range(n):
    step()
print(f"e2e step {1e3 * (time.perf_counter() - t0) / n:.3f} ms")
def enqueue_only():
    d = dens_h.to(dev, non_blocking=True).requires_grad_(True)
    loss = mod(pts_h, st_h.to(dev, non_blocking=True), tgt_h, d)
    loss.backward()
    grad_h.copy_(d.grad, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n):
    enqueue_only()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host enqueue time per step {1e3 * t_host / n:.3f} ms (GPU path ~2.0 ms)")
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
