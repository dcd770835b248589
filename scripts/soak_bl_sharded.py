"""Randomised soak of the point-chunk sharded Bayesian loss: random batches, grids, strides, sigmas, background on/off,
2..4 ranks inside this process on one GPU (LocalComm), random density owners, random chunk sizes, culling on/off, deferred
or immediate loss.  Every rank's loss and the gathered gradient must be BIT-IDENTICAL to one GPU running the same chunk
table (the exchange must not change a bit); additionally within rtol 1e-5 of the ordinary BL module.

    python scripts/soak_bl_sharded.py [seconds]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from dgvcc_b200 import synthetic
from dgvcc_b200.losses import bl as blmod
from test_bl_sharded_gpu import run_sharded, single_gpu_with_table

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(88002)
t0, cases = time.time(), 0
while time.time() - t0 < budget:
    stride = int(rng.choice([4, 8, 16]))
    hp, wp = int(rng.integers(2, 70)), int(rng.integers(2, 100))
    h, w = hp * stride, wp * stride
    b = int(rng.integers(1, 7))
    counts = [int(rng.choice([0, 1, 2, 5, 31, 33, 100, 257, 700, 1500])) for _ in range(b)]
    sigma = float(rng.choice([4.0, 8.0, 5.5, 12.0]))
    use_bg = bool(rng.integers(0, 2))
    bg_ratio = float(rng.choice([1.0, 0.15, 0.5]))
    world = int(rng.integers(2, 5))
    cull = bool(rng.integers(0, 2))
    blmod._CHUNK_POINTS = int(rng.choice([17, 64, 256, 1024]))
    owners = [int(rng.integers(0, world)) for _ in range(b)]
    cfg = int(rng.integers(100, 10000))
    pts, tgt, dens, st = synthetic.bl_batch(cfg, counts, w, h, stride)
    pts = [torch.from_numpy(p) for p in pts]
    tgt = [torch.from_numpy(t) for t in tgt]
    dens, st = torch.from_numpy(dens), torch.from_numpy(st)
    losses, grad, plan = run_sharded(world, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull, owners, steps=2,
                                     defer=bool(rng.integers(0, 2)))
    ref_loss, ref_grad = single_gpu_with_table(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull)
    what = f"cfg={cfg} counts={counts} grid={hp}x{wp} stride={stride} sigma={sigma} bg={use_bg}/{bg_ratio} world={world} owners={owners} cull={cull} chunk={blmod._CHUNK_POINTS}"
    for r, l in enumerate(losses):
        if not torch.equal(l, ref_loss):
            raise SystemExit(f"MISMATCH loss on rank {r}: {float(l)!r} vs {float(ref_loss)!r}: {what}")
    if not torch.equal(grad, ref_grad):
        raise SystemExit(f"MISMATCH gradient in {int((grad != ref_grad).sum())} pixels: {what}")
    cases += 1
print(f"soak ok: {cases} random sharded batches bit-identical to one GPU, {time.time() - t0:.0f} s")
