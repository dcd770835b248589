"""Randomised soak of the batched density-map path against the closed-form oracle (bit-exact maps expected) and of the
batched kNN against scipy's KDTree: many small ragged batches with random shapes, head counts and crowd structure.

    python scripts/soak_dmap.py [seconds]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dgvcc_b200 import synthetic
from dgvcc_b200.utils import dmap_gen
from oracle import dmap_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(20261018)
t0 = time.time()
batches = images = mism = 0
while time.time() - t0 < budget:
    b = int(rng.integers(1, 7))
    shapes, plist = [], []
    for _ in range(b):
        h, w = int(rng.integers(17, 420)), int(rng.integers(17, 420))
        kind = rng.integers(0, 5)
        n = [0, int(rng.integers(1, 5)), int(rng.integers(5, 60)), int(rng.integers(60, 600)), int(rng.integers(600, 4500))][kind]
        if rng.random() < 0.5:
            pts = synthetic.crowd_points(np.random.default_rng(int(rng.integers(1 << 30))), n, w, h, dtype=np.float64)
        else:  # uniform, some heads outside the image (skipped for the splat, still neighbours)
            pts = rng.uniform([-3, -3], [w + 5, h + 5], size=(n, 2))
            pts = np.abs(pts)
        if rng.random() < 0.3:
            pts = pts.astype(np.float32).astype(np.float64)
        if n and len(np.unique(pts, axis=0)) != n:
            continue  # KDTree's tie order is implementation-defined
        shapes.append((h, w)); plist.append(pts)
    if not shapes:
        continue
    for fixed in (False, True):
        outs = dmap_gen.gaussian_filter_density_batch(shapes, plist, fixed=fixed)
        for (h, w), p, o in zip(shapes, plist, outs):
            if len(p) == 0:
                ok = not o.any()
            else:
                ref = dmap_oracle.density_closed_form((h, w), p, fixed=fixed)
                ok = np.array_equal(o, ref)
                if not ok:  # the documented gate: rtol 1e-5 (an fp64 exp may differ in its last ulp)
                    tol = 1e-5 * np.abs(ref) + 1e-7 * float(np.abs(ref).max())
                    if (np.abs(o.astype(np.float64) - ref) > tol).any():
                        raise SystemExit(f"MISMATCH beyond tolerance: {h}x{w} n={len(p)} fixed={fixed}")
                    mism += 1
            assert ok or mism, "empty map not zero"
            images += 1
    batches += 1
print(f"soak ok: {batches} batches, {images} maps, {mism} not bit-identical (within rtol 1e-5), {time.time() - t0:.0f} s")
