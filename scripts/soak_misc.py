"""Randomised soak of the small next-row kernels against their oracles: dataset-side density targets (pad, crop, sum-pool,
flip, occupancy), the CovMatrix_ISW top-k mask, and the Bayesian-dataset crop targets.

    python scripts/soak_misc.py [seconds]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import den_targets_oracle as do, bay_targets_oracle as bo
from oracle.cov_settings_oracle import CovMatrixISW
from dgvcc_b200.datasets import den_targets, bay_targets
from dgvcc_b200.models.ISW.cov_settings import topk_mask

oracle.warm_up()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(4242)
t0, n_den, n_mask, n_bay = time.time(), 0, 0, 0
while time.time() - t0 < budget:
    # ---- density targets
    d = int(rng.choice([1, 2, 4, 8]))
    ch, cw = 16 * d * int(rng.integers(1, 5)), 16 * d * int(rng.integers(1, 5))
    maps, geoms = [], []
    for k in range(int(rng.integers(1, 5))):
        h, w = int(rng.integers(8, 400)), int(rng.integers(8, 400))
        m = np.where(rng.random((h, w)) < 0.03, rng.random((h, w)), 0).astype(np.float32)
        ph, pw = max(h, ch), max(w, cw)
        top, left = (ph - h) // 2, (pw - w) // 2
        geoms.append((left, top, int(rng.integers(0, ph - ch + 1)), int(rng.integers(0, pw - cw + 1)), int(rng.integers(0, 2))))
        maps.append(torch.from_numpy(m).cuda() if rng.random() < 0.5 else m)
    dm, bm = den_targets.train_density_targets(maps, geoms, (ch, cw), d)
    for k, (m, g) in enumerate(zip(maps, geoms)):
        mm = m.cpu().numpy() if isinstance(m, torch.Tensor) else m
        ref = do.train_density(mm, g[0], g[1], g[2], g[3], ch, cw, d, g[4])
        np.testing.assert_allclose(dm[k].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=0)
        assert np.array_equal(bm[k].cpu().numpy(), do.block_occupancy(ref)[0].numpy()), "occupancy map"
        n_den += 1
    # ---- top-k mask (no ties among the positive entries: continuous values)
    c = int(rng.choice([8, 16, 40, 64, 128]))
    stats = [(torch.rand(c, c) ** 3).triu(1) for _ in range(int(rng.integers(1, 5)))]
    cm = CovMatrixISW(c, float(rng.choice([2.0, 3.0, 1.5])))
    for s_ in stats:
        cm.set_variance_of_covariance(s_)
    cm.set_mask_matrix()
    k = int(cm.num_off_diagonal - cm.margin)
    _, mask = topk_mask(torch.stack([s_.reshape(-1) for s_ in stats]).cuda(), len(stats), k)
    assert torch.equal(mask.cpu().view(c, c), cm.mask_matrix), "top-k mask"
    n_mask += 1
    # ---- crop targets (kept set exact)
    n = int(rng.choice([0, 1, 3, 50, 400]))
    gt = rng.uniform(0, 600, size=(n, 2))
    dists = rng.uniform(1, 200, size=(n, 1))
    i, j, hh, ww = int(rng.integers(0, 300)), int(rng.integers(0, 300)), 256, 256
    g_ref, t_ref = bo.crop_targets(gt.copy(), dists, i, j, hh, ww)
    g_got, t_got = bay_targets.crop_targets(gt.copy(), dists, i, j, hh, ww)
    assert len(t_got) == len(t_ref)
    if len(t_ref):
        np.testing.assert_array_equal(g_got, g_ref)
        np.testing.assert_allclose(t_got, t_ref, rtol=1e-12)
    n_bay += 1
print(f"soak ok: {n_den} density-target maps, {n_mask} masks, {n_bay} crop-target sets, {time.time() - t0:.0f} s")
