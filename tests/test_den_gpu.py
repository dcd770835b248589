"""Parity of the CUDA density-target kernel (through the C ABI) with the reference fixtures and the oracle.

Gates: pooled maps rtol 1e-5 (the d x d fp32 sums run in a different order than torch's CPU reduction;
the inputs are non-negative, so no cancellation) -- and bit-exact when d == 1 (a pure copy); occupancy maps
bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import den_targets_oracle as do
from helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "den_cases.npz"))


def geom(fixtures, k):
    left, top, i, j, h, w, down, flip = (int(v) for v in fixtures[f"den_{k}_geom"])
    return (left, top, i, j, flip), (h, w), down


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 5])
def test_matches_reference_fixture(fixtures, k):
    from dgvcc_b200.datasets import den_targets
    g, crop, down = geom(fixtures, k)
    d, b = den_targets.train_density_targets([fixtures[f"den_{k}_dmap"]], [g], crop, down)
    ref_d, ref_b = fixtures[f"den_{k}_ref_dmap"], fixtures[f"den_{k}_ref_bmap"]
    assert d.shape == (1,) + ref_d.shape and b.shape == ref_b.shape
    np.testing.assert_allclose(d[0].cpu().numpy(), ref_d, rtol=1e-5, atol=0)
    if down == 1:
        assert np.array_equal(d[0].cpu().numpy(), ref_d)
    assert np.array_equal(b.cpu().numpy(), ref_b)


def test_ragged_batch_against_oracle():
    """A batch of maps of different sizes (device tensors and numpy mixed), padding on both axes, flips, and the
    den_dataset.py variant without an occupancy map (pooled size not a multiple of 16)."""
    from dgvcc_b200.datasets import den_targets
    rng = np.random.default_rng(31)
    crop, down = (320, 384), 4
    maps, geoms = [], []
    for k, (h, w) in enumerate([(700, 900), (200, 500), (400, 300), (320, 384), (100, 100)]):
        m = np.where(rng.random((h, w)) < 0.02, rng.random((h, w)), 0).astype(np.float32)
        ph, pw = max(h, crop[0]), max(w, crop[1])
        top, left = (ph - h) // 2, (pw - w) // 2
        i, j = int(rng.integers(0, ph - crop[0] + 1)), int(rng.integers(0, pw - crop[1] + 1))
        maps.append(torch.from_numpy(m).cuda() if k % 2 else m)
        geoms.append((left, top, i, j, k % 2))
    d, b = den_targets.train_density_targets(maps, geoms, crop, down)
    assert d.shape == (5, 1, 80, 96) and b.shape == (5, 5, 6)
    for k, (m, g) in enumerate(zip(maps, geoms)):
        mm = m.cpu().numpy() if isinstance(m, torch.Tensor) else m
        ref = do.train_density(mm, g[0], g[1], g[2], g[3], crop[0], crop[1], down, g[4])
        np.testing.assert_allclose(d[k].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=0)
        assert np.array_equal(b[k].cpu().numpy(), do.block_occupancy(ref)[0].numpy())
        # mass inside the crop is conserved by the pooling
        assert abs(float(d[k].sum()) - float(ref.sum())) <= 1e-4 * max(1.0, float(ref.sum()))
    d2, b2 = den_targets.train_density_targets(maps, geoms, (300, 260), 20, with_bmap=False)
    assert b2 is None and d2.shape == (5, 1, 15, 13)
    for k, (m, g) in enumerate(zip(maps, geoms)):
        mm = m.cpu().numpy() if isinstance(m, torch.Tensor) else m
        gi, gj = min(g[2], max(mm.shape[0], 300) - 300), min(g[3], max(mm.shape[1], 260) - 260)
        top, left = max(0, (300 - mm.shape[0]) // 2), max(0, (260 - mm.shape[1]) // 2)
        d3, _ = den_targets.train_density_targets([m], [(left, top, gi, gj, g[4])], (300, 260), 20, with_bmap=False)
        ref = do.train_density(mm, left, top, gi, gj, 300, 260, 20, g[4])
        np.testing.assert_allclose(d3[0].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=0)


def test_straight_from_the_generator():
    """Full-resolution maps produced on the device by dmap_gen feed the target kernel without leaving HBM."""
    from dgvcc_b200 import synthetic
    from dgvcc_b200.datasets import den_targets
    from dgvcc_b200.utils import dmap_gen
    from oracle import dmap_oracle
    h, w = 600, 800
    pts = synthetic.crowd_points(np.random.default_rng(5), 400, w, h, dtype=np.float64)
    dev = torch.device("cuda")
    full = dmap_gen._density_device(h, w, pts, False, dev)
    d, b = den_targets.train_density_targets([full], [(0, 0, 100, 200, 1)], (256, 256), 8)
    ref_full = dmap_oracle.density_closed_form((h, w), pts, fixed=True)
    ref = do.train_density(ref_full, 0, 0, 100, 200, 256, 256, 8, 1)
    np.testing.assert_allclose(d[0].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=0)
    assert np.array_equal(b[0].cpu().numpy(), do.block_occupancy(ref)[0].numpy())


def test_argument_errors():
    from dgvcc_b200.datasets import den_targets
    m = np.zeros((64, 64), dtype=np.float32)
    with pytest.raises(RuntimeError):
        den_targets.train_density_targets([m], [(0, 0, 0, 0, 0)], (60, 64), 8)       # reshape would raise
    with pytest.raises(RuntimeError):
        den_targets.train_density_targets([m], [(0, 0, 0, 0, 0)], (64, 64), 8)       # 8x8 pooled map, no 16x16 blocks
