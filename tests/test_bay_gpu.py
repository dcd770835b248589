"""Parity of the CUDA Bayesian-dataset target preparation with the reference fixtures and the oracle."""
import os

import numpy as np
import pytest

from dgvcc_b200 import synthetic
from oracle import bay_targets_oracle as bo
from helpers import GOLDEN

pytestmark = pytest.mark.gpu
DIST_CASES = ["n0", "n1", "n2", "n3", "n4", "n300", "n200f32"]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "bay_cases.npz"))


def rtol(dtype):
    return 1e-9 if dtype == np.float64 else 2e-3


def assert_dists_close(got, ref, pts):
    """The reference's expansion sq_i - 2 p.p + sq_j cancels: BLAS (FMA or not, summation order) moves d^2 by a few
    ulp of sq ~ |p|^2, i.e. d by  ulp(sq) / (2 d).  Allow 16 ulp(sq) on d^2 on top of the relative tolerance."""
    if ref.size == 0:
        assert got.size == 0
        return
    eps = np.finfo(ref.dtype).eps
    sq_max = float((np.asarray(pts, dtype=np.float64) ** 2).sum(1).max()) if len(pts) else 1.0
    tol = rtol(ref.dtype) * np.abs(ref) + 16 * eps * sq_max / np.maximum(2 * np.abs(ref), 1e-300)
    bad = np.abs(got.astype(np.float64) - ref) > tol
    assert not bad.any(), f"{int(bad.sum())} of {bad.size} distances off; worst {np.abs(got - ref).max()}"


@pytest.mark.parametrize("name", DIST_CASES)
def test_cal_dists_matches_reference_fixture(fixtures, name):
    from dgvcc_b200.datasets import bay_targets
    ref = fixtures[f"dist_{name}_ref"]
    got = bay_targets.cal_dists(fixtures[f"dist_{name}_pts"])
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert_dists_close(got, ref, fixtures[f"dist_{name}_pts"])


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_crop_targets_match_reference_fixture(fixtures, k):
    from dgvcc_b200.datasets import bay_targets
    i, j, h, w = (int(v) for v in fixtures[f"crop_{k}_ijhw"])
    gt, targ = bay_targets.crop_targets(fixtures[f"crop_{k}_gt"].copy(), fixtures[f"crop_{k}_dists"], i, j, h, w)
    ref_gt, ref_targ = fixtures[f"crop_{k}_ref_gt"], fixtures[f"crop_{k}_ref_targ"]
    assert len(targ) == len(ref_targ)          # the kept set (bookkeeping) is exact
    if len(ref_targ):
        np.testing.assert_array_equal(np.asarray(gt, dtype=np.float64).astype(np.float32), ref_gt)
        np.testing.assert_array_equal(np.asarray(targ).astype(np.float32), ref_targ)


@pytest.mark.parametrize("n,dtype", [(12000, np.float64), (5000, np.float32), (257, np.float64)])
def test_qnrf_size_against_oracle(n, dtype):
    """12 000 heads: the reference builds a 1.15 GB matrix; the kernel keeps four values per head in registers."""
    from dgvcc_b200.datasets import bay_targets
    pts = synthetic.crowd_points(np.random.default_rng(8200 + n), n, 2048, 1536, dtype=dtype)
    ref = bo.cal_dists(pts)
    got = bay_targets.cal_dists(pts)
    assert_dists_close(got, ref, pts)
    gt_ref, targ_ref = bo.crop_targets(pts.copy(), ref, 300, 500, 512, 512)
    gt, targ = bay_targets.crop_targets(pts.copy(), ref, 300, 500, 512, 512)
    assert len(targ) == len(targ_ref)
    np.testing.assert_array_equal(gt, gt_ref)
    np.testing.assert_allclose(targ, targ_ref, rtol=1e-12 if dtype == np.float64 else 1e-6)
