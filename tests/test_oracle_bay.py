"""The Bayesian-dataset target oracle vs fixtures produced by the unmodified reference (datasets/bay_dataset.py)."""
import os

import numpy as np
import pytest

from oracle import bay_targets_oracle as bo
from helpers import GOLDEN

DIST_CASES = ["n0", "n1", "n2", "n3", "n4", "n300", "n200f32"]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "bay_cases.npz"))


@pytest.mark.parametrize("name", DIST_CASES)
def test_cal_dists(fixtures, name):
    ref = fixtures[f"dist_{name}_ref"]
    got = bo.cal_dists(fixtures[f"dist_{name}_pts"])
    assert got.shape == ref.shape and got.dtype == ref.dtype
    np.testing.assert_allclose(got, ref, rtol=1e-12 if ref.dtype == np.float64 else 1e-6)


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_crop_targets(fixtures, k):
    i, j, h, w = (int(v) for v in fixtures[f"crop_{k}_ijhw"])
    gt, targ = bo.crop_targets(fixtures[f"crop_{k}_gt"].copy(), fixtures[f"crop_{k}_dists"], i, j, h, w)
    ref_gt, ref_targ = fixtures[f"crop_{k}_ref_gt"], fixtures[f"crop_{k}_ref_targ"]
    assert len(targ) == len(ref_targ)
    if len(ref_targ):
        np.testing.assert_array_equal(np.asarray(gt, dtype=np.float64).astype(np.float32), ref_gt)
        np.testing.assert_array_equal(np.asarray(targ).astype(np.float32), ref_targ)
