"""The density-target oracle vs fixtures produced by the unmodified reference (datasets/den_cls_dataset.py)."""
import os

import numpy as np
import pytest

from oracle import den_targets_oracle as do
from helpers import GOLDEN

CASES = [0, 1, 2, 3, 4, 5]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "den_cases.npz"))


@pytest.mark.parametrize("k", CASES)
def test_train_density_and_occupancy_bit_exact(fixtures, k):
    left, top, i, j, h, w, down, flip = (int(v) for v in fixtures[f"den_{k}_geom"])
    d = do.train_density(fixtures[f"den_{k}_dmap"], left, top, i, j, h, w, down, flip)
    assert np.array_equal(d.numpy(), fixtures[f"den_{k}_ref_dmap"])
    assert np.array_equal(do.block_occupancy(d).numpy(), fixtures[f"den_{k}_ref_bmap"])
