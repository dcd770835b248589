"""The CovMatrix_ISW oracle vs fixtures produced by the unmodified reference (models/ISW/cov_settings.py)."""
import os

import numpy as np
import pytest
import torch

from oracle.cov_settings_oracle import CovMatrixISW
from helpers import GOLDEN

CASES = ["c16", "c64", "c64r3", "c256"]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "cov_cases.npz"))


@pytest.mark.parametrize("name", CASES)
def test_masks_bit_exact(fixtures, name):
    dim, relax, rounds, batches = fixtures[f"{name}_cfg"]
    cm = CovMatrixISW(int(dim), float(relax))
    for r in range(int(rounds)):
        for b in range(int(batches)):
            cm.set_variance_of_covariance(torch.from_numpy(fixtures[f"{name}_var_{r}_{b}"]))
        cm.set_mask_matrix()
        assert np.array_equal(cm.mask_matrix.numpy(), fixtures[f"{name}_mask_{r}"])
        num, margin_ret, margin, n_off = fixtures[f"{name}_num_{r}"]
        assert float(cm.num_sensitive) == num and float(cm.margin) == margin and float(cm.num_off_diagonal) == n_off
        assert margin_ret == 0   # get_mask_matrix always hands margin 0 to the loss (cov_settings.py:47)
