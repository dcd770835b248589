"""The ISW oracle restatement vs fixtures produced by the unmodified reference (models/ISW/instance_whitening.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import isw_oracle
from helpers import GOLDEN, assert_close

CASES = ["c64", "c128", "c48", "margin"]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "isw_cases.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(fixtures, name):
    x = torch.from_numpy(fixtures[f"{name}_x"]).requires_grad_(True)
    mask = torch.from_numpy(fixtures[f"{name}_mask"])
    margin = float(fixtures[f"{name}_margin"])
    c = x.shape[1]
    eye = torch.eye(c)
    w = isw_oracle.instance_standardize(x)
    cov, b = isw_oracle.covariance(w, eye)
    loss = isw_oracle.whitening_loss(w, eye, mask, margin, mask.sum())
    loss.backward()
    assert b == x.shape[0]
    # same torch ops as the reference, but bmm / instance_norm may pick other CPU kernels per machine
    assert_close(w.detach(), fixtures[f"{name}_norm"], 1e-6, 1e-6, "instance norm")
    assert_close(cov.detach(), fixtures[f"{name}_cov"], 1e-6, 1e-7, "covariance")
    assert_close(loss.detach(), fixtures[f"{name}_loss"], 1e-6, 0, "loss")
    assert_close(x.grad, fixtures[f"{name}_grad_x"], 1e-5, 1e-6 * float(np.abs(fixtures[f"{name}_grad_x"]).max()), "grad x")
