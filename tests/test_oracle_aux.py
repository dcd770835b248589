"""The lw / ortho loss oracle vs fixtures produced by the unmodified reference (losses/lw.py, losses/ortho.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import aux_losses_oracle as ao
from helpers import GOLDEN


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "aux_cases.npz"))


@pytest.mark.parametrize("name", ["lw_a", "lw_b", "lw_c", "lw_d"])
def test_lw_loss_bit_exact(fixtures, name):
    x = torch.from_numpy(fixtures[f"{name}_x"]).requires_grad_(True)
    mask = torch.from_numpy(fixtures[f"{name}_mask"]) if f"{name}_mask" in fixtures else None
    loss = ao.lw_loss(x, mask)
    loss.backward()
    assert np.array_equal(loss.detach().numpy(), fixtures[f"{name}_loss"])
    assert np.array_equal(x.grad.numpy(), fixtures[f"{name}_grad"])


@pytest.mark.parametrize("name", ["or_a", "or_b", "or_c", "or_d"])
def test_ortho_loss_bit_exact(fixtures, name):
    x = torch.from_numpy(fixtures[f"{name}_x"]).requires_grad_(True)
    y = torch.from_numpy(fixtures[f"{name}_y"]).requires_grad_(True)
    loss = ao.ortho_loss(x, y)
    loss.backward()
    assert np.array_equal(loss.detach().numpy(), fixtures[f"{name}_loss"])
    assert np.array_equal(x.grad.numpy(), fixtures[f"{name}_gx"]) and np.array_equal(y.grad.numpy(), fixtures[f"{name}_gy"])
