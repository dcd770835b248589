"""Parity of the auxiliary Gram losses (lw_loss, ortho_loss) through the C ABI with the fixtures
produced by the unmodified reference files and with the oracle.

Gates: loss rtol 1e-5; gradients rtol 1e-4 with atol 2e-6 * max|ref| (sums over C terms with cancellation, 3xTF32
Gram: the same gate as the ISW loss gradients)."""
import os

import numpy as np
import pytest
import torch

from oracle import aux_losses_oracle as ao
from helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "aux_cases.npz"))


def close(got, ref, what, rtol=1e-4):
    ref = np.asarray(ref)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=2e-6 * float(np.abs(ref).max()), err_msg=what)


@pytest.mark.parametrize("tc", [1, 0])
@pytest.mark.parametrize("name", ["lw_a", "lw_b", "lw_c", "lw_d"])
def test_lw_loss_matches_reference_fixture(fixtures, name, tc, monkeypatch):
    from dgvcc_b200.losses.lw import lw_loss
    monkeypatch.setenv("DGVCC_ISW_TENSOR_CORES", str(tc))
    x = torch.from_numpy(fixtures[f"{name}_x"]).cuda().requires_grad_(True)
    mask = torch.from_numpy(fixtures[f"{name}_mask"]).cuda() if f"{name}_mask" in fixtures else None
    loss = lw_loss(x, mask)
    loss.backward()
    np.testing.assert_allclose(loss.item(), fixtures[f"{name}_loss"], rtol=1e-5)
    close(x.grad.cpu().numpy(), fixtures[f"{name}_grad"], f"{name} grad")


@pytest.mark.parametrize("tc", [1, 0])
@pytest.mark.parametrize("name", ["or_a", "or_b", "or_c", "or_d"])
def test_ortho_loss_matches_reference_fixture(fixtures, name, tc, monkeypatch):
    from dgvcc_b200.losses.ortho import ortho_loss
    monkeypatch.setenv("DGVCC_ISW_TENSOR_CORES", str(tc))
    x = torch.from_numpy(fixtures[f"{name}_x"]).cuda().requires_grad_(True)
    y = torch.from_numpy(fixtures[f"{name}_y"]).cuda().requires_grad_(True)
    loss = ortho_loss(x, y)
    loss.backward()
    np.testing.assert_allclose(loss.item(), fixtures[f"{name}_loss"], rtol=1e-5)
    close(x.grad.cpu().numpy(), fixtures[f"{name}_gx"], f"{name} grad x")
    close(y.grad.cpu().numpy(), fixtures[f"{name}_gy"], f"{name} grad y")


def test_vgg_shaped_maps_against_oracle():
    """BASELINE config-5 sized feature maps (the shapes the ISW loss sees) through lw_loss, scaled upstream gradient."""
    from dgvcc_b200.losses.lw import lw_loss
    g = torch.Generator().manual_seed(12)
    for (n, c, h, w) in [(4, 64, 40, 40), (2, 256, 20, 20)]:
        x = torch.randn(n, c, h, w, generator=g)
        mask = (torch.rand(n, 1, h, w, generator=g) < 0.5).float()
        xg = x.clone().cuda().requires_grad_(True)
        (0.37 * lw_loss(xg, mask.cuda())).backward()
        xo = x.clone().requires_grad_(True)
        lo = 0.37 * ao.lw_loss(xo, mask)
        lo.backward()
        close(xg.grad.cpu().numpy(), xo.grad.numpy(), f"lw {n}x{c}x{h}x{w}")


def test_cpu_tensors_are_rejected():
    from dgvcc_b200.losses.lw import lw_loss
    from dgvcc_b200.losses.ortho import ortho_loss
    with pytest.raises(RuntimeError):
        lw_loss(torch.randn(1, 32, 4, 4))
    with pytest.raises(RuntimeError):
        ortho_loss(torch.randn(32, 16), torch.randn(32, 16))
