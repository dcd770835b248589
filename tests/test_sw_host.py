"""Host-side checks of the switchable-whitening drop-in that need no GPU: module state, workspace sizing, argument
errors of the C ABI."""
import pytest
import torch

from dgvcc_b200 import _native
from dgvcc_b200.models.ISW import SwitchWhiten2d, SyncSwitchWhiten2d


def test_state_matches_the_reference_class():
    """Parameters, buffers and their shapes as switchwhiten.py:46-71 registers them (after reset_parameters)."""
    m = SwitchWhiten2d(64, num_pergroup=16, sw_type=3)
    sd = m.state_dict()
    assert list(sd) == ["sw_mean_weight", "sw_var_weight", "weight", "bias", "running_mean", "running_cov"]
    assert sd["sw_mean_weight"].shape == (3,) and sd["running_mean"].shape == (4, 16, 1) and sd["running_cov"].shape == (4, 16, 16)
    assert float(sd["running_cov"].abs().sum()) == 0.0 and float(sd["weight"].sum()) == 64.0
    tied = SyncSwitchWhiten2d(32, sw_type=5, tie_weight=True, affine=False)
    assert list(tied.state_dict()) == ["sw_mean_weight", "running_mean", "running_cov"]
    assert tied.sw_var_weight is None and tied.weight is None and tied.bias is None
    assert "num_pergroup=16" in repr(tied) and "sw_type=5" in repr(tied)
    with pytest.raises(ValueError):
        SwitchWhiten2d(32, sw_type=4)
    with pytest.raises(ValueError):
        SyncSwitchWhiten2d(32, sw_type=1)


def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError, match="no CPU path"):
        SwitchWhiten2d(32)(torch.randn(1, 32, 4, 4))


def test_workspace_sizing_and_argument_errors():
    lib = _native.lib()
    assert lib.dgvcc_sw_workspace_bytes(8, 64, 160 * 160, 16) > 0
    assert lib.dgvcc_sw_workspace_bytes(8, 64, 160 * 160, 12) == 0        # groups of 4, 8 or 16 channels
    assert lib.dgvcc_sw_workspace_bytes(8, 60, 100, 16) == 0              # channels not a multiple of the group
    assert lib.dgvcc_sw_workspace_bytes(0, 64, 100, 16) == 0
    # grows with the batch, stays small next to the activations (fp64 statistics + split partials only)
    small, big = lib.dgvcc_sw_workspace_bytes(2, 64, 1600, 16), lib.dgvcc_sw_workspace_bytes(8, 64, 1600, 16)
    assert small < big < 8 * 64 * 1600 * 4
    assert lib.dgvcc_sw_workspace_bytes(8, 512, 40 * 40, 16) % 256 == 0
    # null pointers are argument errors before anything touches the device
    null = _native.ptr(None)
    assert lib.dgvcc_sw_instance_stats(null, 1, 16, 4, 16, null, null, null, 0, null) == -1
    assert lib.dgvcc_sw_batch_mean(null, 1, 16, null, null) == -1
    assert lib.dgvcc_sw_batch_cov(null, null, null, 1, 16, 16, null, null) == -1
    assert lib.dgvcc_sw_update_running(null, null, null, null, 16, 16, 0.99, 0.01, null) == -1
