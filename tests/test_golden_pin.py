"""The fixtures under tests/golden/ ARE outputs of the unmodified reference: wherever the reference's own files are at
hand (/root/reference in the authoring container, or the byte copies oracle/make_ref.sh leaves in oracle/_ref), re-run
them on the inputs stored in the fixtures and compare bit for bit.  CPU only; skipped where neither copy exists.

This is the check SURVEY 8c asks for ("pin the oracle against outputs of the reference itself run here") kept alive as a
test instead of a one-off: a fixture that was edited by hand, or a reference file that is not the one the fixtures came
from, fails here.  (tests/golden/make_golden.py is the generator; this file reads only what it wrote.)
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

needs_bl = pytest.mark.skipif(not ref_loader.available("bl"), reason="no copy of the reference's losses/bl.py here")
needs_dmap = pytest.mark.skipif(not ref_loader.available("dmap_gen"), reason="no copy of the reference's utils/dmap_gen.py here")
needs_isw = pytest.mark.skipif(not ref_loader.available("instance_whitening"),
                               reason="no copy of the reference's models/ISW/instance_whitening.py here")


def test_ref_copies_are_byte_identical_to_the_reference():
    """oracle/_ref (what travels to the GPU box and what bench.py's reference legs time) == /root/reference."""
    if not os.path.isdir("/root/reference") or not os.path.isdir(os.path.join(os.path.dirname(ref_loader.__file__), "_ref")):
        pytest.skip("needs both /root/reference and oracle/_ref")
    for rel in ref_loader.FILES.values():
        a = open(os.path.join("/root/reference", rel), "rb").read()
        b = open(os.path.join(os.path.dirname(ref_loader.__file__), "_ref", rel), "rb").read()
        assert a == b, rel


@needs_bl
@pytest.mark.parametrize("name", ["c1", "mixed", "nobg", "sigma10", "outside", "empty"])
def test_bl_fixture_is_what_the_unmodified_reference_returns(name):
    """losses/bl.py (BL = Post_Prob + Bay_Loss) on the fixture's inputs: loss and density gradient, bit for bit.
    The reference is square-only, so the generator zero-pads the density to the square grid (make_golden.bl_case);
    the stored gradient is the crop of the square one."""
    z = np.load(os.path.join(GOLDEN, f"bl_{name}.npz"))
    bl = ref_loader.load("bl")
    width, height, stride = int(z["in_width"]), int(z["in_height"]), int(z["in_stride"])
    counts = z["in_counts"]
    c_size = max(width, height)
    g, hp, wp = c_size // stride, height // stride, width // stride
    sq = np.zeros((len(counts), 1, g, g), dtype=np.float32)
    sq[:, :, :hp, :wp] = z["in_density"]
    mod = bl.BL(float(z["in_sigma"]), c_size, stride, float(z["in_bg_ratio"]), bool(int(z["in_use_bg"])), "cpu")
    d = torch.from_numpy(sq).requires_grad_(True)
    pts = [torch.from_numpy(z[f"in_points_{i}"].copy()) for i in range(len(counts))]
    tgt = [torch.from_numpy(z[f"in_targets_{i}"].copy()) for i in range(len(counts))]
    loss = mod(pts, torch.from_numpy(z["in_st_sizes"].copy()), tgt, d)
    loss.backward()
    assert np.array_equal(loss.detach().numpy(), z["ref_loss"]), (float(loss), float(z["ref_loss"]))
    grad = d.grad.numpy()[:, :, :hp, :wp]
    assert grad.shape == z["ref_grad"].shape and np.array_equal(grad, z["ref_grad"])
    # nothing of the gradient falls on the zero padding's side of the crop that the fixture does not store
    outside = d.grad.numpy().copy()
    outside[:, :, :hp, :wp] = 0
    assert np.isfinite(outside).all()


@needs_dmap
def test_dmap_fixture_is_what_the_unmodified_reference_returns():
    """utils/dmap_gen.py: gaussian_filter_density (kNN sigma) and gaussian_filter_density_fixed on every stored case."""
    z = np.load(os.path.join(GOLDEN, "dmap_cases.npz"))
    gen = ref_loader.load("dmap_gen")
    cases = sorted(k[:-len("_adaptive")] for k in z.files if k.endswith("_adaptive"))
    assert len(cases) >= 8
    for c in cases:
        shape = tuple(int(v) for v in z[f"{c}_shape"])
        img = np.zeros(shape + (3,), dtype=np.uint8)     # the reference reads only img.shape[:2] (dmap_gen.py:25)
        pts = z[f"{c}_points"]
        assert np.array_equal(gen.gaussian_filter_density(img, pts.copy()), z[f"{c}_adaptive"]), c
        assert np.array_equal(gen.gaussian_filter_density_fixed(img, pts.copy()), z[f"{c}_fixed"]), c


@needs_isw
def test_isw_fixture_is_what_the_unmodified_reference_returns():
    """models/ISW/instance_whitening.py: InstanceWhitening forward (norm + covariance) and instance_whitening_loss with
    its gradient into x, on every stored case."""
    z = np.load(os.path.join(GOLDEN, "isw_cases.npz"))
    iw = ref_loader.load("instance_whitening")
    cases = sorted(k[:-len("_grad_x")] for k in z.files if k.endswith("_grad_x"))
    assert len(cases) >= 4
    for c in cases:
        x = torch.from_numpy(z[f"{c}_x"].copy()).requires_grad_(True)
        dim = x.shape[1]
        eye, mask = torch.eye(dim), torch.from_numpy(z[f"{c}_mask"].copy())
        margin, num_remove = float(z[f"{c}_margin"]), mask.sum()
        norm, w = iw.InstanceWhitening(dim)(x)
        cov, _ = iw.get_covariance_matrix(w, eye=eye)
        loss = iw.instance_whitening_loss(w, eye, mask, margin, num_remove)
        loss.backward()
        assert np.array_equal(norm.detach().numpy(), z[f"{c}_norm"]), c
        assert np.array_equal(cov.detach().numpy(), z[f"{c}_cov"]), c
        assert np.array_equal(loss.detach().numpy(), z[f"{c}_loss"]), c
        assert np.array_equal(x.grad.numpy(), z[f"{c}_grad_x"]), c


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="tests/golden/make_golden.py reads /root/reference")
@pytest.mark.parametrize("family", ["bl", "dmap", "isw", "bay", "den", "cov", "aux", "sw"])
def test_generator_script_reproduces_the_committed_fixtures(family, tmp_path, monkeypatch):
    """`python tests/golden/make_golden.py <family>` -- the committed script that made the fixtures, running the
    UNMODIFIED reference classes -- writes, into a scratch directory, files whose every array equals the committed one
    bit for bit (inputs AND reference outputs; same keys, shapes, dtypes).  Covers all eight fixture families, i.e. also
    the SURVEY 8f rows (dataset targets, mask bookkeeping, lw / ortho, switchable whitening incl. its two-rank gloo runs)."""
    import sys
    monkeypatch.syspath_prepend(GOLDEN)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    import make_golden
    monkeypatch.setattr(make_golden, "HERE", str(tmp_path))
    torch.manual_seed(0)
    getattr(make_golden, f"make_{family}")()
    made = sorted(os.listdir(tmp_path))
    assert made, "the generator wrote nothing"
    for name in made:
        new, old = np.load(tmp_path / name, allow_pickle=True), np.load(os.path.join(GOLDEN, name), allow_pickle=True)
        assert sorted(new.files) == sorted(old.files), name
        for k in new.files:
            a, b = new[k], old[k]
            assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True), (name, k)
