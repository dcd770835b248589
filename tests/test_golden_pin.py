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


# ------------------------------------------------------------------------------------------------------------------
# Beyond the fixtures: the oracle restatements against the unmodified reference on RANDOM inputs, evaluated live
# (small sizes: the whole block runs in seconds).  The fixtures pin a dozen hand-picked cases; these pin the
# restatements on shapes nobody picked.

@needs_bl
@pytest.mark.parametrize("seed", range(8))
def test_bl_oracle_against_the_live_reference_on_random_inputs(seed):
    """oracle.bl_oracle (rectangular grids, per-image evaluation) == losses/bl.py on the zero-padded square grid:
    posteriors bit for bit on the reference's own square grids and within 2 ulp on rectangular ones (the arithmetic is
    per-pixel independent, torch's SIMD CPU kernels are not position independent), loss / gradient to 2e-6."""
    from oracle import bl_oracle
    rng = np.random.default_rng(9000 + seed)
    bl = ref_loader.load("bl")
    stride = int(rng.choice([4, 8, 16]))
    hp, wp = int(rng.integers(2, 20)), int(rng.integers(2, 20))
    height, width = hp * stride, wp * stride
    sigma = float(rng.choice([4.0, 8.0, 6.5, 10.0, 15.0]))
    use_bg, bg_ratio = bool(rng.integers(0, 2)), float(rng.choice([1.0, 0.15, 2.0]))
    b = int(rng.integers(1, 5))
    counts = [int(rng.choice([0, 1, 2, 3, 11, int(rng.integers(4, 60))])) for _ in range(b)]
    pts = [torch.from_numpy((rng.random((n, 2)) * [width + 12, height + 12] - 6).astype(np.float32)) for n in counts]  # some outside
    tgt = [torch.from_numpy(rng.uniform(0.3, 1.0, n).astype(np.float32)) for n in counts]
    dens = torch.from_numpy(np.abs(rng.normal(size=(b, 1, hp, wp))).astype(np.float32))
    st = torch.full((b,), float(min(width, height)))
    c_size = max(width, height)
    g = c_size // stride
    sq = torch.zeros((b, 1, g, g))
    sq[:, :, :hp, :wp] = dens
    sq.requires_grad_(True)
    mod = bl.BL(sigma, c_size, stride, bg_ratio, use_bg, "cpu")
    ref_loss = mod([p.clone() for p in pts], st, tgt, sq)
    ref_loss.backward()
    ref_grad = sq.grad[:, :, :hp, :wp]
    loss, grad, _ = bl_oracle.bl_forward_backward(pts, st, tgt, dens, stride, sigma, bg_ratio, use_bg)
    torch.testing.assert_close(loss, ref_loss.detach(), rtol=2e-6, atol=0)
    torch.testing.assert_close(grad, ref_grad, rtol=2e-6, atol=2e-7 * float(ref_grad.abs().max()) + 1e-30)
    ref_probs = mod.post_prob([p.clone() for p in pts], st)
    for i, n in enumerate(counts):
        mine = bl_oracle.posterior(pts[i], st[i], hp, wp, stride, sigma, bg_ratio, use_bg)
        if n == 0:
            assert mine is None and ref_probs[i] is None
        else:
            mine, ref = mine.view(-1, hp, wp), ref_probs[i].view(-1, g, g)[:, :hp, :wp]
            if hp == wp == g:      # the reference's own layout: the same torch kernels see the same memory positions
                assert torch.equal(mine, ref), i
            else:
                # a crop of the padded square: torch's vectorised CPU exp / softmax round an element by whether it
                # falls into a SIMD body or a scalar tail, i.e. by its position in memory -- measured <= 2.5e-7 (2 ulp)
                torch.testing.assert_close(mine, ref, rtol=5e-7, atol=1e-37)


@needs_dmap
@pytest.mark.parametrize("seed", range(6))
def test_dmap_closed_form_against_the_live_reference_on_random_inputs(seed):
    """oracle.dmap_oracle.density_closed_form -- the checker of every full-size map -- == utils/dmap_gen.py, bit for
    bit, on random small images: heads on the border, outside the image, duplicated, fp32 and fp64 coordinates."""
    from oracle import dmap_oracle
    rng = np.random.default_rng(9100 + seed)
    gen = ref_loader.load("dmap_gen")
    h, w = int(rng.integers(8, 70)), int(rng.integers(8, 70))
    n = int(rng.choice([0, 1, 3, 4, 5, int(rng.integers(6, 40))]))
    dtype = np.float32 if seed % 3 == 2 else np.float64
    pts = (rng.random((n, 2)) * [w * 1.15, h * 1.15]).astype(dtype)          # ~25 % of the heads right of / below the image
    if n >= 6:
        pts[1] = pts[0]                                                      # a duplicate: zero nearest distance
        pts[2] = [w - 1, h - 1]                                              # the last pixel
        pts[3] = [0, 0]
    img = np.zeros((h, w, 3), dtype=np.uint8)
    assert np.array_equal(dmap_oracle.density_closed_form((h, w), pts), gen.gaussian_filter_density(img, pts.copy()))
    assert np.array_equal(dmap_oracle.density_closed_form((h, w), pts, fixed=True), gen.gaussian_filter_density_fixed(img, pts.copy()))


@needs_isw
@pytest.mark.parametrize("seed", range(4))
def test_isw_oracle_against_the_live_reference_on_random_inputs(seed):
    from oracle import isw_oracle
    iw = ref_loader.load("instance_whitening")
    g = torch.Generator().manual_seed(9200 + seed)
    b, c, h, w = [int(v) for v in (torch.randint(1, 5, (1,), generator=g), torch.randint(2, 40, (1,), generator=g),
                                   torch.randint(2, 12, (1,), generator=g), torch.randint(2, 12, (1,), generator=g))]
    x = (torch.randn((b, c, h, w), generator=g) * 1.5 + 0.2).requires_grad_(True)
    eye, mask = torch.eye(c), isw_oracle.upper_mask(c, 0.5, 9300 + seed)
    margin, num = (0.0, 0.01)[seed % 2], mask.sum().clamp_min(1)
    _, wt = iw.InstanceWhitening(c)(x)
    ref = iw.instance_whitening_loss(wt, eye, mask, margin, num)
    ref.backward()
    x2 = x.detach().clone().requires_grad_(True)
    mine = isw_oracle.whitening_loss(isw_oracle.instance_standardize(x2), eye, mask, margin, num)
    mine.backward()
    torch.testing.assert_close(mine, ref.detach(), rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(x2.grad, x.grad, rtol=1e-5, atol=1e-6 * float(x.grad.abs().max()) + 1e-12)
