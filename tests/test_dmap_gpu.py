"""Parity of the CUDA density-map generator (through the C ABI) with the reference fixtures and the oracle.

Gates: kNN neighbour indices and in-bounds bookkeeping bit-exact; sigma <= 1 ulp;
density maps rtol 1e-5 with atol 1e-7*max|ref| (fp64 exp on the GPU may differ from numpy's in the last
ulp; after the two fp32 roundings that is almost always invisible -- the tests also report how many
pixels are not bit-identical)."""
import os

import numpy as np
import pytest

from dgvcc_b200 import synthetic
from oracle import dmap_oracle
from helpers import GOLDEN

pytestmark = pytest.mark.gpu

CASES = ["a40", "a4", "a3", "a1", "a0", "a25f32", "oob", "dup"]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "dmap_cases.npz"))


def assert_map_close(got, ref, what):
    assert got.dtype == np.float32 and got.shape == ref.shape
    tol = 1e-5 * np.abs(ref) + 1e-7 * float(np.abs(ref).max() if ref.size else 0)
    bad = np.abs(got.astype(np.float64) - ref) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} pixels out of tolerance, max err {np.abs(got - ref).max()}"
    return int((got != ref).sum())


@pytest.mark.parametrize("name", CASES)
def test_maps_match_reference_fixture(fixtures, name):
    from dgvcc_b200.utils import dmap_gen
    shape = tuple(int(v) for v in fixtures[f"{name}_shape"])
    pts = fixtures[f"{name}_points"]
    img = np.zeros(shape + (3,), dtype=np.uint8)
    diff_a = assert_map_close(dmap_gen.gaussian_filter_density(img, pts), fixtures[f"{name}_adaptive"], f"{name} adaptive")
    diff_f = assert_map_close(dmap_gen.gaussian_filter_density_fixed(img, pts), fixtures[f"{name}_fixed"], f"{name} fixed")
    print(f"{name}: pixels not bit-identical: adaptive {diff_a}, fixed {diff_f} of {shape[0] * shape[1]}")


@pytest.mark.parametrize("name", ["knn2000", "knn700f32"])
def test_knn_bit_exact(fixtures, name):
    from dgvcc_b200.utils import dmap_gen
    pts = fixtures[f"{name}_points"]
    dist, loc, sigma = dmap_gen.knn_sigma(pts)
    assert np.array_equal(loc, fixtures[f"{name}_loc"]), "neighbour indices"
    assert np.array_equal(dist, fixtures[f"{name}_dist"]), "neighbour distances"
    ref_sigma = (fixtures[f"{name}_dist"][:, 1] + fixtures[f"{name}_dist"][:, 2] + fixtures[f"{name}_dist"][:, 3]) * 0.1
    assert np.array_equal(sigma, ref_sigma)


def test_few_points_knn_convention():
    from dgvcc_b200.utils import dmap_gen
    pts = np.array([[3.0, 4.0], [10.0, 4.0]])
    dist, loc, sigma = dmap_gen.knn_sigma(pts)
    rd, rl = dmap_oracle.knn4(pts)
    assert np.array_equal(dist, rd) and np.array_equal(loc, rl) and np.array_equal(sigma, [15.0, 15.0])


@pytest.mark.parametrize("seed,n,h,w,dtype", [
    (1, 1500, 768, 1024, np.float64),     # ShanghaiTech-like
    (2, 6000, 1100, 1700, np.float32),    # sides that are no multiple of the tile
    (3, 8, 600, 900, np.float64),         # sparse: huge sigmas, stamps larger than the image
    (4, 25000, 2048, 2048, np.float64),   # JHU-shaped maximum of BASELINE config 4
])
def test_full_size_against_closed_form_oracle(seed, n, h, w, dtype):
    from dgvcc_b200.utils import dmap_gen
    rng = np.random.default_rng(4000 + seed)
    pts = synthetic.crowd_points(rng, n, w, h, dtype=dtype)
    img = np.empty((h, w, 0))
    dist, loc, sigma = dmap_gen.knn_sigma(pts)
    rd, rl = dmap_oracle.knn4(pts)
    assert len(np.unique(pts, axis=0)) == n, "synthetic heads must be distinct (tie order is implementation-defined)"
    assert np.array_equal(loc, rl) and np.array_equal(dist, rd)
    ref = dmap_oracle.density_closed_form((h, w), pts, sigmas=(rd[:, 1] + rd[:, 2] + rd[:, 3]) * 0.1)
    got = dmap_gen.gaussian_filter_density(img, pts)
    diff = assert_map_close(got, ref, "adaptive")
    ref_f = dmap_oracle.density_closed_form((h, w), pts, fixed=True)
    diff_f = assert_map_close(dmap_gen.gaussian_filter_density_fixed(img, pts), ref_f, "fixed")
    print(f"n={n} {h}x{w}: pixels not bit-identical: adaptive {diff}, fixed {diff_f} of {h * w}")
    # mass conservation away from the border: every in-bounds head contributes <= 1
    assert got.sum(dtype=np.float64) <= n * (1 + 1e-6)


def test_sparse_images_with_huge_sigma():
    """4..10 heads spread over a 2048 px image: sigma in the hundreds, kernel radius in the thousands
    (deep pairwise splits of the normaliser sum)."""
    from dgvcc_b200.utils import dmap_gen
    for n, seed in ((4, 1), (5, 2), (7, 3), (10, 4)):
        rng = np.random.default_rng(seed)
        h, w = 2048, 1900
        pts = rng.uniform([0, 0], [w - 1, h - 1], size=(n, 2))
        ref = dmap_oracle.density_closed_form((h, w), pts)
        got = dmap_gen.gaussian_filter_density(np.empty((h, w, 0)), pts)
        assert_map_close(got, ref, f"sparse n={n}")


def test_batch_api_and_file_protocol(tmp_path):
    from dgvcc_b200.utils import dmap_gen
    from PIL import Image
    rng = np.random.default_rng(9)
    shapes = [(120, 160), (90, 90), (64, 200)]
    plist = [synthetic.crowd_points(rng, n, s[1], s[0], dtype=np.float64) for n, s in zip((50, 0, 7), shapes)]
    outs = dmap_gen.gaussian_filter_density_batch(shapes, plist, fixed=True)
    for s, p, o in zip(shapes, plist, outs):
        assert_map_close(o, dmap_oracle.density_closed_form(s, p, fixed=True), "batch")
    # run(): <name>.jpg + <name>.npy -> <name>_dmap.npy, idempotent (dmap_gen.py:83-95)
    fn = tmp_path / "img_7.jpg"
    Image.fromarray(np.zeros((120, 160, 3), dtype=np.uint8)).save(fn)
    np.save(tmp_path / "img_7.npy", plist[0])
    dmap_gen.run(str(fn))
    saved = np.load(tmp_path / "img_7_dmap.npy")
    assert_map_close(saved, dmap_oracle.density_closed_form((120, 160), plist[0], fixed=True), "run()")
    mtime = os.path.getmtime(tmp_path / "img_7_dmap.npy")
    dmap_gen.run(str(fn))
    assert os.path.getmtime(tmp_path / "img_7_dmap.npy") == mtime


def test_negative_beyond_size_raises_like_numpy():
    from dgvcc_b200.utils import dmap_gen
    with pytest.raises(IndexError):
        dmap_gen.gaussian_filter_density_fixed(np.empty((20, 20, 0)), np.array([[-30.0, 2.0]]))
