"""The dmap oracle restatements vs fixtures produced by the unmodified reference (utils/dmap_gen.py)."""
import os

import numpy as np
import pytest

from oracle import dmap_oracle
from helpers import GOLDEN

CASES = ["a40", "a4", "a3", "a1", "a0", "a25f32", "oob", "dup"]


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "dmap_cases.npz"))


@pytest.mark.parametrize("name", CASES)
def test_closed_form_is_bit_identical_to_reference(fixtures, name):
    shape = tuple(fixtures[f"{name}_shape"])
    pts = fixtures[f"{name}_points"]
    for fixed in (False, True):
        ref = fixtures[f"{name}_{'fixed' if fixed else 'adaptive'}"]
        got = dmap_oracle.density_closed_form(shape, pts, fixed=fixed)
        assert got.dtype == np.float32 and got.shape == ref.shape
        assert np.array_equal(got, ref), f"{name} fixed={fixed}: {np.abs(got - ref).max()}"


@pytest.mark.parametrize("name", ["a4", "a1", "dup"])
def test_reference_like_matches_fixture(fixtures, name):
    shape = tuple(fixtures[f"{name}_shape"])
    pts = fixtures[f"{name}_points"]
    assert np.array_equal(dmap_oracle.density_reference_like(shape, pts), fixtures[f"{name}_adaptive"])
    assert np.array_equal(dmap_oracle.density_reference_like(shape, pts, fixed=True), fixtures[f"{name}_fixed"])


@pytest.mark.parametrize("name", ["knn2000", "knn700f32"])
def test_knn_bookkeeping(fixtures, name):
    pts = fixtures[f"{name}_points"]
    d, loc = dmap_oracle.knn4(pts)
    assert np.array_equal(loc, fixtures[f"{name}_loc"]) and np.array_equal(d, fixtures[f"{name}_dist"])
    # brute force in fp64 without FMA gives the same neighbours and distances (what the kernel does)
    p = pts.astype(np.float64)
    dx = p[:, None, 0] - p[None, :, 0]
    dy = p[:, None, 1] - p[None, :, 1]
    d2 = dx * dx + dy * dy
    order = np.argsort(d2, axis=1, kind="stable")[:, :4]
    assert np.array_equal(order, loc)
    assert np.array_equal(np.sqrt(np.take_along_axis(d2, order, 1)), d)
