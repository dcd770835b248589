"""The fp32 filter of the batched kNN search (csrc/dmap_kernels.cu: knn_bound32) must never reject a candidate whose
exact fp64 squared distance is below the threshold.  Here the bound's formula is re-stated in numpy and checked against
an emulation of the kernel's fp32 evaluation over millions of random pairs, including the regimes that are hardest for
fp32 (large coordinates, tiny separations).  No GPU needed: this tests the mathematics of the bound."""
import numpy as np

U = np.float32(5.9604645e-8)


def up(x):
    """One fp32 step up: a conservative stand-in for the kernel's round-up intrinsics."""
    return np.nextafter(x.astype(np.float32), np.float32(np.inf))


def bound32(thr, big_m):
    t = up(thr.astype(np.float32))  # __double2float_ru
    um = up(U * big_m)
    e = up(up(np.float32(12) * um) * up(up(np.sqrt(t)) * np.float32(1.000001)))
    e = up(e + up(up(np.float32(16) * U) * t))
    e = up(e + up(np.float32(32) * up(um * um)))
    return up(t + e)


def fp32_sqdist(q, c):
    """The kernel's evaluation: inputs rounded to fp32, fx = qx - cx, fy = qy - cy, fma(fx, fx, fy * fy)."""
    qf, cf = q.astype(np.float32), c.astype(np.float32)
    fx, fy = qf[:, 0] - cf[:, 0], qf[:, 1] - cf[:, 1]
    yy = (fy * fy).astype(np.float32)
    return (fx.astype(np.float64) * fx.astype(np.float64) + yy.astype(np.float64)).astype(np.float32)  # one rounding = fma


def exact_sqdist(q, c):
    """The reference's fp64 evaluation (what decides the neighbour lists)."""
    dx, dy = q[:, 0] - c[:, 0], q[:, 1] - c[:, 1]
    return dx * dx + dy * dy


def test_filter_never_rejects_a_candidate_below_the_threshold():
    rng = np.random.default_rng(11)
    n = 400_000
    for scale, sep in [(2048, 2048), (2048, 1.0), (2048, 1e-3), (1e6, 300), (1e6, 1e-2), (5e4, 5e4), (30, 1e-5), (1e9, 1e3)]:
        q = rng.uniform(-scale, scale, size=(n, 2))
        c = q + rng.normal(0, sep, size=(n, 2))
        if scale == 2048:
            q, c = np.abs(q), np.abs(c)
        big_m = np.float32(np.nextafter(np.float32(max(np.abs(q).max(), np.abs(c).max())), np.float32(np.inf)))
        d2 = exact_sqdist(q, c)
        thr = np.nextafter(d2, np.inf)            # the tightest threshold this candidate is still below
        ok = fp32_sqdist(q, c) <= bound32(thr, big_m)
        assert ok.all(), f"scale {scale}, separation {sep}: {int((~ok).sum())} candidates would have been skipped"
        # and the filter is not vacuous: at image scale a candidate 1.01x (in squared distance: 2 %) farther than
        # the threshold is rejected most of the time
        if scale == 2048 and sep >= 1.0:
            far = fp32_sqdist(q, c) <= bound32(d2 / 1.02, big_m)
            assert far.mean() < 0.05
