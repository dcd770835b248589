"""Host-only checks of the batched density-map plan (dgvcc_dmap_batch_plan makes no CUDA call): the per-image table
the kernels search, the launch totals and the workspace layout."""
import ctypes

import numpy as np

from dgvcc_b200 import _native

FINE, COARSE, CHUNK, KNN_T = 32, 256, 2048, 256


def plan(shapes, counts):
    lib = _native.lib()
    b = len(shapes)
    h = np.ascontiguousarray([s[0] for s in shapes], dtype=np.int32)
    w = np.ascontiguousarray([s[1] for s in shapes], dtype=np.int32)
    c = np.ascontiguousarray(counts, dtype=np.int32)
    meta = np.zeros((b + 1, _native.DMAP_META_COLS), dtype=np.int64)
    pl = _native.DmapPlan()
    rc = lib.dgvcc_dmap_batch_plan(b, h.ctypes.data, w.ctypes.data, c.ctypes.data, meta.ctypes.data, ctypes.byref(pl))
    return rc, meta, pl


def cdiv(a, b):
    return -(-a // b)


def test_plan_table_and_totals():
    rng = np.random.default_rng(1)
    for _ in range(50):
        b = int(rng.integers(1, 12))
        shapes = [(int(rng.integers(1, 2100)), int(rng.integers(1, 2100))) for _ in range(b)]
        counts = [int(rng.choice([0, 1, 3, 255, 256, 257, 2048, 2049, 9000, 25000])) for _ in range(b)]
        rc, meta, pl = plan(shapes, counts)
        assert rc == 0
        # per-image columns
        assert meta[:b, 1].tolist() == counts and meta[:b, 2].tolist() == [s[0] for s in shapes] and meta[:b, 3].tolist() == [s[1] for s in shapes]
        # cumulative columns start at 0 and never decrease; the last row holds the totals
        for col in (0, 4, 5, 6, 7, 8, 9, 10, 11):
            assert meta[0, col] == 0 and (np.diff(meta[:, col]) >= 0).all(), col
        assert meta[b, 0] == pl.total_heads == sum(counts)
        assert meta[b, 4] == pl.total_pixels == sum(h * w for h, w in shapes)
        assert meta[b, 5] == pl.fine_tiles == sum(cdiv(h, FINE) * cdiv(w, FINE) for h, w in shapes)
        ctiles = [cdiv(h, COARSE) * cdiv(w, COARSE) for h, w in shapes]
        assert meta[b, 6] == pl.coarse_tasks == sum(t * cdiv(n, CHUNK) for t, n in zip(ctiles, counts))
        assert meta[b, 7] == sum(t * n for t, n in zip(ctiles, counts))          # coarse-list capacity: every head once per tile
        assert meta[b, 8] == sum(ctiles)
        assert meta[b, 11] == pl.knn_query_blocks == sum(cdiv(n, KNN_T) for n in counts)
        # kNN slices: between 1 and max_slices per image with heads, partial lists sized slices * n * 4
        slices = np.diff(meta[:, 9]) // np.maximum(1, np.asarray([cdiv(n, KNN_T) for n in counts]))
        for s_, n in zip(slices, counts):
            assert (s_ == 0 and n == 0) or 1 <= s_ <= max(1, pl.knn_max_slices)
        assert np.diff(meta[:, 10]).tolist() == [int(s_) * n * 4 for s_, n in zip(slices, counts)]
        assert pl.knn_tasks == meta[b, 9]
        # workspace regions: aligned, ordered, inside the total
        offs = [pl.off_stamps, pl.off_boxes, pl.off_wtab, pl.off_fmask, pl.off_tmpl, pl.off_desc, pl.off_ccount, pl.off_ctotal,
                pl.off_clist]
        assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and pl.splat_workspace_bytes > offs[-1]
        assert pl.off_knn_d2 == 0 and pl.off_knn_idx % 256 == 0 and pl.knn_workspace_bytes > pl.off_knn_max > pl.off_knn_pts32 > pl.off_knn_idx


def test_plan_argument_errors():
    assert plan([], [])[0] == -1                       # DGVCC_ERR_ARG: empty batch
    assert plan([(0, 10)], [1])[0] == -1               # non-positive size
    assert plan([(10, 10)], [-1])[0] == -1             # negative count
    lib = _native.lib()
    assert lib.dgvcc_dmap_batch_plan(1, None, None, None, None, None) == -1
