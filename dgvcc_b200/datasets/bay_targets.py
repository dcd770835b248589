"""Target preparation of the reference's ``BayesianDataset`` on B200.

    cal_dists(pts)                         <- BayesianDataset._cal_dists          (datasets/bay_dataset.py:38-48)
    crop_targets(gt, dists, i, j, h, w)    <- crop block of _train_transform      (datasets/bay_dataset.py:85-107)

Host numpy in / host numpy out with the reference's shapes and dtypes, so ``BayesianDataset`` can call them
in place of its numpy code; the O(N^2) neighbour search never builds the N x N matrix.

Where to call them.  The reference runs this code inside ``__getitem__`` in FORKED DataLoader workers
(``num_workers=16`` in its configs) after the parent has initialised CUDA; a forked child cannot use CUDA, so
these functions raise a clear error there.  Use them (i) in the main process -- ``num_workers=0``, or from the
``collate_fn`` / training loop on the whole batch, the way ``den_targets`` works -- or (ii) in workers started
with ``multiprocessing_context='spawn'`` (each worker then owns a CUDA context and pays one small H2D, kernel
and D2H per sample, which only pays off for crowded images: 12 000 heads cost 0.5 ms here against 1.9 s in numpy).
"""
import numpy as np
import torch

from .. import _native


def _dev(device):
    if torch.cuda._is_in_bad_fork():
        raise RuntimeError(
            "dgvcc_b200.datasets.bay_targets was called in a forked child of a process that had already initialised "
            "CUDA (a DataLoader worker with the default 'fork' start method).  Prepare the targets in the main process "
            "/ collate_fn, or start the workers with multiprocessing_context='spawn' (see the module docstring).")
    if not torch.cuda.is_available():
        raise RuntimeError("dgvcc_b200.datasets.bay_targets needs a CUDA device; there is no CPU path")
    return torch.device(device if device is not None else "cuda")


def _as_xy(a):
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def cal_dists(pts, device=None):
    """[N,1] mean distance to the 3 nearest heads (dtype of ``pts``); the reference's constants for N < 2."""
    if len(pts) == 0:
        return np.array([[]])
    if len(pts) == 1:
        return np.array([[4.0]])
    dev = _dev(device)
    p = _as_xy(pts)
    n = len(p)
    d_pts = torch.from_numpy(p).pin_memory().to(dev, non_blocking=True)
    out = torch.empty((n, 1), dtype=d_pts.dtype, device=dev)
    _native.check(_native.lib().dgvcc_bay_knn_mean(_native.ptr(d_pts), n, int(p.dtype == np.float64), _native.ptr(out),
                                                   _native.stream_ptr(dev)), "dgvcc_bay_knn_mean")
    return out.cpu().numpy()


def crop_targets(gt, dists, i, j, h, w, device=None):
    """(gt_kept [K,2] float64 in crop coordinates, mirrored in x; targ [K]) for the crop window rows i..i+h, cols j..j+w."""
    if len(gt) == 0:
        return gt, np.array([])
    dev = _dev(device)
    g = _as_xy(gt)
    d = np.ascontiguousarray(np.asarray(dists).reshape(-1), dtype=g.dtype)
    n = len(g)
    d_gt = torch.from_numpy(g).pin_memory().to(dev, non_blocking=True)
    d_d = torch.from_numpy(d).pin_memory().to(dev, non_blocking=True)
    gt_out = torch.empty((n, 2), dtype=torch.float64, device=dev)
    targ = torch.empty((n,), dtype=d_gt.dtype, device=dev)
    kept = torch.zeros((1,), dtype=torch.int32, device=dev)
    _native.check(_native.lib().dgvcc_bay_crop_targets(
        _native.ptr(d_gt), _native.ptr(d_d), n, int(g.dtype == np.float64), float(j), float(i), float(j + w), float(i + h),
        _native.ptr(gt_out), _native.ptr(targ), _native.ptr(kept), _native.stream_ptr(dev)), "dgvcc_bay_crop_targets")
    k = int(kept.cpu())
    return gt_out[:k].cpu().numpy(), targ[:k].cpu().numpy()
