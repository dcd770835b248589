"""Drop-in for the reference's ``utils/dmap_gen.py`` (density-map generation) on B200.

Same functions and call signatures as the reference (utils/dmap_gen.py:14, 53, 83):

    gaussian_filter_density(img, points)        -> np.ndarray [H, W] float32   (k=3 kNN-adaptive sigma)
    gaussian_filter_density_fixed(img, points)  -> np.ndarray [H, W] float32   (sigma 4, 15x15 stamps)
    run(img_fn)                                 <name>.jpg + <name>.npy -> <name>_dmap.npy
    python -m dgvcc_b200.utils.dmap_gen --path <root>

``img`` is only inspected for ``.shape[0:2]``; ``points`` is [N,2] (col,row), float32 or float64.
Host numpy in / host numpy out like the reference; the work runs in csrc/dmap_kernels.cu.
CUDA is initialised lazily on first call, so the functions are safe to import before a fork.
"""
import argparse
import os
from glob import glob

import numpy as np
import torch

from .. import _native

ADAPTIVE_TRUNCATE = 4.0      # scipy.ndimage.gaussian_filter default, dmap_gen.py:49
FIXED_SIGMA = 4.0            # dmap_gen.py:78
FIXED_TRUNCATE = 7 / 4       # dmap_gen.py:79: truncate=7/sigma


def _device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("dgvcc_b200.utils.dmap_gen needs a CUDA device; there is no CPU path")
    return torch.device(device if device is not None else "cuda")


def _check_points(points, height, width):
    pts = np.ascontiguousarray(np.asarray(points), dtype=np.float64)  # KDTree(points.copy()) works in float64
    if pts.ndim != 2 or pts.shape[1] != 2:
        raise ValueError(f"points must be [N,2], got {pts.shape}")
    # the reference writes pt2d[int(y), int(x)]: numpy raises IndexError below -size
    if len(pts) and (np.trunc(pts[:, 1]).min() < -height or np.trunc(pts[:, 0]).min() < -width):
        raise IndexError("point index out of bounds for the image (negative beyond its size)")
    return pts


def knn_sigma(points, device=None):
    """(distances [N,4] f64, locations [N,4] i64, sigma [N] f64) of dmap_gen.py:34-36,45-48, as host arrays."""
    dev = _device(device)
    pts = np.ascontiguousarray(np.asarray(points), dtype=np.float64)
    n = len(pts)
    if n == 0:
        return np.zeros((0, 4)), np.zeros((0, 4), dtype=np.int64), np.zeros((0,))
    d_pts = torch.from_numpy(pts).pin_memory().to(dev, non_blocking=True)
    idx = torch.empty((n, 4), dtype=torch.int32, device=dev)
    dist = torch.empty((n, 4), dtype=torch.float64, device=dev)
    sigma = torch.empty((n,), dtype=torch.float64, device=dev)
    kws_bytes = _native.lib().dgvcc_dmap_knn_workspace_bytes(n)
    kws = torch.empty((kws_bytes,), dtype=torch.uint8, device=dev)
    _native.check(_native.lib().dgvcc_dmap_knn_sigma(_native.ptr(d_pts), n, _native.ptr(idx), _native.ptr(dist),
                                                     _native.ptr(sigma), _native.ptr(kws), kws_bytes,
                                                     _native.stream_ptr(dev)), "dgvcc_dmap_knn_sigma")
    return dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64), sigma.cpu().numpy()


def _density_device(height, width, pts, adaptive, dev):
    """Device tensor [H,W] f32 for one image; ``pts`` is a host float64 [N,2] array."""
    lib = _native.lib()
    n = len(pts)
    out = torch.empty((height, width), dtype=torch.float32, device=dev)
    stream = _native.stream_ptr(dev)
    if n == 0:
        d_pts = sigma = ws = None
        ws_bytes = lib.dgvcc_dmap_workspace_bytes(0)
    else:
        d_pts = torch.from_numpy(pts).pin_memory().to(dev, non_blocking=True)
        ws_bytes = lib.dgvcc_dmap_workspace_bytes(n)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        sigma = None
        if adaptive:
            idx = torch.empty((n, 4), dtype=torch.int32, device=dev)
            dist = torch.empty((n, 4), dtype=torch.float64, device=dev)
            sigma = torch.empty((n,), dtype=torch.float64, device=dev)
            kws_bytes = lib.dgvcc_dmap_knn_workspace_bytes(n)
            kws = torch.empty((kws_bytes,), dtype=torch.uint8, device=dev)
            _native.check(lib.dgvcc_dmap_knn_sigma(_native.ptr(d_pts), n, _native.ptr(idx), _native.ptr(dist),
                                                   _native.ptr(sigma), _native.ptr(kws), kws_bytes, stream),
                          "dgvcc_dmap_knn_sigma")
    _native.check(lib.dgvcc_dmap_splat(
        _native.ptr(d_pts), _native.ptr(sigma), FIXED_SIGMA, ADAPTIVE_TRUNCATE if adaptive else FIXED_TRUNCATE, n,
        height, width, _native.ptr(ws), ws_bytes, _native.ptr(out), stream), "dgvcc_dmap_splat")
    return out


def _density(img, points, adaptive, device=None):
    height, width = int(img.shape[0]), int(img.shape[1])
    if len(points) == 0:  # dmap_gen.py:28-29
        return np.zeros((height, width), dtype=np.float32)
    dev = _device(device)
    pts = _check_points(points, height, width)
    host = torch.empty((height, width), dtype=torch.float32).pin_memory()
    host.copy_(_density_device(height, width, pts, adaptive, dev), non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host.numpy().copy()


def gaussian_filter_density(img, points, device=None):
    """Geometry-adaptive density map (reference: utils/dmap_gen.py:14-51)."""
    return _density(img, points, True, device)


def gaussian_filter_density_fixed(img, points, device=None):
    """Fixed-sigma density map, what ``run`` uses (reference: utils/dmap_gen.py:53-81)."""
    return _density(img, points, False, device)


def gaussian_filter_density_batch(shapes, points_list, fixed=False, device=None):
    """Many images back to back on one stream, one synchronisation at the end; returns a list of [H,W] arrays."""
    dev = _device(device)
    outs = []
    for (h, w), pts in zip(shapes, points_list):
        h, w = int(h), int(w)
        host = torch.empty((h, w), dtype=torch.float32).pin_memory()
        p = _check_points(pts, h, w) if len(pts) else np.zeros((0, 2))
        host.copy_(_density_device(h, w, p, not fixed, dev), non_blocking=True)
        outs.append(host)
    torch.cuda.current_stream(dev).synchronize()
    return [o.numpy() for o in outs]


def _image_shape(img_fn):
    from PIL import Image
    with Image.open(img_fn) as im:
        w, h = im.size
    return np.empty((h, w, 0))


def run(img_fn):
    """File protocol of dmap_gen.py:83-95: <name>.<ext> + <name>.npy -> <name>_dmap.npy, skipped when present."""
    img_ext = os.path.splitext(img_fn)[1]
    basename = os.path.basename(img_fn).replace(img_ext, '')
    gt_fn = img_fn.replace(img_ext, '.npy')
    dmap_fn = gt_fn.replace(basename, basename + '_dmap')

    if os.path.exists(dmap_fn):
        return

    img = _image_shape(img_fn)  # the reference decodes the image only for its shape
    gt = np.load(gt_fn)
    dmap = gaussian_filter_density_fixed(img, gt)
    np.save(dmap_fn, dmap)


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--path', type=str)
    args = parser.parse_args()

    path = args.path
    if not os.path.exists(path):
        raise Exception("Path does not exist")

    img_fns = []
    for phase in ['test']:
        img_fns += glob(os.path.join(path, phase, '*.jpg'))
    img_fns = [fn for fn in img_fns if 'aug' not in fn]

    # the reference fans out over Pool(8); one GPU stream is faster than that and needs no fork
    for fn in img_fns:
        run(fn)
