"""Density targets of the reference's density-map datasets on B200.

    train_density_targets(dmaps, geoms, crop_size, downsample, with_bmap=True)
        <- the density part of DenClsDataset._train_transform (datasets/den_cls_dataset.py:109-150, identical
           in datasets/den_dataset.py:86-127): zero padding, crop, d x d sum-pool, horizontal flip
        <- the occupancy map of DenClsDataset.__getitem__ (den_cls_dataset.py:60-61): 16 x 16 block sums > 0

The random draws stay with the caller (``get_padding`` / ``random_crop`` / the flip coin of the dataset class):
each image comes with its geometry ``(pad_left, pad_top, crop_i, crop_j, flip)``.  A whole batch is one launch and
the result is already the stacked ``[B, 1, h/d, w/d]`` / ``[B, h/d/16, w/d/16]`` pair that ``DenClsDataset.collate``
(den_cls_dataset.py:17-24) builds.  Full-resolution maps may be device tensors (straight from
``dgvcc_b200.utils.dmap_gen``, no ``*_dmap.npy`` round trip) or numpy arrays.
"""
import numpy as np
import torch

from .. import _native

BLOCK = 16  # den_cls_dataset.py:60


def _dev(device):
    if not torch.cuda.is_available():
        raise RuntimeError("dgvcc_b200.datasets.den_targets needs a CUDA device; there is no CPU path")
    return torch.device(device if device is not None else "cuda")


def train_density_targets(dmaps, geoms, crop_size, downsample, with_bmap=True, device=None):
    """dmaps: list of [H_i, W_i] float32 maps (CUDA tensors or numpy); geoms: per image
    ``(pad_left, pad_top, crop_i, crop_j, flip)``; crop_size ``(h, w)`` like the dataset's ``crop_size``.
    Returns ``(dmap [B,1,h/d,w/d], bmap [B,h/d/16,w/d/16] or None)`` on the device."""
    if len(dmaps) == 0 or len(dmaps) != len(geoms):
        raise ValueError("need one geometry per density map")
    h, w = int(crop_size[0]), int(crop_size[1])
    d = int(downsample)
    if h % d or w % d:  # the reference's reshape raises on this (den_cls_dataset.py:138)
        raise RuntimeError(f"crop {h}x{w} is not a multiple of the down-sampling factor {d}")
    dh, dw = h // d, w // d
    if with_bmap and (dh % BLOCK or dw % BLOCK):
        raise RuntimeError(f"pooled map {dh}x{dw} is not a multiple of the {BLOCK}x{BLOCK} occupancy block")
    on_device = [isinstance(m, torch.Tensor) and m.is_cuda for m in dmaps]
    dev = dmaps[on_device.index(True)].device if any(on_device) else _dev(device)
    flat, meta, off = [], np.zeros((len(dmaps), _native.DEN_META_COLS), dtype=np.int64), 0
    for k, (m, g) in enumerate(zip(dmaps, geoms)):
        if m.ndim != 2:
            raise ValueError(f"density maps must be [H,W], got {tuple(m.shape)}")
        left, top, ci, cj, flip = (int(v) for v in g)
        hh, ww = int(m.shape[0]), int(m.shape[1])
        if min(left, top, ci, cj) < 0:
            raise ValueError(f"image {k}: negative padding / crop origin")
        meta[k] = (off, hh, ww, left, top, ci, cj, int(bool(flip)))
        if isinstance(m, torch.Tensor):
            if m.device != dev:
                _native.require_cuda(m, "train_density_targets")
            flat.append(m.detach().to(device=dev, dtype=torch.float32).reshape(-1))
        else:
            flat.append(torch.from_numpy(np.ascontiguousarray(m, dtype=np.float32).reshape(-1)).pin_memory()
                        .to(dev, non_blocking=True))
        off += hh * ww
    maps = flat[0] if len(flat) == 1 else torch.cat(flat)
    d_meta = torch.from_numpy(meta).pin_memory().to(dev, non_blocking=True)
    out = torch.empty((len(dmaps), 1, dh, dw), dtype=torch.float32, device=dev)
    bmap = torch.empty((len(dmaps), dh // BLOCK, dw // BLOCK), dtype=torch.float32, device=dev) if with_bmap else None
    _native.check(_native.lib().dgvcc_den_train_targets(_native.ptr(maps), _native.ptr(d_meta), len(dmaps), h, w, d,
                                                        _native.ptr(out), _native.ptr(bmap), _native.stream_ptr(dev)),
                  "dgvcc_den_train_targets")
    return out, bmap
