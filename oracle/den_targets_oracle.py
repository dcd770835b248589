"""CPU oracle for the dataset-side density handling -- TEST INFRASTRUCTURE, not product code.

Restates the density part of /root/reference/datasets/den_cls_dataset.py:

* ``train_density``  <- DenClsDataset._train_transform (den_cls_dataset.py:109-150; the same block is in
  den_dataset.py:86-127): zero padding ``F.pad(dmap, (left, top, right, bottom))``, ``F.crop(dmap, i, j, h, w)``,
  sum-pool ``reshape([1, h/d, d, w/d, d]).sum(dim=(2, 4))`` and the horizontal flip;
* ``block_occupancy`` <- DenClsDataset.__getitem__ (den_cls_dataset.py:60-61): 16 x 16 block sums > 0.

Same torch-CPU primitives as the reference.  Pinned by tests/golden/den_cases.npz, produced by running the
unmodified class (tests/golden/make_golden.py: make_den).
"""
import torch


def train_density(dmap, left, top, i, j, h, w, downsample, flip):
    """dmap [H,W] f32 (numpy or tensor) -> [1, h/downsample, w/downsample] f32 tensor."""
    d = torch.as_tensor(dmap, dtype=torch.float32).unsqueeze(0)
    hh, ww = d.shape[1:]
    # get_padding (utils/misc.py:19-37) pads only up to the crop size, so right / bottom follow from it
    new_h, new_w = max(hh, h) if top or hh < h else hh, max(ww, w) if left or ww < w else ww
    d = torch.nn.functional.pad(d, (left, new_w - ww - left, top, new_h - hh - top))
    d = d[:, i:i + h, j:j + w]
    dh, dw = h // downsample, w // downsample
    d = d.reshape([1, dh, downsample, dw, downsample]).sum(dim=(2, 4))
    if flip:
        d = d.flip(-1)
    return d.float()


def block_occupancy(dmap, block=16):
    """[1, H', W'] -> [1, H'/block, W'/block] in {0., 1.} (den_cls_dataset.py:60-61)."""
    s = dmap.clone().reshape(1, dmap.shape[1] // block, block, dmap.shape[2] // block, block).sum(dim=(2, 4))
    return (s > 0).float()
