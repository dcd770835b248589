"""CPU oracle for the Bayesian-dataset target preparation -- TEST INFRASTRUCTURE, not product code.

Restates the two per-point steps of /root/reference/datasets/bay_dataset.py that feed the Bayesian loss
(SURVEY.md section 8f, rank 1):

* ``cal_dists``     <- BayesianDataset._cal_dists (bay_dataset.py:38-48): mean distance to the 3 nearest
  heads through the expansion sq_i - 2 p_i.p_j + sq_j, with its N < 4 special cases;
* ``crop_targets``  <- the crop block of _train_transform (bay_dataset.py:85-107): clipped box / crop
  overlap ratio, the >= 0.3 filter, shift into crop coordinates and the unconditional mirror.

Pinned by tests/golden/bay_cases.npz, produced by the unmodified reference class.
"""
import numpy as np


def cal_dists(pts):
    if len(pts) == 0:
        return np.array([[]])
    elif len(pts) == 1:
        return np.array([[4.0]])
    square = np.sum(pts * pts, axis=1)
    dists = np.sqrt(np.maximum(square[:, None] - 2 * np.matmul(pts, pts.T) + square[None, :], 0.0))
    if len(pts) < 4:
        return np.mean(dists[:, 1:], axis=1, keepdims=True)
    return np.mean(np.partition(dists, 3, axis=1)[:, 1:4], axis=1, keepdims=True)


def inner_area(c_left, c_up, c_right, c_down, bbox):
    """utils/misc.py:39-45."""
    il = np.maximum(c_left, bbox[:, 0])
    iu = np.maximum(c_up, bbox[:, 1])
    ir = np.minimum(c_right, bbox[:, 2])
    idn = np.minimum(c_down, bbox[:, 3])
    return np.maximum(ir - il, 0.0) * np.maximum(idn - iu, 0.0)


def crop_targets(gt, dists, i, j, h, w):
    """(kept points in crop coordinates, mirrored in x; their targets).  bay_dataset.py:85-107."""
    if len(gt) == 0:
        return gt, np.array([])
    nearest = np.clip(dists, 4.0, 128.0)
    bbox = np.concatenate((gt - nearest / 2.0, gt + nearest / 2.0), axis=1)
    area = inner_area(j, i, j + w, i + h, bbox)
    origin = np.squeeze(nearest * nearest, axis=-1)
    ratio = np.clip(1.0 * area / origin, 0.0, 1.0)
    mask = ratio >= 0.3
    targ = ratio[mask]
    out = gt[mask] - [j, i]
    if len(out) > 0:
        out[:, 0] = w - out[:, 0]
    return out, targ
