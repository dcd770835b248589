"""CPU oracle for the ISW instance-whitening covariance loss -- TEST INFRASTRUCTURE, not product code.

Restates /root/reference/models/ISW/instance_whitening.py:5-39 with torch-CPU ops (fp32, or fp64
for an accuracy yardstick).  Pinned by tests/golden/isw_cases.npz, produced by the unmodified
reference file loaded by path (its package __init__ needs the absent ``kmeans1d``).
"""
import torch


def instance_standardize(x, eps=1e-5):
    """nn.InstanceNorm2d(dim, affine=False): biased variance over H*W per (b, c).  instance_whitening.py:9-12."""
    return torch.nn.functional.instance_norm(x, eps=eps)


def covariance(f_map, eye):
    """bmm(X, X^T) / (HW - 1) + 1e-5 * eye  ->  ([B,C,C], B).  instance_whitening.py:30-39."""
    b, c, h, w = f_map.shape
    x = f_map.contiguous().view(b, c, -1)
    return torch.bmm(x, x.transpose(1, 2)).div(h * w - 1) + (1e-5 * eye), b


def whitening_loss(f_map, eye, mask_matrix, margin, num_remove_cov):
    """sum_b clamp((sum |f_cor * mask| - margin) / num_remove_cov, min=0) / B.  instance_whitening.py:19-27."""
    f_cor, b = covariance(f_map, eye)
    off = torch.sum(torch.abs(f_cor * mask_matrix), dim=(1, 2), keepdim=True) - margin
    return torch.sum(torch.clamp(torch.div(off, num_remove_cov), min=0)) / b


def upper_mask(c, keep_fraction, seed):
    """Strictly-upper-triangular 0/1 mask keeping a random fraction of the entries
    (the relax_denom branch of cov_settings.py:63-73 selects a subset of triu(1))."""
    g = torch.Generator().manual_seed(seed)
    tri = torch.triu(torch.ones(c, c), diagonal=1)
    return tri * (torch.rand(c, c, generator=g) < keep_fraction).float()


def covstat_variance(f_map, eye, reverse_eye):
    """cal_covstat, models/ISW/__init__.py:93-104: var over the batch of the masked covariance.
    (The enclosing reference module cannot be imported here -- it needs kmeans1d -- so this restatement is
    pinned only to the same torch ops, not to a reference-executed fixture.)"""
    f_cor, _ = covariance(f_map, eye)
    return torch.var(f_cor * reverse_eye, dim=0)
