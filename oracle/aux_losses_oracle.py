"""CPU oracle for the auxiliary Gram-type losses -- TEST INFRASTRUCTURE, not product code.

Restates /root/reference/losses/lw.py:5-18 (``lw_loss``) and /root/reference/losses/ortho.py:5-11
(``ortho_loss``) with the same torch-CPU ops.  Pinned by tests/golden/aux_cases.npz,
produced by the unmodified files loaded by path (tests/golden/make_golden.py: make_aux).
"""
import torch


def lw_loss(x, mask=None):
    """Per-(n, c) standardisation with the unbiased variance, optional spatial mask, Gram, sum of squared
    strictly-upper-triangular entries.  lw.py:10-18."""
    n, c, h, w = x.shape
    x = x.view(n, c, -1)
    x = x - torch.mean(x, dim=2, keepdim=True)
    x = x / torch.sqrt(torch.var(x, dim=2, keepdim=True) + 1e-5)
    if mask is not None:
        x = x * mask.view(n, 1, -1)
    gram = torch.matmul(x, x.transpose(1, 2))
    return torch.sum(torch.square(torch.triu(gram, diagonal=1)))


def ortho_loss(x, y):
    """mean over all C*C entries of triu(x y^T, 1)^2.  ortho.py:9-11."""
    gram = torch.matmul(x, y.t())
    return torch.mean(torch.square(torch.triu(gram, diagonal=1)))
