"""Drop-in for the reference's ``models/ISW/cov_settings.py`` (mask bookkeeping of the ISW loss) on B200.

Same classes and methods as the reference (models/ISW/cov_settings.py:16-107), so ``ISWCounter_*`` can keep
calling ``get_eye_matrix`` / ``set_variance_of_covariance`` / ``set_mask_matrix`` / ``get_mask_matrix`` /
``reset_mask_matrix`` (models/ISW/__init__.py:43-50, 93-116).  The statistics stay on the device; the average,
the top-k selection and the AND with the previous mask are one kernel (csrc/isw_kernels.cu:
isw_topk_mask_kernel).  ``relax_denom == 0`` needs the third-party ``kmeans1d`` exactly like the reference
(cov_settings.py:4,58); the shipped default is 2.0 (models/ISW/__init__.py:23).
"""
import torch

from ... import _native


def _cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("dgvcc_b200.models.ISW.cov_settings needs a CUDA device; there is no CPU path")
    return torch.device(device if device is not None else "cuda")


def make_cov_index_matrix(dim):  # cov_settings.py:7-13
    matrix = torch.LongTensor()
    s_index = 0
    for i in range(dim):
        matrix = torch.cat([matrix, torch.arange(s_index, s_index + dim).unsqueeze(0)], dim=0)
        s_index += (dim - (2 + i))
    return matrix.triu(diagonal=1).transpose(0, 1) + matrix.triu(diagonal=1)


def topk_mask(stats, count, k, prev_mask=None):
    """(values [n], mask [n]): values = sum(stats) / count, mask = 1 at the k largest values (AND prev_mask)."""
    stats = stats.contiguous()
    _native.require_cuda(stats, "topk_mask")
    n_stats, n = stats.shape
    values = torch.empty((n,), dtype=torch.float32, device=stats.device)
    mask = torch.empty((n,), dtype=torch.float32, device=stats.device)
    prev = prev_mask.contiguous().reshape(-1).to(torch.float32) if prev_mask is not None else None
    _native.check(_native.lib().dgvcc_isw_topk_mask(_native.ptr(stats), n_stats, int(count), n, int(k), _native.ptr(prev),
                                                    _native.ptr(values), _native.ptr(mask),
                                                    _native.stream_ptr(stats.device)), "dgvcc_isw_topk_mask")
    return values, mask


class CovMatrix_ISW:
    def __init__(self, dim, relax_denom=0, clusters=50, device=None):
        dev = _cuda(device)
        self.dim = dim
        self.i = torch.eye(dim, dim, device=dev)
        self.reversal_i = torch.ones(dim, dim, device=dev).triu(diagonal=1)
        self.num_off_diagonal = torch.sum(self.reversal_i)
        self.num_sensitive = 0
        self.var_matrix = None          # running fp32 sum of the statistics since the last set_mask_matrix
        self.count_var_cov = 0
        self.mask_matrix = None
        self.clusters = clusters
        if relax_denom == 0:    # kmeans1d clustering setting for ISW (cov_settings.py:35-38)
            self.margin = 0
        else:
            self.margin = self.num_off_diagonal // relax_denom

    def get_eye_matrix(self):
        return self.i, self.reversal_i

    def get_mask_matrix(self, mask=True):
        if self.mask_matrix is None:
            self.set_mask_matrix()
        return self.i, self.mask_matrix, 0, self.num_sensitive

    def reset_mask_matrix(self):
        self.mask_matrix = None

    def set_mask_matrix(self):
        stats = self.var_matrix.reshape(1, -1)  # the sum; the kernel divides by count (cov_settings.py:53)
        if self.margin == 0:    # cov_settings.py:57-61
            import kmeans1d  # third-party, like the reference; not part of this package
            var_flatten = stats.sum(0) / self.count_var_cov
            clusters, _ = kmeans1d.cluster(var_flatten.cpu(), self.clusters)
            num_sensitive = var_flatten.numel() - clusters.count(0)
        else:                   # cov_settings.py:62-65
            num_sensitive = self.num_off_diagonal - self.margin
        _, mask = topk_mask(stats, self.count_var_cov, int(num_sensitive), self.mask_matrix)
        self.mask_matrix = mask.view(self.dim, self.dim)
        self.num_sensitive = torch.sum(self.mask_matrix)
        self.var_matrix = None
        self.count_var_cov = 0

    def set_variance_of_covariance(self, var_cov):
        # one running sum like cov_settings.py:84-88 (same fp32 summation order), not a list that grows with every
        # validation patch
        var_cov = var_cov.detach().to(torch.float32)
        self.var_matrix = var_cov.clone() if self.var_matrix is None else self.var_matrix + var_cov
        self.count_var_cov += 1


class CovMatrix_IRW:
    def __init__(self, dim, relax_denom=0, device=None):
        dev = _cuda(device)
        self.dim = dim
        self.i = torch.eye(dim, dim, device=dev)
        self.reversal_i = torch.ones(dim, dim, device=dev).triu(diagonal=1)
        self.num_off_diagonal = torch.sum(self.reversal_i)
        if relax_denom == 0:
            self.margin = 0
        else:
            self.margin = self.num_off_diagonal // relax_denom

    def get_mask_matrix(self):
        return self.i, self.reversal_i, self.margin, self.num_off_diagonal
