"""CPU oracle for the ISW mask bookkeeping -- TEST INFRASTRUCTURE, not product code.

Restates CovMatrix_ISW of /root/reference/models/ISW/cov_settings.py:16-89 with the
same torch-CPU primitives (topk): accumulate the variance-of-covariance statistic, average it, keep the
``num_off_diagonal - margin`` most variable entries as the mask, AND with the previous mask.  Only the
``relax_denom != 0`` branch (the shipped default 2.0, models/ISW/__init__.py:23); the ``relax_denom == 0`` branch
needs the third-party ``kmeans1d``, which is absent here.
Pinned by tests/golden/cov_cases.npz, produced by the unmodified class.
"""
import torch


class CovMatrixISW:
    def __init__(self, dim, relax_denom):
        self.dim = dim
        self.i = torch.eye(dim, dim)
        self.reversal_i = torch.ones(dim, dim).triu(diagonal=1)
        self.num_off_diagonal = torch.sum(self.reversal_i)
        self.margin = self.num_off_diagonal // relax_denom            # cov_settings.py:39
        self.var_matrix, self.count_var_cov, self.mask_matrix, self.num_sensitive = None, 0, None, 0

    def set_variance_of_covariance(self, var_cov):                    # cov_settings.py:84-89
        self.var_matrix = var_cov if self.var_matrix is None else self.var_matrix + var_cov
        self.count_var_cov += 1

    def set_mask_matrix(self):                                        # cov_settings.py:52-82
        var_flatten = torch.flatten(self.var_matrix / self.count_var_cov)
        num_sensitive = self.num_off_diagonal - self.margin
        _, indices = torch.topk(var_flatten, k=int(num_sensitive))
        mask = torch.zeros(self.dim * self.dim)
        mask[indices] = 1
        mask = mask.view(self.dim, self.dim)
        self.mask_matrix = mask if self.mask_matrix is None else (self.mask_matrix.int() & mask.int()).float()
        self.num_sensitive = torch.sum(self.mask_matrix)
        self.var_matrix, self.count_var_cov = None, 0
