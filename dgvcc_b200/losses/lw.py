"""Drop-in for the reference's ``losses/lw.py`` (instance-whitening Gram loss) on B200.

    lw_loss(x, mask=None) -> 0-dim loss          (reference: losses/lw.py:5-18; imported in trainers/dgtrainer.py:24)

x [N, C, H, W], mask [N, 1, H, W] or None.  Per-(n, c) standardisation with the unbiased variance, optional spatial
mask, Gram on the tensor cores (csrc/isw_gram_tc.cu, the kernel of the ISW path), sum of the squared strictly-upper
entries; the backward is the ISW path's dX = S X GEMM followed by the standardisation backward.
"""
import os

import torch

from .. import _native


def _use_tc():
    return int(os.environ.get("DGVCC_ISW_TENSOR_CORES", "1"))


class _LwLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        _native.require_cuda(x, "lw_loss")
        n, c, h, w = x.shape
        hw = h * w
        dev = x.device
        lib = _native.lib()
        stream = _native.stream_ptr(dev)
        xc = x.detach().to(torch.float32).contiguous()
        m = None
        if mask is not None:
            m = mask.detach().to(device=dev, dtype=torch.float32).reshape(n, hw).contiguous()
        yhat = torch.empty((n, c, hw), dtype=torch.float32, device=dev)
        ym = torch.empty_like(yhat) if m is not None else None
        invstd = torch.empty((n * c,), dtype=torch.float32, device=dev)
        _native.check(lib.dgvcc_lw_standardize_forward(_native.ptr(xc), _native.ptr(m), n, c, hw, 1e-5, _native.ptr(yhat),
                                                       _native.ptr(ym), _native.ptr(invstd), stream),
                      "dgvcc_lw_standardize_forward")
        y = ym if m is not None else yhat
        nbytes = lib.dgvcc_isw_workspace_bytes(n, c, hw)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        gram = torch.empty((n, c, c), dtype=torch.float32, device=dev)
        _native.check(lib.dgvcc_isw_gram(_native.ptr(y), n, c, hw, _use_tc(), _native.ptr(ws), nbytes, _native.ptr(gram),
                                         stream), "dgvcc_isw_gram")
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        _native.check(lib.dgvcc_lw_loss_forward(_native.ptr(gram), n, c, hw, _native.ptr(ws), nbytes, _native.ptr(loss),
                                                stream), "dgvcc_lw_loss_forward")
        ctx.save_for_backward(yhat, y, invstd, gram, ws)
        ctx.mask = m
        ctx.meta = (x.shape, x.dtype, nbytes)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        yhat, y, invstd, gram, ws = ctx.saved_tensors
        shape, dtype, nbytes = ctx.meta
        n, c, hw = yhat.shape
        dev = yhat.device
        lib = _native.lib()
        stream = _native.stream_ptr(dev)
        g = grad_loss.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dy = torch.empty_like(yhat)
        _native.check(lib.dgvcc_lw_loss_backward(_native.ptr(y), _native.ptr(gram), _native.ptr(g), n, c, hw, _use_tc(),
                                                 _native.ptr(ws), nbytes, _native.ptr(dy), stream),
                      "dgvcc_lw_loss_backward")
        dx = torch.empty_like(yhat)
        _native.check(lib.dgvcc_lw_standardize_backward(_native.ptr(dy), _native.ptr(yhat), _native.ptr(invstd),
                                                        _native.ptr(ctx.mask), n, c, hw, _native.ptr(dx), stream),
                      "dgvcc_lw_standardize_backward")
        return dx.view(shape).to(dtype), None


def lw_loss(x, mask=None):
    # Instance Whitening Loss
    # x: (N, C, H, W)
    # mask: (N, 1, H, W)
    # return: scalar
    return _LwLoss.apply(x, mask)
