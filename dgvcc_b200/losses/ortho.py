"""Drop-in for the reference's ``losses/ortho.py`` (orthogonality loss between two feature banks) on B200.

    ortho_loss(x, y) -> 0-dim loss               (reference: losses/ortho.py:5-11; imported in trainers/dgtrainer.py:23)

x, y [C, P].  ``x y^T`` is the off-diagonal block of the Gram of the stacked operand ``[x; y]``, so the forward is
the ISW path's tensor-core Gram kernel; the two gradients are two dX = S X GEMMs of the same path.
"""
import os

import torch

from .. import _native


def _use_tc():
    return int(os.environ.get("DGVCC_ISW_TENSOR_CORES", "1"))


class _OrthoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        _native.require_cuda(x, "ortho_loss")
        _native.require_cuda(y, "ortho_loss")
        if x.dim() != 2 or x.shape != y.shape:
            raise ValueError(f"ortho_loss expects two [C, P] tensors, got {tuple(x.shape)} and {tuple(y.shape)}")
        c, p = x.shape
        dev = x.device
        lib = _native.lib()
        stream = _native.stream_ptr(dev)
        z = torch.cat([x.detach().to(torch.float32), y.detach().to(torch.float32)], dim=0).contiguous()  # [2C, P]
        nbytes = lib.dgvcc_isw_workspace_bytes(1, 2 * c, p)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        gram = torch.empty((2 * c, 2 * c), dtype=torch.float32, device=dev)
        _native.check(lib.dgvcc_isw_gram(_native.ptr(z), 1, 2 * c, p, _use_tc(), _native.ptr(ws), nbytes, _native.ptr(gram),
                                         stream), "dgvcc_isw_gram")
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        _native.check(lib.dgvcc_ortho_loss_forward(_native.ptr(gram), c, p, _native.ptr(ws), nbytes, _native.ptr(loss), stream),
                      "dgvcc_ortho_loss_forward")
        ctx.save_for_backward(z, gram, ws)
        ctx.meta = (c, p, nbytes, x.dtype, y.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        z, gram, ws = ctx.saved_tensors
        c, p, nbytes, xdt, ydt = ctx.meta
        dev = z.device
        g = grad_loss.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        gx = torch.empty((c, p), dtype=torch.float32, device=dev)
        gy = torch.empty((c, p), dtype=torch.float32, device=dev)
        _native.check(_native.lib().dgvcc_ortho_loss_backward(
            _native.ptr(z[:c]), _native.ptr(z[c:]), _native.ptr(gram), _native.ptr(g), c, p, _use_tc(), _native.ptr(ws),
            nbytes, _native.ptr(gx), _native.ptr(gy), _native.stream_ptr(dev)), "dgvcc_ortho_loss_backward")
        return gx.to(xdt), gy.to(ydt)


def ortho_loss(x, y):
    # x: (C, P)
    # y: (C, P)
    # return: scalar
    return _OrthoLoss.apply(x, y)
