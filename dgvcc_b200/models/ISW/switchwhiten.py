"""Drop-in for the reference's ``models/ISW/switchwhiten.py`` (and its copy ``models/SW/ops/switchwhiten.py``) on B200.

    SwitchWhiten2d(num_features, num_pergroup=16, sw_type=2, T=5, tie_weight=False, eps=1e-5, momentum=0.99, affine=True)

Same parameters (``sw_mean_weight``, ``sw_var_weight``, ``weight``, ``bias``), buffers (``running_mean`` [g, c, 1],
``running_cov`` [g, c, c]) and ``state_dict`` keys as the reference class (switchwhiten.py:22-77), so checkpoints load
either way.  ``forward`` (switchwhiten.py:84-183) runs on the kernels of csrc/sw_kernels.cu: one pass over ``x`` for
the per-sample moments, the batch / layer statistics derived from them, Newton's iteration in fp64 on the c x c
matrices, one pass for ``y``; the backward is hand-derived (closed-form adjoint of the iteration) and costs one more
pass over ``x`` and ``grad_y`` for the moments and one for ``grad_x``.  CUDA only.
"""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from ... import _native


class _Exchange:
    """The four all-reduces of SyncMeanCov (sync_switchwhiten.py:21,25,44,45); ``None`` in the plain layer."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world_size = dist.get_world_size(group)

    def sum_(self, t):
        self.dist.all_reduce(t, group=self.group)
        return t


class _SwitchWhitenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sw_mean_weight, sw_var_weight, weight, bias, running_mean, running_cov, cfg):
        _native.require_cuda(x, "SwitchWhiten2d")
        cper, sw_type, T, eps, momentum, training, exchange, sync = cfg
        n, ch, h, w = x.shape
        hw, groups, dev = h * w, ch // cper, x.device
        lib, stream = _native.lib(), _native.stream_ptr(dev)
        f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()
        xc, mw, vw, wt, bs = f32(x), f32(sw_mean_weight), f32(sw_var_weight), f32(weight), f32(bias)
        nbytes = lib.dgvcc_sw_workspace_bytes(n, ch, hw, cper)
        if nbytes == 0:
            raise ValueError(f"SwitchWhiten2d: unsupported shape {tuple(x.shape)} with num_pergroup={cper} "
                             "(groups of 4, 8 or 16 channels)")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
        mean_in, cov_in = f64(n, ch), f64(n, groups, cper, cper)
        _native.check(lib.dgvcc_sw_instance_stats(_native.ptr(xc), n, ch, hw, cper, _native.ptr(mean_in), _native.ptr(cov_in),
                                                  _native.ptr(ws), nbytes, stream), "dgvcc_sw_instance_stats")
        world = 1
        if training:
            mean_bn, cov_bn = f64(ch), f64(groups, cper, cper)
            _native.check(lib.dgvcc_sw_batch_mean(_native.ptr(mean_in), n, ch, _native.ptr(mean_bn), stream),
                          "dgvcc_sw_batch_mean")
            if exchange is not None:
                world = exchange.world_size
                exchange.sum_(mean_bn).div_(world)
            _native.check(lib.dgvcc_sw_batch_cov(_native.ptr(mean_in), _native.ptr(cov_in), _native.ptr(mean_bn), n, ch, cper,
                                                 _native.ptr(cov_bn), stream), "dgvcc_sw_batch_cov")
            if exchange is not None:
                exchange.sum_(cov_bn).div_(world)
            plain = all(b.dtype == torch.float32 and b.is_contiguous() and b.device == dev for b in (running_mean, running_cov))
            if plain:    # switchwhiten.py:101-104, one launch
                _native.check(lib.dgvcc_sw_update_running(_native.ptr(running_mean), _native.ptr(running_cov),
                                                          _native.ptr(mean_bn), _native.ptr(cov_bn), ch, cper, momentum,
                                                          1 - momentum, stream), "dgvcc_sw_update_running")
            else:
                with torch.no_grad():
                    running_mean.mul_(momentum)
                    running_mean.add_((1 - momentum) * mean_bn.to(running_mean).view_as(running_mean))
                    running_cov.mul_(momentum)
                    running_cov.add_((1 - momentum) * cov_bn.to(running_cov))
        else:
            mean_bn = running_mean.detach().to(device=dev, dtype=torch.float64).reshape(ch).contiguous()
            cov_bn = running_cov.detach().to(device=dev, dtype=torch.float64).contiguous()
        a_fwd = torch.empty((n, groups, cper, cper), dtype=torch.float32, device=dev)
        y = torch.empty_like(xc)
        _native.check(lib.dgvcc_sw_whiten_forward(_native.ptr(xc), _native.ptr(mean_in), _native.ptr(cov_in),
                                                  _native.ptr(mean_bn), _native.ptr(cov_bn), _native.ptr(mw), _native.ptr(vw),
                                                  _native.ptr(wt), _native.ptr(bs), n, ch, hw, cper, sw_type, T, eps,
                                                  _native.ptr(a_fwd), _native.ptr(y), _native.ptr(ws), nbytes, stream),
                      "dgvcc_sw_whiten_forward")
        ctx.save_for_backward(xc, mean_in, cov_in, mean_bn, cov_bn, mw, vw, wt, a_fwd)
        if not training:
            bn_scale = 1.0 / (n * hw) if sync else 0.0   # sync_switchwhiten.py:48-55, see the C header
        else:
            bn_scale = 1.0 / (n * hw * world)
        ctx.cfg = (cper, sw_type, T, eps, training, exchange, bn_scale, x.dtype, nbytes)
        ctx.dtypes = tuple(None if t is None else t.dtype for t in (sw_mean_weight, sw_var_weight, weight, bias))
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_y):
        xc, mean_in, cov_in, mean_bn, cov_bn, mw, vw, wt, a_fwd = ctx.saved_tensors
        cper, sw_type, T, eps, training, exchange, bn_scale, x_dtype, nbytes = ctx.cfg
        n, ch, h, w = xc.shape
        hw, groups, dev = h * w, ch // cper, xc.device
        lib, stream = _native.lib(), _native.stream_ptr(dev)
        gy = grad_y.detach().to(device=dev, dtype=torch.float32).contiguous()
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        g_mw = torch.empty_like(mw)
        g_vw = None if vw is None else torch.empty_like(vw)
        g_wt = None if wt is None else torch.empty_like(wt)
        g_bs = None if wt is None else torch.empty_like(wt)
        g_mean_bn = torch.empty((ch,), dtype=torch.float64, device=dev)
        g_cov_bn = torch.empty((groups, cper, cper), dtype=torch.float64, device=dev)
        _native.check(lib.dgvcc_sw_backward_stats(_native.ptr(xc), _native.ptr(gy), _native.ptr(mean_in), _native.ptr(cov_in),
                                                  _native.ptr(mean_bn), _native.ptr(cov_bn), _native.ptr(mw), _native.ptr(vw),
                                                  _native.ptr(wt), n, ch, hw, cper, sw_type, T, eps, _native.ptr(g_mw),
                                                  _native.ptr(g_vw), _native.ptr(g_wt), _native.ptr(g_bs),
                                                  _native.ptr(g_mean_bn), _native.ptr(g_cov_bn), _native.ptr(ws), nbytes,
                                                  stream), "dgvcc_sw_backward_stats")
        if training and exchange is not None:
            exchange.sum_(g_mean_bn)
            exchange.sum_(g_cov_bn)
        gx = torch.empty_like(xc)
        _native.check(lib.dgvcc_sw_backward_apply(_native.ptr(xc), _native.ptr(gy), _native.ptr(a_fwd), _native.ptr(mean_in),
                                                  _native.ptr(mean_bn), _native.ptr(g_mean_bn), _native.ptr(g_cov_bn),
                                                  _native.ptr(mw), _native.ptr(vw), bn_scale, n, ch, hw, cper, sw_type,
                                                  _native.ptr(gx), _native.ptr(ws), nbytes, stream),
                      "dgvcc_sw_backward_apply")
        cast = lambda g, dt: None if g is None else g.to(dt)
        d_mw, d_vw, d_wt, d_bs = ctx.dtypes
        return gx.to(x_dtype), cast(g_mw, d_mw), cast(g_vw, d_vw), cast(g_wt, d_wt), cast(g_bs, d_bs), None, None, None


class SwitchWhiten2d(nn.Module):
    """Switchable whitening (BW + IW [+ LN [+ BN + IN]]); constructor and state of switchwhiten.py:7-77."""

    _sw_types = (2, 3, 5)
    _sync = False

    def __init__(self, num_features, num_pergroup=16, sw_type=2, T=5, tie_weight=False, eps=1e-5, momentum=0.99,
                 affine=True):
        super().__init__()
        if sw_type not in self._sw_types:
            raise ValueError('sw_type should be in {}, but got {}'.format(list(self._sw_types), sw_type))
        assert num_features % num_pergroup == 0
        self.num_features, self.num_pergroup, self.num_groups = num_features, num_pergroup, num_features // num_pergroup
        self.sw_type, self.T, self.tie_weight, self.eps, self.momentum, self.affine = sw_type, T, tie_weight, eps, momentum, affine
        self.sw_mean_weight = Parameter(torch.ones(sw_type))
        if tie_weight:
            self.register_parameter('sw_var_weight', None)
        else:
            self.sw_var_weight = Parameter(torch.ones(sw_type))
        if affine:
            self.weight = Parameter(torch.ones(num_features))
            self.bias = Parameter(torch.zeros(num_features))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)
        # both buffers start at zero: reset_parameters overrides the identity the constructor registers
        # (switchwhiten.py:61-71)
        self.register_buffer('running_mean', torch.zeros(self.num_groups, num_pergroup, 1))
        self.register_buffer('running_cov', torch.zeros(self.num_groups, num_pergroup, num_pergroup))

    def reset_parameters(self):
        with torch.no_grad():
            self.running_mean.zero_()
            self.running_cov.zero_()
            self.sw_mean_weight.fill_(1.0)
            if self.sw_var_weight is not None:
                self.sw_var_weight.fill_(1.0)
            if self.affine:
                self.weight.fill_(1.0)
                self.bias.zero_()

    def extra_repr(self):
        return ('{num_features}, num_pergroup={num_pergroup}, sw_type={sw_type}, T={T}, tie_weight={tie_weight}, '
                'eps={eps}, momentum={momentum}, affine={affine}'.format(**self.__dict__))

    def _exchange(self):
        return None

    def forward(self, x):
        if self.sw_type not in (2, 3, 5):
            # the synchronised reference class accepts sw_type 4 but has no branch for it (sync_switchwhiten.py:176-195)
            raise RuntimeError(f"sw_type={self.sw_type} has no forward in the reference either")
        cfg = (self.num_pergroup, self.sw_type, self.T, self.eps, self.momentum, self.training, self._exchange(), self._sync)
        return _SwitchWhitenFn.apply(x, self.sw_mean_weight, self.sw_var_weight, self.weight, self.bias, self.running_mean,
                                     self.running_cov, cfg)
