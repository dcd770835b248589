"""Scratch timing of the ISW kernels on BASELINE config-5 shapes (CUDA events, L2 flushed between reps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200 import _native
from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss

dev = torch.device("cuda:0")
lib = _native.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (b, c, h, w) in [(8, 64, 160, 160), (8, 256, 80, 80), (8, 512, 40, 40)]:
    hw = h * w
    x = torch.randn(b, c, hw, device=dev)
    eye = torch.eye(c, device=dev)
    n = lib.dgvcc_isw_workspace_bytes(b, c, hw)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    fc = torch.empty(b, c, c, device=dev)
    for tc in (0, 1):
        ts = []
        for rep in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            rc = lib.dgvcc_isw_covariance(_native.ptr(x), _native.ptr(eye), b, c, hw, tc, _native.ptr(ws), n, _native.ptr(fc), _native.stream_ptr(dev))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0
            if rep: ts.append(e0.elapsed_time(e1))
        t = min(ts)
        flops = 2.0 * b * c * c * hw
        print(f"C={c} HW={hw} tc={tc}: covariance {t*1e3:.1f} us  {flops/t/1e9:.1f} TFLOP/s(dense-equiv)  read {4*b*c*hw/t/1e6:.0f} GB/s")
    # torch reference on the same box
    xt = x.clone()
    for rep in range(3):
        flush.zero_(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); g = torch.bmm(xt, xt.transpose(1, 2)); e1.record(); torch.cuda.synchronize()
    print(f"   torch.bmm fp32: {e0.elapsed_time(e1)*1e3:.1f} us")
    # full module fwd+bwd
    xin = torch.randn(b, c, h, w, device=dev, requires_grad=True)
    mask = torch.triu(torch.ones(c, c, device=dev), 1)
    iw = InstanceWhitening(c)
    for rep in range(4):
        flush.zero_(); xin.grad = None
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); y, wt = iw(xin); loss = instance_whitening_loss(wt, eye, mask, 0, mask.sum()); loss.backward(); e1.record(); torch.cuda.synchronize()
    print(f"   IN + loss fwd+bwd (module): {e0.elapsed_time(e1)*1e3:.1f} us")
