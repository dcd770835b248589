"""How does the tensor-core Gram's error grow with the length of the in-TMEM accumulation chain?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200 import _native
DEV = "cuda:0"
b, c = 2, 128
for hw, kps in [(256, 256), (1024, 1024), (4096, 4096), (16384, 16384), (16384, 1024), (16384, 256), (65536, 65536), (65536, 1024)]:
    g = torch.Generator().manual_seed(hw)
    x = torch.randn((b, c, hw), generator=g)
    xd = x.to(DEV)
    splits = (hw + kps - 1) // kps
    part = torch.zeros((b, splits, 1, 128, 128), device=DEV)  # c = 128: one 128x128 tile
    rc = _native.lib().dgvcc_isw_gram_tc_partials(_native.ptr(xd), b, c, hw, splits, kps, _native.ptr(part), _native.stream_ptr(torch.device(DEV)))
    assert rc == 0
    torch.cuda.synchronize()
    got = part.sum(1)[:, 0].cpu().double()
    ref = torch.bmm(x.double(), x.double().transpose(1, 2))
    ref32 = torch.bmm(xd, xd.transpose(1, 2)).cpu().double()
    d = torch.diagonal(got - ref, dim1=1, dim2=2)
    print(f"hw={hw:6d} kps={kps:6d}: max|err|/max|G| ours {((got-ref).abs().max()/ref.abs().max()).item():.2e} (diag mean signed {d.mean().item()/hw:+.2e} rel)  cuBLAS fp32 {((ref32-ref).abs().max()/ref.abs().max()).item():.2e}")
