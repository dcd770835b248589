"""Scratch timing of the fused BL on config 3 plus the MUFU/FFMA probes (not the bench contract)."""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200 import synthetic, _native
from dgvcc_b200.losses.bl import BL

dev = torch.device("cuda:0")
lib = _native.lib()
sink = torch.zeros(4, device=dev)
for name in ("dgvcc_probe_ex2", "dgvcc_probe_ffma"):
    fn = getattr(lib, name)
    ops = ctypes.c_int64(0)
    for _ in range(2):
        fn(_native.ptr(sink), 2000, ctypes.byref(ops), _native.stream_ptr(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); fn(_native.ptr(sink), 20000, ctypes.byref(ops), _native.stream_ptr(dev)); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ops.value / ms / 1e9:.2f} Tops/s  ({ms:.2f} ms)")

for cfg in (3, 2, 1):
    counts = synthetic.config_counts(cfg)
    w, h = synthetic.CONFIG_SHAPES[cfg]
    pts, tgt, dens, st = synthetic.bl_batch(cfg, counts, w, h)
    pts = [torch.from_numpy(p).to(dev) for p in pts]
    tgt = [torch.from_numpy(t).to(dev) for t in tgt]
    dens = torch.from_numpy(dens).to(dev).requires_grad_(True)
    st = torch.from_numpy(st).to(dev)
    mod = BL(8.0, max(w, h), 8, 1.0, True, dev)
    pairs = sum(counts) * (w // 8) * (h // 8)
    for it in range(3):
        dens.grad = None
        e = [torch.cuda.Event(True) for _ in range(3)]
        e[0].record(); loss = mod(pts, st, tgt, dens); e[1].record(); loss.backward(); e[2].record()
        torch.cuda.synchronize()
        f, b = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        print(f"cfg{cfg} B={len(counts)} pts={sum(counts)} fwd {f:.3f} ms bwd {b:.3f} ms  img/s {len(counts)/(f+b)*1e3:.0f}  "
              f"Gexp/s {3*pairs/(f+b)/1e6:.0f} loss {float(loss):.5f}")
