"""InstanceWhitening + instance_whitening_loss forward + backward on the three BASELINE config-5 shapes, 3 steps each;
run under ncu (`-k regex:isw`) for a launch list, or plain for CUDA-event and wall-clock step times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss

dev = torch.device("cuda:0")
for (b, c, h, w) in [(8, 64, 160, 160), (8, 256, 80, 80), (8, 512, 40, 40)]:
    xin = torch.randn(b, c, h, w, device=dev, requires_grad=True)
    eye = torch.eye(c, device=dev)
    mask = torch.triu(torch.ones(c, c, device=dev), 1)
    num = mask.sum()
    iw = InstanceWhitening(c)
    ev, wall = [], []
    for rep in range(6):
        xin.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter()
        e0.record()
        y, wt = iw(xin)
        loss = instance_whitening_loss(wt, eye, mask, 0, num)
        loss.backward()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        ev.append(e0.elapsed_time(e1) * 1e3)
        wall.append((t1 - t0) * 1e6)
    print(f"C={c} HW={h*w}: step {min(ev):.0f} us on the device timeline, host issue time {min(wall):.0f} us", flush=True)
