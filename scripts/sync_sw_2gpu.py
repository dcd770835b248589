"""SyncSwitchWhiten2d over NCCL, one rank per GPU: parity with the two-rank fixtures of the reference, equivalence with
the plain layer on the concatenated batch at a training-size shape, and timing (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
        scripts/sync_sw_2gpu.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import load_sw_cases  # noqa: E402
from dgvcc_b200.models.ISW import SwitchWhiten2d, SyncSwitchWhiten2d  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
failures = []


def worst(got, ref):
    got, ref = got.detach().double().cpu().numpy().reshape(np.shape(ref)), np.asarray(ref, np.float64)
    return float((np.abs(got - ref) / (np.abs(ref) + np.abs(ref).max())).max())


def load(m, c):
    with torch.no_grad():
        m.sw_mean_weight.copy_(c["mw"])
        if c["vw"] is not None:
            m.sw_var_weight.copy_(c["vw"])
        if c["weight"] is not None:
            m.weight.copy_(c["weight"])
            m.bias.copy_(c["bias"])
        m.running_mean.copy_(c["rmean"])
        m.running_cov.copy_(c["rcov"])


def grads(m, x):
    out = {"gx": x.grad, "gmw": m.sw_mean_weight.grad, "rmean": m.running_mean, "rcov": m.running_cov}
    if m.sw_var_weight is not None:
        out["gvw"] = m.sw_var_weight.grad
    if m.weight is not None:
        out["gweight"], out["gbias"] = m.weight.grad, m.bias.grad
    return out


# 1. the reference's own two-rank outputs (tests/golden/sw_cases.npz, kinds "sync2")
if world == 2:
    for name, c in load_sw_cases().items():
        if c["kind"] != "sync2":
            continue
        per = c["x"].shape[0] // 2
        m = SyncSwitchWhiten2d(c["x"].shape[1], num_pergroup=c["num_pergroup"], sw_type=c["sw_type"], tie_weight=c["tie"],
                               affine=c["affine"]).to(dev)
        load(m, c)
        x = c["x"][rank * per:(rank + 1) * per].to(dev).requires_grad_(True)
        y = m(x)
        y.backward(c["gy"][rank * per:(rank + 1) * per].to(dev))
        res = dict(grads(m, x), y=y)
        for k, ref in c["ref"][rank].items():
            e = worst(res[k], ref)
            if e > 2e-4:
                failures.append(f"{name} rank {rank} {k}: {e:.2e}")

# 2. sync over `world` ranks on slices == plain layer on the whole batch
shape, sw_type = (4 * world, 64, 80, 80), 5
g = torch.Generator().manual_seed(99)
full_x = torch.randn(*shape, generator=g) * 1.3 + 0.4
full_x = full_x + 0.4 * full_x.roll(1, dims=1)
full_gy = torch.randn(*shape, generator=g)
c = {"mw": torch.randn(sw_type, generator=g), "vw": torch.randn(sw_type, generator=g), "weight": 1 + 0.3 * torch.randn(64, generator=g),
     "bias": 0.3 * torch.randn(64, generator=g), "rmean": torch.zeros(4, 16, 1), "rcov": torch.zeros(4, 16, 16)}
plain, sync = SwitchWhiten2d(64, sw_type=sw_type).to(dev), SyncSwitchWhiten2d(64, sw_type=sw_type).to(dev)
load(plain, c)
load(sync, c)
xf = full_x.to(dev).requires_grad_(True)
yf = plain(xf)
yf.backward(full_gy.to(dev))
per = shape[0] // world
sl = slice(rank * per, (rank + 1) * per)
xs = full_x[sl].to(dev).requires_grad_(True)
ys = sync(xs)
ys.backward(full_gy[sl].to(dev))
pg, sg = grads(plain, xf), grads(sync, xs)
for k in ("gmw", "gvw", "gweight", "gbias"):          # parameter gradients are per-rank partial sums, like the reference's
    t = sg[k].clone()
    dist.all_reduce(t)
    sg[k] = t
checks = {"y": (ys, yf[sl]), "gx": (sg["gx"], pg["gx"][sl])}
checks.update({k: (sg[k], pg[k]) for k in ("gmw", "gvw", "gweight", "gbias", "rmean", "rcov")})
for k, (a, b) in checks.items():
    e = worst(a, b.detach().double().cpu().numpy())
    if e > 2e-5:
        failures.append(f"sync == plain, rank {rank} {k}: {e:.2e}")

# 3. timing: forward + backward of the synchronised layer on this rank's slice, max over ranks
gy = full_gy[sl].to(dev)
ts = []
for rep in range(12):
    xs.grad = None
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sync(xs).backward(gy)
    e1.record()
    torch.cuda.synchronize()
    if rep >= 2:
        ts.append(e0.elapsed_time(e1))
t = torch.tensor([min(ts)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
bad = torch.tensor([len(failures)], device=dev)
dist.all_reduce(bad)
for f in failures:
    print("MISMATCH", f, flush=True)
if rank == 0:
    print(json.dumps({"workload": f"SyncSwitchWhiten2d(64, sw_type={sw_type}) forward + backward, {per} x 64 x 80 x 80 per rank, NCCL",
                      "n_gpus": world, "ms_per_step_max_over_ranks": float(t), "mismatches": int(bad),
                      "checked": "two-rank reference fixtures (2e-4), sync on slices == plain on the whole batch (2e-5)"}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(bad) else 0)
