"""ISW covariance loss with the batch split over N GPUs (BASELINE config 5: "B=8 ... 8xB200").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/isw_multi_gpu.py [--steps 20]

Samples are independent units of the loss (instance_whitening.py:23-25: a mean over the batch), so rank r evaluates
its own samples and ``dgvcc_b200.sharding.sharded_mean_over_batch`` all-reduces the scalar (NCCL); gradients stay
local.  Checks, for the three whitened layers of the sta_final.yml-shaped step ((B,C,H,W) = (8,64,160,160),
(8,256,80,80), (8,512,40,40)): the sharded loss equals the whole-batch loss on one GPU and every rank's gradient equals
its slice of the whole-batch gradient, rtol 1e-6.  Then times InstanceWhitening + loss forward + backward per layer
(CUDA events, max over ranks).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dgvcc_b200 import synthetic
    from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss
    from dgvcc_b200.sharding import sharded_mean_over_batch

    out = {"world": world, "layers": {}}
    ok_all = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (b, c, h, w) in synthetic.CONFIG5_SHAPES:
        gen = torch.Generator().manual_seed(5000 + c)
        x_all = torch.randn(b, c, h, w, generator=gen)
        mask = (torch.rand(c, c, generator=gen) < 0.5).float().triu(1).to(dev)
        eye = torch.eye(c, device=dev)
        nrm = mask.sum()
        per = (b + world - 1) // world
        lo, hi = min(b, rank * per), min(b, (rank + 1) * per)
        iw = InstanceWhitening(c)

        def step(x):
            x.grad = None
            _, wt = iw(x)
            loss = instance_whitening_loss(wt, eye, mask, 0, nrm)
            return loss

        # whole batch on this GPU (the reference value), then this rank's samples
        xw = x_all.to(dev).requires_grad_(True)
        lw = step(xw)
        lw.backward()
        line = {"samples_per_rank": per}
        if hi > lo:
            xs = x_all[lo:hi].to(dev).requires_grad_(True)
            ls = sharded_mean_over_batch(step(xs), hi - lo, b)
            ls.backward()
            err_l = abs(float(ls) - float(lw)) / abs(float(lw))
            ref_g = xw.grad[lo:hi]
            err_g = float((xs.grad - ref_g).abs().max() / ref_g.abs().max())
        else:  # more ranks than samples: take part in the all-reduce with a zero contribution
            ls = sharded_mean_over_batch(torch.zeros((), device=dev), 0, b)
            err_l, err_g = abs(float(ls) - float(lw)) / abs(float(lw)), 0.0
        worst = torch.tensor([err_l, err_g], device=dev, dtype=torch.float64)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        line["loss_rel_err"], line["grad_rel_err_to_max"] = float(worst[0]), float(worst[1])
        line["ok"] = bool(worst[0] <= 1e-6 and worst[1] <= 1e-6)
        ok_all = ok_all and line["ok"]

        def timed(fn, reps):
            for _ in range(3):
                fn()
                flush.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            evs = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                evs.append((e0, e1))
                flush.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            ms = torch.tensor([sum(a.elapsed_time(z) for a, z in evs) / reps], device=dev, dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms[0])

        def sharded_step():
            if hi > lo:
                sharded_mean_over_batch(step(xs), hi - lo, b).backward()
            else:
                sharded_mean_over_batch(torch.zeros((), device=dev), 0, b)

        line["ms_whole_batch_one_gpu"] = timed(lambda: step(xw).backward(), args.steps)
        line["ms_sharded"] = timed(sharded_step, args.steps)
        line["steps_per_s_sharded"] = 1e3 / line["ms_sharded"]
        out["layers"][f"B{b}_C{c}_HW{h * w}"] = line
    out["ok"] = ok_all
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
