"""Point-chunk sharded Bayesian loss on N real GPUs (torchrun, one process per GPU, CUDA IPC peer memory over NVLink).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/shard_bl_multi_gpu.py [--steps 20]

Checks: the loss of every rank and the gathered gradient are BIT-IDENTICAL to one GPU running the same chunk table
(rank 0 computes that reference), for the golden 'mixed' case with a small chunk size, BASELINE config 2 and config 3;
then times config 3 (per-step CUDA events, max over ranks).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--chunks", default="", help="comma-separated chunk sizes to time on config 3 (default: the plan's own choice)")
    ap.add_argument("--no-checks", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dgvcc_b200 import synthetic
    from dgvcc_b200.losses import bl as blmod
    from dgvcc_b200.losses.bl_sharded import ChunkShardedBL, IpcComm, plan_shards
    from helpers import load_bl_golden

    comm = IpcComm(device=dev)
    out = out_checks = {"world": world, "checks": {}}

    def case(name):
        if name.startswith("config"):
            cfg = int(name[-1])
            w, h = synthetic.CONFIG_SHAPES[cfg]
            pts, tgt, dens, st = synthetic.bl_batch(cfg, synthetic.config_counts(cfg), w, h, 8)
            return ([torch.from_numpy(p) for p in pts], [torch.from_numpy(t) for t in tgt], torch.from_numpy(st),
                    torch.from_numpy(dens), 8, 8.0, 1.0, True)
        c = load_bl_golden(name)
        return c["points"], c["targets"], c["st_sizes"], c["density"], c["stride"], c["sigma"], c["bg_ratio"], c["use_bg"]

    def reference(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull):
        packed = types.SimpleNamespace(
            pts=torch.cat([p.reshape(-1, 2) for p in pts]).to(dev), meta=torch.from_numpy(plan.meta_all()).to(dev),
            total_rows=plan.total_rows, total_chunks=plan.total_chunks, multi_chunk=plan.multi_chunk, batch=plan.batch)
        tg = torch.cat([t.reshape(-1) for t in tgt]).to(dev)
        d = dens.to(dev).clone().requires_grad_(True)
        loss = blmod._FusedBL.apply(d, packed, tg, st.to(dev), float(stride), float(sigma), float(bg_ratio), bool(use_bg),
                                    1.0 / plan.batch, None, cull)
        loss.backward()
        return loss.detach(), d.grad

    ok_all = True
    for name, chunk in (() if args.no_checks else (("mixed", 29), ("config2", 1024), ("config3", 1024))):
        blmod._CHUNK_POINTS = chunk   # upper bound; the sharded plan picks smaller chunks for larger worlds
        pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = case(name)
        b, _, hp, wp = dens.shape
        plan = plan_shards([len(p) for p in pts], use_bg, world, None, hp, wp)
        mod = ChunkShardedBL(sigma, max(hp, wp) * stride, stride, bg_ratio, use_bg, dev, comm)
        for cull in (False, True):
            mod.exact_cull = cull
            mod.defer_loss = cull
            local_d = dens[plan.owned[rank]].to(dev).clone().requires_grad_(True)
            for _ in range(3):  # several steps: flags, epochs and buffer re-use
                local_d.grad = None
                loss = mod(pts, st.to(dev), tgt, local_d)
                loss.backward()
            mod.check()
            ref_loss, ref_grad = reference(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull)
            mine = torch.equal(loss.detach().reshape(()), ref_loss.reshape(()))
            if len(plan.owned[rank]):
                mine = mine and torch.equal(local_d.grad, ref_grad[plan.owned[rank]])
            flag = torch.tensor([int(mine)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            out["checks"][f"{name}_cull{int(cull)}"] = bool(flag.item())
            ok_all = ok_all and bool(flag.item())

    results = {}
    for forced in ([None if c == "auto" else int(c) for c in args.chunks.split(",")] if args.chunks else [None]):
        # ---- timing: config 3, dense
        blmod._CHUNK_POINTS = 1024
        from dgvcc_b200.losses import bl_sharded
        bl_sharded.FORCE_CHUNK = forced
        pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = case("config3")
        b, _, hp, wp = dens.shape
        plan = plan_shards([len(p) for p in pts], use_bg, world, None, hp, wp)
        mod = ChunkShardedBL(sigma, 2048, stride, bg_ratio, use_bg, dev, comm)
        mod.exact_cull = False
        mod.defer_loss = True   # a training step: loss.backward() first, the value is read afterwards
        local_d = dens[plan.owned[rank]].to(dev).clone().requires_grad_(True)
        st_d = st.to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def step():
            local_d.grad = None
            loss = mod(pts, st_d, tgt, local_d)
            loss.backward()

        for _ in range(5):
            step()
            flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        evs = []
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            evs.append((e0, e1))
            flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        mod.check()
        ms = sum(a.elapsed_time(c) for a, c in evs) / args.steps
        allms = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([ms], device=dev, dtype=torch.float64))
        allms = [float(x[0]) for x in allms]
        # per-phase device times (CUDA events recorded between the launches), averaged over a few steps
        mod.profile = True
        acc = {}
        for _ in range(5):
            step()
            for k, v in mod.phase_ms().items():
                acc[k] = acc.get(k, 0.0) + v / 5
            flush.zero_()
        mod.profile = False
        phases = [None] * world
        dist.all_gather_object(phases, acc)
        if rank == 0:
            out = results.setdefault(str(forced or "auto"), {})
            out["phase_ms_rank0"] = {k: round(v, 4) for k, v in phases[0].items()}
            out["phase_ms_max_over_ranks"] = {k: round(max(p[k] for p in phases), 4) for k in phases[0]}
            out["phase_sum_ms_per_rank"] = [round(sum(p.values()), 4) for p in phases]
            out.update({"config3_ms_per_rank": allms, "config3_ms_per_step": max(allms), "images_per_s": b / (max(allms) * 1e-3),
                        "chunks": plan.total_chunks, "chunk_points": int(plan.c_cnt.max()), "groups": [len(g) for g in plan.groups]})
    if rank == 0:
        print(json.dumps({"world": world, "checks": out_checks["checks"], "ok": ok_all, "timing": results}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
