"""One Bayesian-loss batch over N real GPUs (torchrun, one process per GPU, CUDA IPC peer memory over NVLink).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/shard_bl_multi_gpu.py [--mode band|chunk] [--steps 20] [--chunks auto,256,512]

--mode chunk (point-chunk split, bl_sharded.py): the loss of every rank and the gathered gradient are BIT-IDENTICAL to
one GPU running the same chunk table.  --mode band (row bands of the grid, bl_banded.py): every rank holds the SAME
loss bits, the gradient is bit-identical to one GPU running the same chunk table and the loss within 1e-6 of it (the
counts are added band by band, then rank by rank).  Cases: the golden 'mixed' case with a small chunk size, BASELINE
config 2 and config 3; then config 3 is timed (per-step CUDA events, max over ranks).  One JSON line on rank 0.
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
    ap.add_argument("--mode", default="chunk", choices=["chunk", "band"])
    ap.add_argument("--tiles", default="0", help="band mode: comma-separated pixel tiles to time (0 = by task count, 82, 81, 41, 21)")
    ap.add_argument("--min-cell", type=int, default=0, help="smallest cell of the minima's point grid (dgvcc_bl_set_option)")
    ap.add_argument("--tail-splits", default="0", help="band mode: comma-separated tail_split values of build_meta to time")
    args = ap.parse_args()
    band = args.mode == "band"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dgvcc_b200 import synthetic
    from dgvcc_b200.losses import bl as blmod
    from dgvcc_b200.losses import bl_banded, bl_sharded
    from dgvcc_b200.losses.bl_banded import BandShardedBL, plan_bands
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
            pts=torch.cat([p.reshape(-1, 2) for p in pts]).to(dev),
            meta=torch.from_numpy(plan.meta if band else plan.meta_all()).to(dev),
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
        if band:
            plan = plan_bands([len(p) for p in pts], use_bg, world, None, hp, wp, chunk if name == "mixed" else None)
            mod = BandShardedBL(sigma, max(hp, wp) * stride, stride, bg_ratio, use_bg, dev, comm)
            mod.chunk = plan.chunk
        else:
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
            if band:   # the same bits on every rank; against one GPU only the order of the count sums differs
                everyone = [torch.zeros(1, device=dev) for _ in range(world)]
                dist.all_gather(everyone, loss.detach().reshape(1))
                mine = all(torch.equal(v, everyone[0]) for v in everyone)
                mine = mine and abs(float(loss) - float(ref_loss)) <= 1e-6 * abs(float(ref_loss))
            else:
                mine = torch.equal(loss.detach().reshape(()), ref_loss.reshape(()))
            if len(plan.owned[rank]):
                mine = mine and torch.equal(local_d.grad, ref_grad[plan.owned[rank]])
            flag = torch.tensor([int(mine)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            out["checks"][f"{name}_cull{int(cull)}"] = bool(flag.item())
            ok_all = ok_all and bool(flag.item())

    from dgvcc_b200 import _native
    if args.min_cell:
        _native.check(_native.lib().dgvcc_bl_set_option(_native.BL_OPT_MIN_CELL, args.min_cell), "dgvcc_bl_set_option")
    results = {}
    sweep = [(tile, forced, ts) for tile in ([int(t) for t in args.tiles.split(",")] if band else [0])
             for forced in ([None if c == "auto" else int(c) for c in args.chunks.split(",")] if args.chunks else [None])
             for ts in ([int(t) for t in args.tail_splits.split(",")] if band else [0])]
    for tile, forced, ts in sweep:
        bl_banded.TAIL_SPLIT = ts
        _native.check(_native.lib().dgvcc_bl_set_option(_native.BL_OPT_BAND_TILE, tile), "dgvcc_bl_set_option")
        bl_banded._plan_cache.clear()
        # ---- timing: config 3, dense
        blmod._CHUNK_POINTS = 1024
        bl_sharded.FORCE_CHUNK = bl_banded.FORCE_CHUNK = forced
        pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = case("config3")
        b, _, hp, wp = dens.shape
        if band:
            plan = plan_bands([len(p) for p in pts], use_bg, world, None, hp, wp)
            mod = BandShardedBL(sigma, 2048, stride, bg_ratio, use_bg, dev, comm)
        else:
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
            out = results.setdefault(f"{forced or 'auto'}" + (f"/tile{tile}" if tile else "") + (f"/tail{ts}" if ts else ""), {})
            out["phase_ms_rank0"] = {k: round(v, 4) for k, v in phases[0].items()}
            out["phase_ms_max_over_ranks"] = {k: round(max(p[k] for p in phases), 4) for k in phases[0]}
            out["phase_sum_ms_per_rank"] = [round(sum(p.values()), 4) for p in phases]
            out.update({"config3_ms_per_rank": allms, "config3_ms_per_step": max(allms), "images_per_s": b / (max(allms) * 1e-3),
                        "chunks": plan.total_chunks})
            if band:
                out.update({"chunk_points": plan.chunk, "bands": [[int(a), int(c)] for a, c in zip(plan.band_lo, plan.band_hi)],
                            "pixel_tile": [plan.layout.rows_per_thread, plan.layout.cols_per_thread]})
            else:
                out.update({"chunk_points": int(plan.c_cnt.max()), "groups": [len(g) for g in plan.groups]})
    if rank == 0:
        print(json.dumps({"world": world, "mode": args.mode, "checks": out_checks["checks"], "ok": ok_all, "timing": results}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
