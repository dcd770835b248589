"""Secondary measurements of SURVEY.md section 8d: density-map generation (BASELINE config 4) and the ISW
covariance loss (config 5).  One JSON line per workload; CUDA-event timing, L2 flushed between repetitions.

    python scripts/bench_aux.py [--cpu]      # --cpu also times the oracle (reference algorithm) on a bounded sample
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dgvcc_b200 import _native, synthetic  # noqa: E402

dev = torch.device("cuda:0")
HBM_GBS = 6446.0  # MEASURED_PEAKS.json (driver-written) copy bandwidth of this pool's B200
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    HBM_GBS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", HBM_GBS)


def ev():
    return torch.cuda.Event(enable_timing=True)


config4_images = synthetic.config4_images


def probe_tf32_peak(device=None, iters=20000):
    """TF32 flop/s of the tensor pipe (tcgen05.mma.kind::tf32 issued back to back on every SM), best of 3."""
    import ctypes
    d = device or dev
    lib = _native.lib()
    sink = torch.zeros(4, device=d)
    flops = ctypes.c_int64(0)
    best = 0.0
    for rep in range(4):
        e0, e1 = ev(), ev()
        e0.record()
        _native.check(lib.dgvcc_probe_tf32(_native.ptr(sink), iters, ctypes.byref(flops), _native.stream_ptr(d)), "dgvcc_probe_tf32")
        e1.record()
        torch.cuda.synchronize(d)
        if rep:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3))
    return best


def _reference_module(name):
    """The unmodified reference file from oracle/_ref (oracle/make_ref.sh) or None when it did not travel."""
    try:
        from oracle import ref_loader
        return ref_loader.load(name) if ref_loader.available(name) else None
    except Exception:
        return None


def dmap_lines(cpu, cpu_seconds=10.0):
    """BASELINE config 4 (64 JHU-shaped images): {splat_adaptive, splat_fixed} lines with roofline / e2e / cpu_baseline."""
    import ctypes
    from dgvcc_b200.utils import dmap_gen
    lib = _native.lib()
    images = config4_images()
    shapes = [s for s, _ in images]
    order = np.argsort([-len(p) for _, p in images], kind="stable")  # launch order of the public batch API
    pts_list = [np.ascontiguousarray(images[i][1], dtype=np.float64) for i in order]
    plan = dmap_gen._Plan([shapes[i] for i in order], [len(p) for p in pts_list])
    pl = plan.plan
    meta = torch.from_numpy(plan.meta).to(dev)
    d_pts = torch.from_numpy(np.concatenate([p for p in pts_list if len(p)], axis=0)).to(dev)
    out = torch.empty((pl.total_pixels,), dtype=torch.float32, device=dev)
    ws = torch.empty((pl.splat_workspace_bytes,), dtype=torch.uint8, device=dev)
    sig = torch.empty((pl.total_heads,), dtype=torch.float64, device=dev)
    kws = torch.empty((pl.knn_workspace_bytes,), dtype=torch.uint8, device=dev)
    bytes_alg = 4 * int(pl.total_pixels) + 16 * int(pl.total_heads)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = _native.stream_ptr(dev)
    lines = {}
    for adaptive in (True, False):
        t_dev, t_splat = [], []
        for rep in range(8):  # the whole set is ONE batched launch set (inputs resident, L2 flushed before each)
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            if adaptive:
                _native.check(lib.dgvcc_dmap_knn_sigma_batch(_native.ptr(d_pts), len(images), _native.ptr(meta), ctypes.byref(pl),
                                                             None, None, _native.ptr(sig), _native.ptr(kws),
                                                             pl.knn_workspace_bytes, st), "knn")
            e1.record()
            _native.check(lib.dgvcc_dmap_splat_batch(_native.ptr(d_pts), _native.ptr(sig) if adaptive else None, 4.0,
                                                     4.0 if adaptive else 1.75, len(images), _native.ptr(meta), ctypes.byref(pl),
                                                     _native.ptr(ws), pl.splat_workspace_bytes, _native.ptr(out), st), "splat")
            e2.record()
            torch.cuda.synchronize()
            if rep >= 3:
                t_dev.append(e0.elapsed_time(e2))
                t_splat.append(e1.elapsed_time(e2))
        t_dev, t_splat = float(np.mean(t_dev)), float(np.mean(t_splat))
        # end to end through the public API: host numpy in, host numpy out (H2D of heads + plan, D2H of the maps inside)
        fn = dmap_gen.gaussian_filter_density if adaptive else dmap_gen.gaussian_filter_density_fixed
        for (h, w), pts in images[:8]:
            fn(np.empty((h, w, 0)), pts)  # warm the pinned-block cache
        t0 = time.perf_counter()
        for (h, w), pts in images:
            fn(np.empty((h, w, 0)), pts)
        e2e_s = time.perf_counter() - t0
        pts_in = [p for _, p in images]
        for _ in range(2):
            maps = dmap_gen.gaussian_filter_density_batch(shapes, pts_in, fixed=not adaptive)
            del maps
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            maps = dmap_gen.gaussian_filter_density_batch(shapes, pts_in, fixed=not adaptive)
            times.append(time.perf_counter() - t0)
            del maps
        e2e_batch_s = float(np.mean(times))
        achieved = bytes_alg / (t_splat * 1e-3) / 1e9
        line = {
            "workload": f"BASELINE config 4: 64 JHU-shaped images (512..2048 px sides, 0..25000 heads, {int(pl.total_heads)} in all, "
                        f"{int(pl.total_pixels) * 4 / 1e6:.0f} MB of maps), {'adaptive kNN sigma' if adaptive else 'fixed sigma 4'}",
            "metric": "density maps/s", "unit": "maps/s", "value": 64 / (t_dev * 1e-3), "ms_total_device": t_dev,
            "ms_splat": t_splat, "ms_knn": t_dev - t_splat,
            "e2e": {"value": 64 / e2e_batch_s, "unit": "maps/s", "h2d_bytes_per_step": 16 * int(pl.total_heads) + plan.meta.nbytes,
                    "d2h_bytes_per_step": 4 * int(pl.total_pixels),
                    "note": "gaussian_filter_density_batch: host numpy heads in, host numpy maps out (one launch set, one packed D2H)",
                    "per_image_api_maps_per_s": 64 / e2e_s},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": HBM_GBS, "unit": "GB/s", "frac": achieved / HBM_GBS,
                         "traffic": None,
                         "note": "algorithmic bytes 4*H*W + 16*N per image over the time of the prepare + culling + splat kernels "
                                 "of the batch (kNN time reported separately: N^2 fp64 distance evaluations, compute only)"},
        }
        if cpu:
            ref = _reference_module("dmap_gen")
            cand = sorted([im for im in images if 15 <= len(im[1]) <= 400], key=lambda im: len(im[1]) * im[0][0] * im[0][1])
            heads, done, t0 = 0, 0, time.perf_counter()
            for (h, w), pts in cand:  # cheapest images with a non-trivial crowd first, until the budget is spent
                if done and time.perf_counter() - t0 > cpu_seconds / 2:
                    break
                if ref is not None:
                    (ref.gaussian_filter_density if adaptive else ref.gaussian_filter_density_fixed)(np.empty((h, w, 3)), pts)
                else:
                    from oracle import dmap_oracle
                    dmap_oracle.density_reference_like((h, w), pts, fixed=not adaptive)
                heads += len(pts)
                done += 1
            dt = time.perf_counter() - t0
            total_heads = sum(len(p) for _, p in images)
            line["cpu_baseline"] = {
                "value": 64 / (dt / max(heads, 1) * total_heads), "unit": "maps/s", "cores": 1,
                "kind": "reference" if ref is not None else "port",
                "sample": (f"{'unmodified utils/dmap_gen.py' if ref is not None else 'oracle port of utils/dmap_gen.py'} "
                           f"(scipy gaussian_filter per head, one process) on the {done} cheapest non-trivial images, {heads} heads in "
                           f"{dt:.1f} s; per-head cost extrapolated to the {total_heads} heads of the set (SURVEY 8d: the full "
                           f"set would take hours)")}
        lines["splat_adaptive" if adaptive else "splat_fixed"] = line
    return lines


def bench_dmap(cpu):
    for line in dmap_lines(cpu).values():
        print(json.dumps(line), flush=True)


def isw_lines(cpu, tf32_peak=None):
    """BASELINE config 5: {gram_c64, gram_c256, gram_c512} lines (covariance kernel roofline, module step, e2e, cpu_baseline)."""
    from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss
    lib = _native.lib()
    tf32_peak = tf32_peak or probe_tf32_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = {}
    for (b, c, h, w) in synthetic.CONFIG5_SHAPES:
        hw = h * w
        gen = torch.Generator().manual_seed(5000 + c)
        x_host = torch.randn(b, c, h, w, generator=gen)
        mask_host = (torch.rand(c, c, generator=gen) < 0.5).float().triu(1)   # relax_denom = 2 branch: half of the upper triangle
        x = x_host.to(dev).view(b, c, hw)
        eye = torch.eye(c, device=dev)
        n = lib.dgvcc_isw_workspace_bytes(b, c, hw)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        fc = torch.empty(b, c, c, device=dev)
        ts = []
        for rep in range(8):
            flush.zero_()
            e0, e1 = ev(), ev()
            e0.record()
            lib.dgvcc_isw_covariance(_native.ptr(x), _native.ptr(eye), b, c, hw, 1, _native.ptr(ws), n, _native.ptr(fc), _native.stream_ptr(dev))
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts.append(e0.elapsed_time(e1))
        t_cov = float(np.mean(ts))
        xin = x_host.to(dev).requires_grad_(True)
        mask = mask_host.to(dev)
        nrm = mask.sum()
        iw = InstanceWhitening(c)
        tm = []
        for rep in range(8):
            flush.zero_()
            xin.grad = None
            e0, e1 = ev(), ev()
            e0.record()
            y, wt = iw(xin)
            loss = instance_whitening_loss(wt, eye, mask, 0, nrm)
            loss.backward()
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                tm.append(e0.elapsed_time(e1))
        t_mod = float(np.mean(tm))
        # the same step captured once as a CUDA graph and replayed (what it costs inside a graph-captured training step:
        # the module is seven short kernels, eager launches leave gaps between them)
        t_graph = None
        try:
            xg = x_host.to(dev).requires_grad_(True)   # a fresh leaf: its gradient accumulator is born on the side stream
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    xg.grad = None
                    instance_whitening_loss(iw(xg)[1], eye, mask, 0, nrm).backward()
            torch.cuda.current_stream().wait_stream(side)
            xg.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                instance_whitening_loss(iw(xg)[1], eye, mask, 0, nrm).backward()
            tg = []
            for rep in range(8):
                flush.zero_()
                e0, e1 = ev(), ev()
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                if rep >= 2:
                    tg.append(e0.elapsed_time(e1))
            t_graph = float(np.mean(tg))
            del graph, xg
        except Exception as exc:  # keep the eager numbers
            t_graph = f"{type(exc).__name__}: {exc}"
        # end to end: pinned host feature map in, loss + gradient out
        x_pin, g_pin = x_host.pin_memory(), torch.empty_like(x_host).pin_memory()

        def host_step():
            xd = x_pin.to(dev, non_blocking=True).requires_grad_(True)
            _, wt_ = iw(xd)
            l_ = instance_whitening_loss(wt_, eye, mask, 0, nrm)
            l_.backward()
            g_pin.copy_(xd.grad, non_blocking=True)
            return float(l_.detach())
        for _ in range(3):
            host_step()
        t0 = time.perf_counter()
        for _ in range(10):
            host_step()
        e2e_s = (time.perf_counter() - t0) / 10
        flops_alg = 3.0 * b * c * c * hw        # 2 B C^2 HW x 3 TF32 MMAs per product / 2 (upper-triangular tiles)
        bytes_alg = 4.0 * b * c * hw            # the feature map read once
        t_tensor, t_hbm = flops_alg / tf32_peak, bytes_alg / (HBM_GBS * 1e9)
        bound = "tensor" if t_tensor >= t_hbm else "hbm"
        t1 = (c + 127) // 128
        tiles = t1 * (t1 + 1) // 2
        executed = 3 * 2.0 * (b if c > 64 else (b + 1) // 2) * tiles * 128 * 128 * hw
        roof = {"bound": bound, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                "achieved": (flops_alg / (t_cov * 1e-3) / 1e12) if bound == "tensor" else (bytes_alg / (t_cov * 1e-3) / 1e9),
                "peak": (tf32_peak / 1e12) if bound == "tensor" else HBM_GBS,
                "frac": max(t_tensor, t_hbm) / (t_cov * 1e-3), "traffic": None,
                "tf32_peak_tflops_probed": tf32_peak / 1e12, "executed_mma_tflops": executed / (t_cov * 1e-3) / 1e12,
                "hbm_gbs": bytes_alg / (t_cov * 1e-3) / 1e9,
                "note": ("dgvcc_isw_covariance (tcgen05 3xTF32 Gram partials + split-K finish); roofline time = max(3 B C^2 HW / "
                         "TF32 peak probed in this run by dgvcc_probe_tf32, 4 B C HW / measured HBM copy peak); frac = roofline "
                         "time / measured time")}
        line = {"workload": f"BASELINE config 5: ISW covariance loss, B={b} C={c} HW={hw} (fp32 in, 3xTF32 tensor cores)",
                "metric": "steps/s (InstanceWhitening + instance_whitening_loss fwd+bwd)", "unit": "steps/s",
                "value": 1e3 / t_mod, "ms_fwd_bwd": t_mod, "ms_fwd_bwd_cuda_graph": t_graph, "us_covariance": t_cov * 1e3,
                "roofline": roof,
                "e2e": {"value": 1 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": x_host.numel() * 4 + 4}}
        if cpu:
            ref = _reference_module("instance_whitening")
            xc = x_host.clone().requires_grad_(True)
            t0 = time.perf_counter()
            if ref is not None:
                _, wr = ref.InstanceWhitening(c)(xc)
                ref.instance_whitening_loss(wr, torch.eye(c), mask_host, 0, mask_host.sum()).backward()
            else:
                from oracle import isw_oracle
                wr = isw_oracle.instance_standardize(xc)
                isw_oracle.whitening_loss(wr, torch.eye(c), mask_host, 0, mask_host.sum()).backward()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1 / dt, "unit": "steps/s", "cores": torch.get_num_threads(),
                                    "kind": "reference" if ref is not None else "port",
                                    "sample": ("unmodified models/ISW/instance_whitening.py" if ref is not None else "oracle port of "
                                               "instance_whitening.py") + " on torch CPU, same tensors, one forward + backward"}
        lines[f"gram_c{c}"] = line
    return lines


def bench_isw(cpu):
    for line in isw_lines(cpu).values():
        print(json.dumps(line), flush=True)


def bench_bay(cpu):
    """Row f1: BayesianDataset._cal_dists + crop targets for one QNRF-size annotation set."""
    from dgvcc_b200.datasets import bay_targets
    lib = _native.lib()
    n = 12000
    pts = synthetic.crowd_points(np.random.default_rng(8212), n, 2048, 1536, dtype=np.float64)
    d_pts = torch.from_numpy(pts).to(dev)
    out = torch.empty((n, 1), dtype=torch.float64, device=dev)
    ts = []
    for rep in range(6):
        e0, e1 = ev(), ev()
        e0.record()
        lib.dgvcc_bay_knn_mean(_native.ptr(d_pts), n, 1, _native.ptr(out), _native.stream_ptr(dev))
        e1.record()
        torch.cuda.synchronize()
        if rep:
            ts.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    d = bay_targets.cal_dists(pts)
    bay_targets.crop_targets(pts, d, 300, 500, 512, 512)
    e2e = time.perf_counter() - t0
    line = {"workload": f"SURVEY 8f rank 1: BayesianDataset._cal_dists + crop targets, {n} heads (float64)",
            "metric": "annotation sets/s", "value_device_knn": 1e3 / min(ts), "ms_knn": min(ts),
            "value_e2e_host_numpy": 1 / e2e, "pairs_per_s": n * n / (min(ts) * 1e-3)}
    if cpu:
        from oracle import bay_targets_oracle as bo
        t0 = time.perf_counter()
        dr = bo.cal_dists(pts)
        bo.crop_targets(pts.copy(), dr, 300, 500, 512, 512)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1 / dt, "unit": "annotation sets/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle port (numpy N x N matrix + partition), the same 12 000 heads, once"}
    print(json.dumps(line), flush=True)


def bench_den(cpu):
    """Row f3: dataset-side density targets (crop + 8x8 sum-pool + flip + 16x16 occupancy) for a batch of maps."""
    from dgvcc_b200.datasets import den_targets
    rng = np.random.default_rng(8600)
    b, h, w, crop, down = 64, 1536, 2048, (1024, 1024), 8
    maps = [torch.from_numpy(np.where(rng.random((h, w)) < 0.01, rng.random((h, w)), 0).astype(np.float32)).to(dev) for _ in range(b)]
    geoms = [(0, 0, int(rng.integers(0, h - crop[0] + 1)), int(rng.integers(0, w - crop[1] + 1)), k % 2) for k in range(b)]
    lib = _native.lib()
    flat = torch.cat([m.reshape(-1) for m in maps])
    meta = np.zeros((b, _native.DEN_META_COLS), dtype=np.int64)
    for k, g in enumerate(geoms):
        meta[k] = (k * h * w, h, w, g[0], g[1], g[2], g[3], g[4])
    d_meta = torch.from_numpy(meta).to(dev)
    dh, dw = crop[0] // down, crop[1] // down
    out = torch.empty((b, 1, dh, dw), dtype=torch.float32, device=dev)
    bm = torch.empty((b, dh // 16, dw // 16), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for rep in range(6):
        flush.zero_()
        e0, e1 = ev(), ev()
        e0.record()
        _native.check(lib.dgvcc_den_train_targets(_native.ptr(flat), _native.ptr(d_meta), b, crop[0], crop[1], down, _native.ptr(out),
                                                  _native.ptr(bm), _native.stream_ptr(dev)), "den")
        e1.record()
        torch.cuda.synchronize()
        if rep:
            ts.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    den_targets.train_density_targets(maps, geoms, crop, down)
    torch.cuda.synchronize()
    api_s = time.perf_counter() - t0
    bytes_alg = b * (crop[0] * crop[1] * 4 + dh * dw * 4 + (dh // 16) * (dw // 16) * 4)
    line = {"workload": f"SURVEY 8f rank 3: density targets of {b} maps {h}x{w}: crop {crop[0]}x{crop[1]}, {down}x{down} sum-pool, flip, 16x16 occupancy",
            "metric": "maps/s", "value_device": b / (min(ts) * 1e-3), "ms": min(ts), "value_public_api_device_maps": b / api_s,
            "roofline": {"bound": "hbm", "achieved": bytes_alg / (min(ts) * 1e-3) / 1e9, "peak": HBM_GBS, "unit": "GB/s",
                         "frac": bytes_alg / (min(ts) * 1e-3) / 1e9 / HBM_GBS,
                         "note": "algorithmic bytes: every source pixel of the crops read once, targets written once"}}
    if cpu:
        from oracle import den_targets_oracle as do
        host = [m.cpu().numpy() for m in maps[:8]]
        t0 = time.perf_counter()
        for m, g in zip(host, geoms[:8]):
            do.block_occupancy(do.train_density(m, g[0], g[1], g[2], g[3], crop[0], crop[1], down, g[4]))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 8 / dt, "unit": "maps/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle port of den_cls_dataset.py:109-150 + :60-61 (torch CPU ops) on 8 of the maps"}
    print(json.dumps(line), flush=True)


def bench_cov(cpu):
    """Row f2: CovMatrix_ISW.set_mask_matrix (average of the statistics, top-k mask, AND with the previous mask)."""
    from dgvcc_b200.models.ISW.cov_settings import topk_mask
    c, n_stats = 512, 4
    g = torch.Generator().manual_seed(8700)
    stats = (torch.rand(n_stats, c * c, generator=g) ** 3).to(dev)
    prev = (torch.rand(c * c, generator=g) < 0.7).float().to(dev)
    k = (c * c - c) // 2 - (c * c - c) // 4
    ts = []
    for rep in range(6):
        e0, e1 = ev(), ev()
        e0.record()
        topk_mask(stats, n_stats, k, prev)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            ts.append(e0.elapsed_time(e1))
    line = {"workload": f"SURVEY 8f rank 2: CovMatrix_ISW mask construction, C={c} ({c * c} entries, {n_stats} statistics, k={k})",
            "metric": "masks/s", "value": 1e3 / min(ts), "ms": min(ts)}
    if cpu:
        from oracle.cov_settings_oracle import CovMatrixISW
        cm = CovMatrixISW(c, 2.0)
        cm.mask_matrix = prev.cpu().view(c, c)
        sc = stats.cpu()
        t0 = time.perf_counter()
        for s_ in sc:
            cm.set_variance_of_covariance(s_.view(c, c))
        cm.set_mask_matrix()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1 / dt, "unit": "masks/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle port of cov_settings.py:52-89 (torch CPU topk), same statistics"}
    print(json.dumps(line), flush=True)


def bench_auxloss(cpu):
    """Row f4: lw_loss / ortho_loss forward + backward (the Gram and dX = S X kernels of the ISW path)."""
    from dgvcc_b200.losses.lw import lw_loss
    from dgvcc_b200.losses.ortho import ortho_loss
    g = torch.Generator().manual_seed(8800)
    x = torch.randn(8, 256, 80, 80, generator=g)
    mask = (torch.rand(8, 1, 80, 80, generator=g) < 0.5).float()
    a, b_ = torch.randn(256, 16384, generator=g), torch.randn(256, 16384, generator=g)
    for name, fn, args in (("lw_loss (8,256,80,80) + mask", lw_loss, (x, mask)), ("ortho_loss (256,16384) x2", ortho_loss, (a, b_))):
        dargs = [t.to(dev).requires_grad_(t.dim() != 4 or t.shape[1] != 1) for t in args]
        ts = []
        for rep in range(6):
            for t in dargs:
                t.grad = None
            e0, e1 = ev(), ev()
            e0.record()
            fn(*dargs).backward()
            e1.record()
            torch.cuda.synchronize()
            if rep:
                ts.append(e0.elapsed_time(e1))
        line = {"workload": f"SURVEY 8f rank 4: {name}, forward + backward", "metric": "steps/s", "value": 1e3 / min(ts), "ms": min(ts)}
        if cpu:
            from oracle import aux_losses_oracle as ao
            cargs = [t.clone().requires_grad_(t.dim() != 4 or t.shape[1] != 1) for t in args]
            t0 = time.perf_counter()
            (ao.lw_loss if fn is lw_loss else ao.ortho_loss)(*cargs).backward()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1 / dt, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "oracle port (torch CPU), same tensors, one forward + backward"}
        print(json.dumps(line), flush=True)


def bench_sw(cpu):
    """Row f4: SwitchWhiten2d forward + backward at the ISW feature-map shapes; the four big passes are HBM-bound
    (forward 4+8 bytes per element, backward 8+12), so the line reports achieved HBM GB/s over those 32 bytes."""
    from dgvcc_b200.models.ISW.switchwhiten import SwitchWhiten2d
    g = torch.Generator().manual_seed(8900)
    for shape, sw_type in (((8, 64, 160, 160), 2), ((8, 256, 80, 80), 2), ((8, 256, 80, 80), 5), ((8, 512, 40, 40), 3)):
        x = torch.randn(*shape, generator=g)
        m = SwitchWhiten2d(shape[1], sw_type=sw_type).to(dev)
        xd, gy = x.to(dev).requires_grad_(True), torch.randn(*shape, generator=g).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        tf, tb = [], []
        for rep in range(6):
            xd.grad = None
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            y = m(xd)
            e1.record()
            y.backward(gy)
            e2.record()
            torch.cuda.synchronize()
            if rep:
                tf.append(e0.elapsed_time(e1))
                tb.append(e1.elapsed_time(e2))
        elems = x.numel()
        line = {"workload": f"SURVEY 8f rank 4: SwitchWhiten2d{shape} sw_type={sw_type} T=5, forward + backward (L2 flushed)",
                "metric": "steps/s", "value": 1e3 / (min(tf) + min(tb)), "ms_forward": min(tf), "ms_backward": min(tb),
                "roofline": {"bound": "hbm", "achieved": 32 * elems / ((min(tf) + min(tb)) * 1e-3) / 1e9, "peak": HBM_GBS,
                             "unit": "GB/s", "frac": 32 * elems / ((min(tf) + min(tb)) * 1e-3) / 1e9 / HBM_GBS,
                             "note": "algorithmic bytes: 4 (moments) + 8 (y) + 8 (backward moments) + 12 (grad_x) per element"}}
        if cpu:
            from oracle import switchwhiten_oracle as so
            mc = {k: v.detach().cpu() for k, v in m.state_dict().items()}
            t0 = time.perf_counter()
            so.forward_backward(x, gy.cpu(), mc["sw_mean_weight"], mc["sw_var_weight"], mc["weight"], mc["bias"],
                                mc["running_mean"], mc["running_cov"], sw_type=sw_type)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1 / dt, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "oracle port (the reference's torch ops on CPU), same tensors, one forward + backward"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.only in ("", "dmap"):
        bench_dmap(a.cpu)
    if a.only in ("", "isw"):
        bench_isw(a.cpu)
    if a.only in ("", "bay"):
        bench_bay(a.cpu)
    if a.only in ("", "den"):
        bench_den(a.cpu)
    if a.only in ("", "cov"):
        bench_cov(a.cpu)
    if a.only in ("", "auxloss"):
        bench_auxloss(a.cpu)
    if a.only in ("", "sw"):
        bench_sw(a.cpu)
