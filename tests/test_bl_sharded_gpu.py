"""Point-chunk sharding of the Bayesian loss (dgvcc_b200/losses/bl_sharded.py) on ONE GPU: `world` ranks run inside
this process, one CUDA stream each, exchanging partials through the same kernels, flags and phase order as one process
per GPU does over NVLink (LocalComm: plain pointers instead of CUDA IPC mappings).

Gates: loss and every owner's gradient BIT-IDENTICAL to one GPU running the same chunk table through
dgvcc_bl_forward / _backward (the exchange must not change a bit), and within the usual rtol 1e-5 of the CPU oracle.
"""
import types

import numpy as np
import pytest
import torch

from dgvcc_b200 import synthetic
from oracle import bl_oracle
from helpers import assert_close, load_bl_golden

pytestmark = pytest.mark.gpu


def single_gpu_with_table(plan, points, targets, st, dens, stride, sigma, bg_ratio, use_bg, cull):
    """One GPU, the plan's chunk table, the ordinary fused path."""
    from dgvcc_b200.losses import bl as blmod
    dev = torch.device("cuda:0")
    packed = types.SimpleNamespace(
        pts=torch.cat([p.reshape(-1, 2) for p in points]).to(dev) if plan.total_points else torch.zeros((1, 2), device=dev),
        meta=torch.from_numpy(plan.meta_all()).to(dev), total_rows=plan.total_rows, total_chunks=plan.total_chunks,
        multi_chunk=plan.multi_chunk, batch=plan.batch)
    tg = torch.cat([t.reshape(-1) for t in targets]).to(dev) if plan.total_points else torch.zeros((1,), device=dev)
    d = dens.to(dev).clone().requires_grad_(True)
    loss = blmod._FusedBL.apply(d, packed, tg, st.to(dev), float(stride), float(sigma), float(bg_ratio), bool(use_bg),
                                1.0 / plan.batch, None, cull)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), d.grad.cpu()


def run_sharded(world, points, targets, st, dens, stride, sigma, bg_ratio, use_bg, cull, owners=None, steps=1, defer=False):
    from dgvcc_b200.losses.bl_sharded import ChunkShardedBL, LocalComm, plan_shards
    dev = torch.device("cuda:0")
    b, _, hp, wp = dens.shape
    comms = LocalComm.make(world, dev, nbytes=192 << 20)
    plan = plan_shards([len(p) for p in points], use_bg, world, owners, hp, wp)
    mods = [ChunkShardedBL(sigma, max(hp, wp) * stride, stride, bg_ratio, use_bg, dev, c) for c in comms]
    for m in mods:
        m.exact_cull = cull
        m.defer_loss = defer   # the loss value then completes with the backward launches
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    # torch's own kernels are loaded lazily too: run the ones a step uses once before any wait kernel can be spinning
    warm = torch.zeros((2, 1, hp, wp), device=dev, requires_grad=True)
    (warm * 2.0).sum().backward()
    warm.grad.clone().to(torch.float32).contiguous()
    st_d = st.to(dev)
    locals_ = [dens[plan.owned[r]].to(dev).clone().requires_grad_(True) for r in range(world)]
    torch.cuda.synchronize()
    for _ in range(steps):
        losses = []
        for r in range(world):
            locals_[r].grad = None
            with torch.cuda.stream(streams[r]):
                losses.append(mods[r](points, st_d, targets, locals_[r], owners))
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                losses[r].backward()
        for m in mods:
            m.check()
    grad = torch.zeros_like(dens)
    for r in range(world):
        if len(plan.owned[r]):
            grad[plan.owned[r]] = locals_[r].grad.cpu()
    return [l.detach().cpu() for l in losses], grad, plan


def _case(name):
    if name.startswith("config"):
        cfg = int(name[-1])
        w, h = synthetic.CONFIG_SHAPES[cfg]
        counts = synthetic.config_counts(cfg)
        pts, tgt, dens, st = synthetic.bl_batch(cfg, counts, w, h, 8)
        return ([torch.from_numpy(p) for p in pts], [torch.from_numpy(t) for t in tgt], torch.from_numpy(st),
                torch.from_numpy(dens), 8, 8.0, 1.0, True)
    c = load_bl_golden(name)
    return c["points"], c["targets"], c["st_sizes"], c["density"], c["stride"], c["sigma"], c["bg_ratio"], c["use_bg"]


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", ["mixed", "nobg", "empty", "outside", "config2"])
def test_sharded_ranks_reproduce_one_gpu_bit_for_bit(name, world, monkeypatch):
    monkeypatch.setattr("dgvcc_b200.losses.bl._CHUNK_POINTS", 1024 if name == "config2" else 29)
    pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = _case(name)
    owners = [(3 * i + 1) % world for i in range(len(pts))]          # density owners unrelated to where the points fall
    for cull in (False, True):
        losses, grad, plan = run_sharded(world, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull, owners, steps=2,
                                         defer=cull)
        ref_loss, ref_grad = single_gpu_with_table(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull)
        for r, l in enumerate(losses):
            assert torch.equal(l, ref_loss), f"loss on rank {r}: {float(l)!r} vs {float(ref_loss)!r}"
        assert torch.equal(grad, ref_grad), f"gradient differs in {int((grad != ref_grad).sum())} pixels"
    o_loss, o_grad, _ = bl_oracle.bl_forward_backward(pts, st, tgt, dens, stride, sigma, bg_ratio, use_bg)
    assert_close(losses[0], o_loss, 1e-5, 0, "sharded loss vs oracle")
    assert_close(grad, o_grad, 1e-5, 2e-7 * float(o_grad.abs().max()), "sharded gradient vs oracle")


def test_config3_batch_on_four_emulated_ranks():
    """BASELINE config 3 (16 images, 49 697 heads): the 12 000-head image is spread over several ranks."""
    pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = _case("config3")
    losses, grad, plan = run_sharded(4, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, False)
    assert max(len(g) for g in plan.groups) >= 2
    per_rank = [int(plan.c_cnt[plan.chunk_lo[r]:plan.chunk_hi[r]].sum()) for r in range(4)]
    assert max(per_rank) - min(per_rank) <= 1
    ref_loss, ref_grad = single_gpu_with_table(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, False)
    assert all(torch.equal(l, ref_loss) for l in losses) and torch.equal(grad, ref_grad)
    from dgvcc_b200.losses.bl import BL
    dev = torch.device("cuda:0")
    d = dens.to(dev).requires_grad_(True)
    mod = BL(sigma, 2048, stride, bg_ratio, use_bg, dev)
    l1 = mod([p.to(dev) for p in pts], st.to(dev), [t.to(dev) for t in tgt], d)
    l1.backward()
    # against the ordinary single-GPU module (another chunk table: other summation order of the chunk partials)
    assert_close(losses[0], l1.detach().cpu(), 2e-6, 0, "sharded vs BL loss")
    assert_close(grad, d.grad.cpu(), 1e-5, 1e-6 * float(d.grad.abs().max()), "sharded vs BL gradient")


# ------------------------------------------------------------------------------------------------ row-band sharding
def run_banded(world, points, targets, st, dens, stride, sigma, bg_ratio, use_bg, cull, owners=None, steps=1, chunk=None):
    """`world` ranks of BandShardedBL inside this process (LocalComm), one stream each."""
    from dgvcc_b200.losses.bl_banded import BandShardedBL, LocalComm, plan_bands
    dev = torch.device("cuda:0")
    b, _, hp, wp = dens.shape
    comms = LocalComm.make(world, dev, nbytes=192 << 20)
    plan = plan_bands([len(p) for p in points], use_bg, world, owners, hp, wp, chunk)
    mods = [BandShardedBL(sigma, max(hp, wp) * stride, stride, bg_ratio, use_bg, dev, c) for c in comms]
    for m in mods:
        m.exact_cull, m.chunk = cull, chunk
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    warm = torch.zeros((2, 1, hp, wp), device=dev, requires_grad=True)
    (warm * 2.0).sum().backward()
    warm.grad.clone().to(torch.float32).contiguous()
    st_d = st.to(dev)
    locals_ = [dens[plan.owned[r]].to(dev).clone().requires_grad_(True) for r in range(world)]
    torch.cuda.synchronize()
    for _ in range(steps):
        losses = []
        for r in range(world):
            locals_[r].grad = None
            with torch.cuda.stream(streams[r]):
                losses.append(mods[r](points, st_d, targets, locals_[r], owners))
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                losses[r].backward()
        for m in mods:
            m.check()
    grad = torch.zeros_like(dens)
    for r in range(world):
        if len(plan.owned[r]):
            grad[plan.owned[r]] = locals_[r].grad.cpu()
    return [l.detach().cpu() for l in losses], grad, plan


def _band_table_on_one_gpu(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull):
    shim = types.SimpleNamespace(total_points=plan.total_points, meta_all=lambda: plan.meta, total_rows=plan.total_rows,
                                 total_chunks=plan.total_chunks, multi_chunk=plan.multi_chunk, batch=plan.batch)
    return single_gpu_with_table(shim, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull)


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("name", ["mixed", "nobg", "empty", "outside", "config2"])
def test_banded_ranks_agree_and_match_one_gpu(name, world):
    """Row-band sharding on emulated ranks: every rank returns the SAME loss bits; the gradient -- per-pixel work that
    never leaves its rank -- is BIT-IDENTICAL to one GPU sweeping the same chunk table; the loss differs from one GPU
    only by the two-level (band, then rank) order of the count sums (bit-identical for world == 1); both within
    rtol 1e-5 of the CPU oracle."""
    pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = _case(name)
    chunk = 1024 if name == "config2" else 29
    owners = [(3 * i + 1) % world for i in range(len(pts))]
    for cull in (False, True):
        losses, grad, plan = run_banded(world, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull, owners, steps=2,
                                        chunk=chunk)
        assert all(torch.equal(l, losses[0]) for l in losses), [float(l) for l in losses]
        ref_loss, ref_grad = _band_table_on_one_gpu(plan, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, cull)
        if world == 1:
            assert torch.equal(losses[0], ref_loss)
        assert_close(losses[0], ref_loss, 1e-6, 0, "banded loss vs one GPU")
        assert torch.equal(grad, ref_grad), f"gradient differs in {int((grad != ref_grad).sum())} pixels"
    o_loss, o_grad, _ = bl_oracle.bl_forward_backward(pts, st, tgt, dens, stride, sigma, bg_ratio, use_bg)
    assert_close(losses[0], o_loss, 1e-5, 0, "banded loss vs oracle")
    assert_close(grad, o_grad, 1e-5, 2e-7 * float(o_grad.abs().max()), "banded gradient vs oracle")


def test_config3_batch_on_eight_emulated_band_ranks():
    """BASELINE config 3 (16 images, 49 697 heads, 192 x 256 grid) over 8 bands of 24 grid rows."""
    pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = _case("config3")
    losses, grad, plan = run_banded(8, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, False)
    assert (plan.band_hi - plan.band_lo).tolist() == [24] * 8
    assert all(torch.equal(l, losses[0]) for l in losses)
    from dgvcc_b200.losses.bl import BL
    dev = torch.device("cuda:0")
    d = dens.to(dev).requires_grad_(True)
    mod = BL(sigma, 2048, stride, bg_ratio, use_bg, dev)
    l1 = mod([p.to(dev) for p in pts], st.to(dev), [t.to(dev) for t in tgt], d)
    l1.backward()
    # against the ordinary single-GPU module (1024-point chunks: another summation order of the chunk partials)
    assert_close(losses[0], l1.detach().cpu(), 2e-6, 0, "banded vs BL loss")
    assert_close(grad, d.grad.cpu(), 1e-5, 1e-6 * float(d.grad.abs().max()), "banded vs BL gradient")
    culled, grad_c, _ = run_banded(8, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, True)
    assert torch.equal(culled[0], losses[0]) and torch.equal(grad_c, grad)   # exact-zero culling changes no bit


def _torchrun(script, nproc, *args, timeout=600):
    import os
    import socket
    import subprocess
    import sys
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", script)
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                           "--master-addr", "127.0.0.1", "--master-port", str(port), path, *args],
                          capture_output=True, text=True, timeout=timeout)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    return proc.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one process per GPU, CUDA IPC peer memory)")
def test_real_processes_over_peer_memory():
    """scripts/shard_bl_multi_gpu.py under torchrun on every GPU of the box (at most 8): loss and gradients of the
    golden 'mixed' case, BASELINE config 2 and config 3 bit-identical to one GPU running the same chunk table."""
    import json
    out = _torchrun("shard_bl_multi_gpu.py", min(8, torch.cuda.device_count()), "--steps", "3")
    line = json.loads(out[out.index("{"):].splitlines()[0])
    assert line["ok"] and all(line["checks"].values()), line["checks"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_isw_loss_sharded_over_processes():
    """scripts/isw_multi_gpu.py: the ISW covariance loss with the batch split over the GPUs equals the whole batch on one
    GPU (loss and gradients, rtol 1e-6) for the three config-5 layers."""
    import json
    out = _torchrun("isw_multi_gpu.py", min(8, torch.cuda.device_count()), "--steps", "2")
    line = json.loads(out[out.index("{"):].splitlines()[0])
    assert line["ok"], line
