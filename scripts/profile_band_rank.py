"""One rank's share of a row-band sharded step (BASELINE config 3, 8 ranks emulated inside this process on one GPU:
the kernels, grids and arguments of rank r are what a real 8-GPU run launches on GPU r).  Run under ncu with
-k regex:'bl_(z|counts|grad|gridmin|grid_build)' to profile the per-rank sweeps.

    python scripts/profile_band_rank.py [world]
"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from test_bl_sharded_gpu import _case, run_banded  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg = _case("config3")
losses, grad, plan = run_banded(world, pts, tgt, st, dens, stride, sigma, bg_ratio, use_bg, False, steps=2)
torch.cuda.synchronize()
print("loss", float(losses[0]), "chunks", plan.total_chunks, "chunk", plan.chunk, "tile", plan.layout.rows_per_thread,
      plan.layout.cols_per_thread)
