"""Single-GPU tuning sweep of the Bayesian-loss path on BASELINE config 3: per-kernel device times (CUDA events between the
launches, L2 flushed) for the tuning knobs of dgvcc_bl_set_option.  One JSON line per setting.

    python scripts/bl_tune_1gpu.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dgvcc_b200 import _native  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.workload(0)
lib = _native.lib()
for cell in (64, 32):
    _native.check(lib.dgvcc_bl_set_option(_native.BL_OPT_MIN_CELL, cell), "dgvcc_bl_set_option")
    kernels, kept, packed = bench.kernel_breakdown(wl, dev, reps=8)
    print(json.dumps({"min_cell": cell, "path_ms": round(sum(kernels.values()), 4),
                      "kernels_ms": {k: round(v, 4) for k, v in kernels.items()}}), flush=True)
lib.dgvcc_bl_set_option(_native.BL_OPT_MIN_CELL, 64)
# points per chunk: more, shorter warp tasks (less wave quantisation at the tail of a sweep, more per-task prologues)
from dgvcc_b200.losses import bl as blmod  # noqa: E402
for chunk in (1024, 512, 256):
    blmod._CHUNK_POINTS = chunk
    kernels, kept, packed = bench.kernel_breakdown(wl, dev, reps=8)
    print(json.dumps({"chunk_points": chunk, "chunks": int(packed.total_chunks), "path_ms": round(sum(kernels.values()), 4),
                      "kernels_ms": {k: round(v, 4) for k, v in kernels.items()}}), flush=True)
blmod._CHUNK_POINTS = 1024
