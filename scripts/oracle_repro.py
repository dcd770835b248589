"""Is the CPU checker reproducible across fresh processes on this host?  ("What's weak": the oracle
needed a warm-up call on the GPU box's 16-thread host.)

    python scripts/oracle_repro.py [processes]       # spawns fresh interpreters, default threads and 1 thread

Each child evaluates the smoke-test batch twice with the BL oracle and prints loss bits + gradient hash of the FIRST
and of the SECOND evaluation; the parent reports how many distinct results each setting produced.
"""
import collections
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, hashlib
sys.path.insert(0, %r)
import torch
threads = int(sys.argv[1])
if threads > 0:
    torch.set_num_threads(threads)
from dgvcc_b200 import synthetic
from oracle import bl_oracle
counts, (w, h), stride, sigma = [200, 0, 37], (512, 384), 8, 8.0
pts, tgt, dens, st = synthetic.bl_batch(5, counts, w, h, stride)
pts = [torch.from_numpy(p) for p in pts]; tgt = [torch.from_numpy(t) for t in tgt]
dens, st = torch.from_numpy(dens), torch.from_numpy(st)
out = []
for _ in range(2):
    l, g, _ = bl_oracle.bl_forward_backward(pts, st, tgt, dens, stride, sigma)
    out.append(float(l).hex() + ":" + hashlib.md5(g.numpy().tobytes()).hexdigest()[:8])
print(torch.get_num_threads(), *out)
''' % ROOT


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    for threads in (0, 8, 1):
        first, second = collections.Counter(), collections.Counter()
        for _ in range(n):
            r = subprocess.run([sys.executable, "-c", CHILD, str(threads)], capture_output=True, text=True)
            parts = r.stdout.split()
            if len(parts) != 3:
                print("child failed:", r.stderr[-300:])
                continue
            first[parts[1]] += 1
            second[parts[2]] += 1
            used = parts[0]
        print(f"threads={'default' if threads == 0 else threads} (torch used {used}): {n} fresh processes -> "
              f"{len(first)} distinct FIRST results {dict(first)}, {len(second)} distinct SECOND results {dict(second)}", flush=True)


if __name__ == "__main__":
    main()
