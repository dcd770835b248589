#!/bin/bash
# Sweep the BL tuning knobs (point-chunk size, pixel tile per thread) on BASELINE config 3; prints one line per setting.
for v in 0 1; do for c in 512 768 1024 1536 2048; do
  DGVCC_BL_VARIANT=$v DGVCC_BL_CHUNK=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels_ms']
print('variant $v chunk $c: value %.0f frac %.3f path %.3f ms  ' % (d['value'], d['roofline']['frac'], d['roofline']['path_ms']), {a: round(b,3) for a,b in k.items()})"
done; done
