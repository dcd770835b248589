#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, checker reproducibility, and the ncu launch list.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 $out/${tag}_bench.json
timeout 400 python scripts/oracle_repro.py 12 > $out/${tag}_oracle_repro.log 2>&1
cat $out/${tag}_oracle_repro.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux --no-eager > $out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
