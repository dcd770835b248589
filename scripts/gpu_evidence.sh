#!/bin/bash
# Round evidence on ONE GPU: bash scripts/gpu_evidence.sh <tag>
# GPU tests, the bench line, the reference arm, ncu launch list and the two ncu --set full captures that
# scripts/summarize_profiles.py turns into profiles/<tag>_*.
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --cpu-seconds 60 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference rc=$?"
python scripts/bench_aux.py --cpu > $out/${tag}_aux_bench.jsonl 2> $out/${tag}_aux_bench.err; echo "aux rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux --no-eager > $out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
python scripts/profile_bl.py > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bl_ -s 7 -c 7 \
    -f -o $out/bl_${tag} python scripts/profile_bl.py > $out/${tag}_ncu_bl.log 2>&1; echo "ncu bl rc=$?"
python scripts/profile_aux.py > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:isw_|dmap_' -s 23 -c 23 \
    -f -o $out/aux_${tag} python scripts/profile_aux.py > $out/${tag}_ncu_aux.log 2>&1; echo "ncu aux rc=$?"
tail -n 3 $out/${tag}_pytest.log
