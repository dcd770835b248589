tag=r2c
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux --no-eager > $out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
python scripts/profile_bl.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:bl_ -s 7 -c 7 \
    -f -o $out/bl_${tag} python scripts/profile_bl.py > $out/${tag}_ncu_bl.log 2>&1; echo "ncu bl rc=$?"
tail -n 2 $out/${tag}_pytest.log
