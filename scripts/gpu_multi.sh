#!/bin/bash
# Multi-GPU round: bash scripts/gpu_multi.sh <tag> <N>
tag=${1:-r2m}; N=${2:-2}
out=gpurun_out; mkdir -p $out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/shard_bl_multi_gpu.py --steps 20 > $out/${tag}_shard_${N}gpu.json 2> $out/${tag}_shard_${N}gpu.err; echo "shard rc=$?"
tail -c 1500 $out/${tag}_shard_${N}gpu.json; tail -5 $out/${tag}_shard_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-aux --no-eager --no-cpu-baseline > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_${N}gpu.json"))
print("value",d["value"],"e2e",d["e2e"]["value"])
print(json.dumps(d.get("strong"),indent=1)[:3000])
PY
tail -5 $out/${tag}_bench_${N}gpu.err
