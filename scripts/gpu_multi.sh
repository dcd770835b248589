#!/bin/bash
# Multi-GPU round: bash scripts/gpu_multi.sh <tag> <N> [chunks]
tag=${1:-r2m}; N=${2:-2}; chunks=${3:-}
out=gpurun_out; mkdir -p $out
export NCCL_DEBUG=WARN
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 scripts/shard_bl_multi_gpu.py --steps 20 ${chunks:+--chunks $chunks} > $out/${tag}_shard_${N}gpu.json 2> $out/${tag}_shard_${N}gpu.err; echo "shard rc=$?"
run 29513 scripts/isw_multi_gpu.py --steps 20 > $out/${tag}_isw_${N}gpu.json 2> $out/${tag}_isw_${N}gpu.err; echo "isw rc=$?"
run 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-aux --no-eager --no-cpu-baseline > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
for f in shard isw bench; do tail -n 3 $out/${tag}_${f}_${N}gpu.err | cut -c1-300; done
