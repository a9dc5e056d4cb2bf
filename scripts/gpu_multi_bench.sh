#!/bin/bash
# Multi-GPU bench line only (no tests, no reference arm): bash scripts/gpu_multi_bench.sh <N> <tag>
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
cut -c1-300 gpurun_out/bench_${N}gpu_$TAG.json; echo
tail -3 gpurun_out/bench_${N}gpu_$TAG.err
