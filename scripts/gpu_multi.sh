#!/bin/bash
# Multi-GPU bench lines on one box: bash scripts/gpu_multi.sh <N> <tag>
N=$1; TAG=$2
mkdir -p gpurun_out
run() {  # name, port, extra bench args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 \
      bench.py --gpus $N ${@:3} > gpurun_out/bench_${N}gpu_$1_$TAG.json 2>> gpurun_out/bench_${N}gpu_$TAG.err
  cut -c1-260 gpurun_out/bench_${N}gpu_$1_$TAG.json; echo
}
run cfg2 29521 --steps 10 --warmup 3
run cfg4 29522 --steps 3 --warmup 3 --config cfg4 --voices-per-gpu 32768 --no-cpu-baseline
run cfg5 29523 --steps 3 --warmup 3 --config cfg5
tail -3 gpurun_out/bench_${N}gpu_$TAG.err
