#!/bin/bash
# Multi-GPU on one box: the NCCL result tests, then the default bench line (headline + cfg4_shard + cfg5 + cfg5_balanced
# sub-records) and the reference arm the way the driver launches them.   bash scripts/gpu_multi.sh <N> <tag>
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi_${N}gpu_$TAG.log 2>&1
tail -8 gpurun_out/pytest_multi_${N}gpu_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
cut -c1-400 gpurun_out/bench_${N}gpu_$TAG.json; echo
tail -3 gpurun_out/bench_${N}gpu_$TAG.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu_$TAG.json 2>> gpurun_out/bench_${N}gpu_$TAG.err
cut -c1-300 gpurun_out/bench_ref_${N}gpu_$TAG.json; echo
