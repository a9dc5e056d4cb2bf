#!/bin/bash
# ncu --set full capture of the voice kernel for one bench configuration.
# Usage: bash scripts/gpu_profile.sh <tag> [bench.py args...]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_voices -s 3 -c 1 \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
