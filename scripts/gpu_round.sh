#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and full captures of the voice kernel.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_round.sh [tag] [quick|full|notest]
TAG=${1:-r01}
MODE=${2:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
if [ "$MODE" != "notest" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
  tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-1500 gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
if [ "$MODE" = "quick" ]; then exit 0; fi
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-600 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1
# headline kernel (cfg2 @ 4096 voices) and the full chip (cfg2 @ 65536): one launch each, full set, source on
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_voices|srk_fused" -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG}_cfg2_4096 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_voices|srk_fused" -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG}_cfg2_65536 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --voices-per-gpu 65536 >> gpurun_out/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_voices|srk_fused" -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG}_cfg4_32768 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --config cfg4 --voices-per-gpu 32768 >> gpurun_out/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_voices|srk_fused" -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG}_cfg3_65536 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --config cfg3 --voices-per-gpu 65536 >> gpurun_out/ncu_full_$TAG.log 2>&1
grep -o '"kernel": "[^"]*"' gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out | tail -12
