#!/bin/bash
# One gpurun call after a kernel-source change: GPU parity tests, fresh schedule decisions for the BASELINE launch
# shapes (-> gpurun_out/tuned_$TAG, to be shipped as s-rack_b200/tuned/), the sine A/B, a bench line.
TAG=${1:-r06a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/tune_all.py gpurun_out/tuned_$TAG > gpurun_out/tune_all_$TAG.txt 2>&1
tail -3 gpurun_out/tune_all_$TAG.txt
if [ -z "$NO_AB" ]; then
{ echo "== sine port: CUDA sin + tie band (default)"; SRK_TUNE=0 timeout 300 python scripts/tune_report.py cfg1:65536 cfg3:65536 cfg4:32768
  echo "== sine port: every sample through the restated glibc sin (SRK_FUSED_DEFINE=SRK_SIN_TIE_BAND=268435456)"
  SRK_TUNE=0 SRK_FUSED_DEFINE=SRK_SIN_TIE_BAND=268435456 timeout 300 python scripts/tune_report.py cfg1:65536 cfg3:65536 cfg4:32768; } > gpurun_out/sine_ab_$TAG.txt 2>&1
grep -E "^==|kernel" gpurun_out/sine_ab_$TAG.txt
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-1200 gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_$TAG.err
