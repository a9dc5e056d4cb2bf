#!/bin/bash
# A/B of a fused-kernel macro on one GPU:  bash scripts/gpu_ab.sh TAG "NAME=VALUE" cfg:voices ...
TAG=$1; DEF=$2; shift 2
mkdir -p gpurun_out
{ echo "== default"; SRK_TUNE=0 timeout 600 python scripts/tune_report.py "$@"
  echo "== SRK_FUSED_DEFINE=$DEF"; SRK_TUNE=0 SRK_FUSED_DEFINE=$DEF timeout 600 python scripts/tune_report.py "$@"; } > gpurun_out/ab_$TAG.txt 2>&1
grep -E "^==|kernel|rror" gpurun_out/ab_$TAG.txt
