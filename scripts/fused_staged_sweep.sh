#!/bin/bash
# kernel-time sweep of the staged fused kernels (few voice groups per SM): samples per straight-line group x stage count
# usage: bash scripts/fused_staged_sweep.sh cfg:voices [cfg:voices ...]
export SRK_FUSED=1
for spec in "$@"; do
  for g in 4 8 16; do for st in 0 3 4 5 6; do
    if [ "$st" = "0" ]; then unset SRK_FUSED_STAGES; else export SRK_FUSED_STAGES=$st; fi
    echo -n "group=$g stages_forced=$st  "
    SRK_FUSED_GROUP=$g python scripts/sweep.py $spec:0:0 2>&1 | tail -1
  done; done
done
