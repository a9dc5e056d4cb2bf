#!/bin/bash
# kernel-time sweep of the staged fused kernels' build knobs: prefetch x ladder-coefficient split x samples per group (x stages)
# usage: bash scripts/fused_knob_sweep.sh cfg:voices [cfg:voices ...]
export SRK_FUSED=1
for spec in "$@"; do
  for st in 0 4 5 6; do for g in 4 8; do for sp in 0 1; do for pf in 0 1; do
    if [ "$st" = "0" ]; then unset SRK_FUSED_STAGES; else export SRK_FUSED_STAGES=$st; fi
    echo -n "stages_forced=$st group=$g split=$sp prefetch=$pf  "
    SRK_FUSED_GROUP=$g SRK_FUSED_SPLIT_MOOG=$sp SRK_FUSED_PREFETCH=$pf python scripts/sweep.py $spec:0:0 2>&1 | tail -1
  done; done; done; done
done
