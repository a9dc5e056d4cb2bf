#!/bin/bash
# kernel-time sweep of the fused kernel's build knobs: SRK_FUSED_GROUP (samples per straight-line group) x SRK_FUSED_MINB
# (register cap = 65536 / (128 * MINB)) x SRK_FUSED_OSC_RATE (-: by rule, 0: always the group test, 1: never)
# usage: bash scripts/fused_sweep.sh cfg:voices [cfg:voices ...]
export SRK_FUSED=1
for spec in "$@"; do
  for g in 4 8; do for mb in 4 2 3; do for rate in - 0 1; do
    if [ "$rate" = "-" ]; then unset SRK_FUSED_OSC_RATE; else export SRK_FUSED_OSC_RATE=$rate; fi
    echo -n "group=$g minb=$mb rate=$rate  "
    SRK_FUSED_GROUP=$g SRK_FUSED_MINB=$mb python scripts/sweep.py $spec:0:0 2>&1 | tail -1
  done; done; done
done
