#!/usr/bin/env python
"""Small renders of every BASELINE patch under both schedules, meant to be run under
compute-sanitizer (memcheck / racecheck) on the GPU box:
    compute-sanitizer --tool racecheck python scripts/sanitize.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import srack_b200 as srk

for warps in ("1", "16"):
    os.environ["SRK_WARPS"] = warps
    for name in ("cfg1", "cfg2", "cfg3", "cfg3b", "cfg4", "sequenced", "sampler"):
        p = srk.Patch(srk.AudioConfig(48000, 256, 2))
        (srk.patches.CONFIGS[name][0] if name in srk.patches.CONFIGS else getattr(srk.patches, name))(p, 70)
        p.plan()
        st, mx = p.render(70, 1500, stems=True, mix=True)
        st2, _ = p.render(70, 37, stems=True, mix=True)
        assert np.isfinite(st).all() and np.isfinite(mx).all()
        print(name, "warps", warps, p.program_info(70)["n_warps"], float(np.abs(st).max()), flush=True)
# several voice groups per block in the one-warp schedule, ragged last block and a table reload
os.environ["SRK_WARPS"] = "1"
for groups in ("3", "16"):
    os.environ["SRK_SOLO_GROUPS"] = groups
    p = srk.Patch(srk.AudioConfig(48000, 256, 2))
    h = srk.patches.sampler(p, 333)
    p.plan()
    st, mx = p.render(333, 700, stems=True, mix=True)
    h["sample"].set_sample(np.linspace(-1, 1, 50).astype(np.float32), 96000.0)
    st2, _ = p.render(333, 300, stems=True, mix=True)
    assert np.isfinite(st).all() and np.isfinite(st2).all()
    print("sampler groups per block", groups, float(np.abs(st).max()), flush=True)
print("sanitize ok")
