#!/usr/bin/env python
"""Small renders of every BASELINE patch under every kernel family and launch shape, meant to be run under
compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box:
    compute-sanitizer --tool racecheck python scripts/sanitize.py [barriers|flags|all]
`barriers`: every kernel whose warps synchronise through barriers only (fused one warp per group, the interpreter
kernels); `flags`: the staged fused kernels, whose stages hand tiles over through acquire / release counters in shared
memory -- racecheck models barriers, not flags, and reports that hand-over as hazards; memcheck and synccheck apply.
Fused kernels: one warp per voice group, 3 and 5 pipeline stages (tile rings + acquire/release counters in shared
memory), with the TMA stems path (voices % 4 == 0) and the per-lane store path; interpreter kernels: one warp per group
and pipelined."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import srack_b200 as srk

KNOBS = ("SRK_FUSED", "SRK_FUSED_STAGES", "SRK_FUSED_GROUP", "SRK_WARPS", "SRK_SOLO_GROUPS", "SRK_FUSED_DEFINE", "SRK_FUSED_SPLIT_MOOG")


def run(label, env, names, V, N=1500):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    for name in names:
        p = srk.Patch(srk.AudioConfig(48000, 256, 2))
        (srk.patches.CONFIGS[name][0] if name in srk.patches.CONFIGS else getattr(srk.patches, name))(p, V)
        p.plan()
        st, mx = p.render(V, N, stems=True, mix=True)
        st2, _ = p.render(V, 37, stems=True, mix=True)
        _, mx3 = p.render(V, 64, stems=False, mix=True)
        assert np.isfinite(st).all() and np.isfinite(mx).all() and np.isfinite(mx3).all()
        info = p.program_info(V)
        print(f"{label:28s} {name:10s} V={V} fused={info['fused']} warps={info['n_warps']} max|x|={float(np.abs(st).max()):.3f}", flush=True)


ALL = ("cfg1", "cfg2", "cfg3", "cfg3b", "cfg4", "sequenced", "sampler")
SECTION = sys.argv[1] if len(sys.argv) > 1 else "all"
if SECTION == "libm":
    # the table-driven libm restatements (libm_glibc.cuh): the fast sine's coefficient rows and exp2's table on the common
    # route, and -- with the tie band at its widest -- glibc's sin restated (the 440-entry __sincostab) on every sample
    run("fused, sine: fast + ties", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": "1"}, ("cfg1", "cfg3", "cfg3b", "cfg4"), 72, N=6000)
    run("fused, sine: all restated", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": "1", "SRK_FUSED_DEFINE": "SRK_SIN_TIE_BAND=268435456"},
        ("cfg1", "cfg3", "cfg3b", "cfg4"), 72, N=6000)
    run("fused staged, all restated", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": "4", "SRK_FUSED_DEFINE": "SRK_SIN_TIE_BAND=268435456"}, ("cfg4",), 72, N=3000)
    run("interpreter", {"SRK_FUSED": "0", "SRK_WARPS": "16"}, ("cfg1", "cfg3", "cfg4"), 70, N=3000)
    sys.exit(0)
for stages in ("1", "3", "5"):
    if (stages == "1" and SECTION == "flags") or (stages != "1" and SECTION == "barriers"):
        continue
    run(f"fused stages<={stages} (TMA)", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": stages}, ALL, 72)
    run(f"fused stages<={stages} (no TMA)", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": stages}, ("cfg2", "cfg3b", "cfg4"), 70)
if SECTION != "barriers":
    run("fused stages<=4, groups of 8", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": "4", "SRK_FUSED_GROUP": "8"}, ("cfg2", "cfg4", "cfg3b"), 72)
    run("fused stages<=6, no coef split", {"SRK_FUSED": "1", "SRK_FUSED_STAGES": "6", "SRK_FUSED_SPLIT_MOOG": "0"}, ("cfg2", "cfg4"), 72)
for warps in ("1", "16"):
    if SECTION == "flags":
        break
    run(f"interpreter warps<={warps}", {"SRK_FUSED": "0", "SRK_WARPS": warps}, ALL, 70)
# several voice groups per block in the interpreter's one-warp schedule, ragged last block and a table reload
for groups in ("3", "16") if SECTION != "flags" else ():
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({"SRK_FUSED": "0", "SRK_WARPS": "1", "SRK_SOLO_GROUPS": groups})
    p = srk.Patch(srk.AudioConfig(48000, 256, 2))
    h = srk.patches.sampler(p, 333)
    p.plan()
    st, mx = p.render(333, 700, stems=True, mix=True)
    h["sample"].set_sample(np.linspace(-1, 1, 50).astype(np.float32), 96000.0)
    st2, _ = p.render(333, 300, stems=True, mix=True)
    assert np.isfinite(st).all() and np.isfinite(st2).all()
    print("sampler groups per block", groups, float(np.abs(st).max()), flush=True)
print("sanitize ok")
