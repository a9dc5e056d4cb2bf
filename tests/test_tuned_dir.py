"""s-rack_b200/tuned/*.tune -- the schedule decisions measured on B200 that ship with the library -- name
kernels by id, and an id is a hash of the kernel's source, the headers, the NVRTC options and version: any edit of
fused_ops.cuh / libm_glibc.cuh / fused_gen.cpp silently orphans every decision (the library then measures again on
first use, correct but not the launch the committed ncu captures describe).  build() compiles the cost model's kernel
and every alternative of the BASELINE launch shapes into kernel_cache/; a shipped decision is live exactly when the
kernel it picks is among them."""
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TUNED = os.path.join(ROOT, "s-rack_b200", "tuned")
CACHE = os.path.join(ROOT, "s-rack_b200", "kernel_cache")


def test_every_shipped_decision_picks_a_kernel_the_current_sources_generate(srk):
    cubins = {os.path.basename(f).split(".")[0] for f in glob.glob(os.path.join(CACHE, "*.cubin"))}
    if len(cubins) < 20:
        pytest.skip("no precompiled kernel cache (python -c 'import __graft_entry__ as g; g.build()' fills it; needs NVRTC)")
    files = sorted(glob.glob(os.path.join(TUNED, "*.tune")))
    assert len(files) >= 20  # cfg2 @ 4096, cfg2 / cfg3 / cfg3b @ 65536, cfg5's eight graphs @ 32768 and @ 16384
    stale = []
    for f in files:
        lines = open(f).read().splitlines()
        pick, cands, shape = lines[0], lines[2:-1], lines[-1]
        assert pick in cands and shape.startswith("V="), f
        assert "whole-render kernel ms" in lines[1], f  # scripts/tune_all.py's second pass decided, not the short window
        missing = [c for c in cands if c.startswith("fused:") and c.split(":")[1] not in cubins]
        if missing:
            stale.append((os.path.basename(f), shape, len(missing), len(cands)))
    assert not stale, f"decisions whose candidates the current sources no longer generate (re-run scripts/tune_all.py on a B200): {stale}"
