"""Generates tests/golden/*.npz: small frozen outputs of the CPU oracle for the BASELINE
patches (a few voices, a few thousand samples).  The reference publishes no golden audio
and cannot be compiled here (Rust), so these vectors pin the *oracle* against regressions
and give the GPU tests a fixture that does not need the oracle at run time; the oracle
itself is pinned to the reference by tests/test_oracle_kat.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import orc  # noqa: E402
import srack_b200 as srk  # noqa: E402  (only for the patch descriptions)

CASES = {  # name: (n_voices, n_samples, buffer_size)
    "cfg1": (1, 4096, 1024),
    "cfg2": (5, 30000, 1024),   # long enough to cross attack/decay/sustain/release of the 2 Hz gate
    "cfg3": (6, 4096, 1024),
    "cfg3b": (6, 4096, 256),
    "cfg4": (4, 30000, 1024),
    "sequenced": (4, 24000, 1024),  # Grid + Pattern sequencers driving a voice
    "sampler": (4, 24000, 1024),    # Sample module with a bent playback rate
}
BUILDERS = {name: srk.patches.CONFIGS[name][0] for name in ("cfg1", "cfg2", "cfg3", "cfg3b", "cfg4")}
BUILDERS["sequenced"] = srk.patches.sequenced
BUILDERS["sampler"] = srk.patches.sampler
ONLY = sys.argv[1:]  # e.g. `make_golden.py sequenced` adds one fixture without touching the others

for name, (V, N, B) in CASES.items():
    if ONLY and name not in ONLY:
        continue
    p = orc.OraclePatch(48000, B, 2)
    BUILDERS[name](p, V)
    stems, _ = p.render(V, N)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), stems=stems, n_voices=V, n_samples=N, buffer_size=B)
    print(name, stems.shape, float(np.abs(stems).max()))
