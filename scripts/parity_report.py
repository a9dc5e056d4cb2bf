#!/usr/bin/env python
"""GPU: parity statistics of the BASELINE graphs against the oracle (max error, worst error / tolerance, max ulp distance,
fraction of bit-identical samples) -- the numbers the test thresholds are set from.   python scripts/parity_report.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import srack_b200 as srk  # noqa: E402
from oracle import orc  # noqa: E402
from util import build_both, parity_stats  # noqa: E402

cases = [(n, c[0], 1024) for n, c in srk.patches.CONFIGS.items()] + [("cfg3b_B1", srk.patches.cfg3b, 1), ("sequenced", srk.patches.sequenced, 1024),
                                                                      ("sampler", srk.patches.sampler, 1024)]
cases += [(g.__name__, g, 1024) for g in srk.patches.CFG5_GRAPHS[4:]]
for name, builder, B in cases:
    V, N = 96, 24000
    gp, op, _, _ = build_both(srk, orc, builder, V, buffer_size=B)
    gp.plan()
    g, _ = gp.render(V, N, stems=True)
    o, _ = op.render(V, N)
    s = parity_stats(g, o)
    print(f"{name:16s} B={B:5d} bit_identical {s['bit_identical']:.6f} max_ulp {s['max_ulp']:6d} max_abs {s['max_abs']:.3g} worst err/tol {s['worst_ratio']:.3g}", flush=True)
