#!/usr/bin/env python
"""How good are the library's schedule decisions?  From a scripts/tune_all.py log: per launch shape the whole-render
kernel time of the cost model's choice, of the window measurement's pick (engine.cu tune_schedule) and of the best
candidate.   python scripts/tuner_regret.py profiles/r06i_tune_all.txt"""
import re
import sys

txt = open(sys.argv[1]).read().splitlines()
window = {}
for l in txt:
    m = re.match(r"(\S+):(\d+)\s+kernel\s+([\d.]+) ms\s+(\S+)", l)
    if m and "ships" not in l:
        window[(m.group(1), int(m.group(2)))] = m.group(4)
rows = []
for l in txt:
    m = re.match(r"(\S+):(\d+)\s+ships (\S+)\s+([\d.]+) ms\s+\((.*)\)", l)
    if not m:
        continue
    key = (m.group(1), int(m.group(2)))
    full = [(c.split()[0], float(c.split()[1])) for c in m.group(5).split(", ")]
    best = min(t for _, t in full)
    w = dict(full)[window[key]]
    rows.append((key, full[0][1], w, best, len(full)))
print(f"{'launch':24s} {'cands':>5s} {'model':>8s} {'window':>8s} {'best':>8s}   model / window regret")
for key, model, w, best, n in rows:
    print(f"{key[0] + ' @ ' + str(key[1]):24s} {n:5d} {model:8.3f} {w:8.3f} {best:8.3f}   {100 * (model / best - 1):5.1f} % / {100 * (w / best - 1):4.1f} %")
n = len(rows)
print(f"mean regret over {n} shapes: cost model {sum(100 * (m / b - 1) for _, m, _, b, _ in rows) / n:.1f} %, "
      f"window measurement {sum(100 * (w / b - 1) for _, _, w, b, _ in rows) / n:.2f} % "
      f"(worst {max(100 * (w / b - 1) for _, _, w, b, _ in rows):.1f} %, exact on {sum(1 for _, _, w, b, _ in rows if w == b)} of {n})")
