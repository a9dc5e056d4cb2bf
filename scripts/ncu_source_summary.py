#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instructions executed and stall samples per
source line (needs -lineinfo and --import-source on) and the stall-reason totals.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda... > src.csv ; python ncu_source_summary.py src.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
tot_inst = 0
tot_samp = 0
stalls = collections.Counter()
by_op = collections.Counter()
samp_by_op = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(float(r[col["Instructions Executed"]] or 0))
        s = int(float(r[col["# Samples"]] or 0))
    except ValueError:
        continue
    tot_inst += n
    tot_samp += s
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    by_op[op] += n
    samp_by_op[op] += s
    for h in stall_cols:
        try:
            stalls[h] += int(float(r[col[h]] or 0))
        except ValueError:
            pass
    lines.append((s, n, r[col["Address"]], r[col["Source"]]))
print(f"total warp instructions {tot_inst}, stall samples {tot_samp}")
print("stall reasons:", ", ".join(f"{k[6:]} {v} ({100*v/max(tot_samp,1):.1f}%)" for k, v in stalls.most_common(8)))
print("top opcodes by executed count:")
for op, n in by_op.most_common(25):
    print(f"  {op:10s} {n:12d} {100*n/tot_inst:5.1f}%   samples {100*samp_by_op[op]/max(tot_samp,1):5.1f}%")
if len(sys.argv) > 2:
    print("top SASS lines by samples:")
    for s, n, a, src in sorted(lines, reverse=True)[: int(sys.argv[2])]:
        print(f"  {s:7d} {n:10d} {a} {src[:100]}")
