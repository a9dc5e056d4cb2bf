#!/usr/bin/env python
"""Attribute an `ncu --page source --csv --print-source sass,cuda` dump of render_voices_kernel to the
interpreter's ops: SASS is laid out op by op, so every instruction is assigned to the op whose body its
source line falls in (dsp.cuh function ranges, engine.cu for the interpreter loop / output / rings).
Prints, per op: warp instructions executed, non-barrier stall samples (= time the warp spends there),
barrier samples.  usage: ncu_ops_summary.py dump.csv [path/to/dsp.cuh]"""
import csv, sys, re, collections, os

dump = sys.argv[1]
dsp = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "s-rack_b200", "csrc", "dsp.cuh")
# function ranges in dsp.cuh: "__device__ ... name(" starts a range
starts = []
for i, line in enumerate(open(dsp), 1):
    m = re.match(r"struct (\w+Op) \{", line) or re.match(r"__device__ \S+ \S+ (\w+)\(", line)
    if m:
        starts.append((i, m.group(1)))
def fn_of(line_no):
    name = "?"
    for s, n in starts:
        if s <= line_no:
            name = n
    return name

# voice_kernel.cuh: the output / mix / ring helpers are regions of their own; the rest of that file (resident
# loops, barriers, the interpreter) stays with the op whose body precedes it in the SASS
vk = os.path.join(os.path.dirname(dsp), "voice_kernel.cuh")
vk_starts = []
if os.path.exists(vk):
    for i, line in enumerate(open(vk), 1):
        m = re.match(r"__device__ __forceinline__ void (run_output|run_mix|run_ring_load|run_ring_store)\(", line)
        if m:
            vk_starts.append((i, m.group(1)))
        elif re.match(r"(template|// A warp that owns ONE instruction)", line) and vk_starts and vk_starts[-1][1] != "":
            vk_starts.append((i, ""))
def vk_fn_of(line_no):
    name = ""
    for s_, n in vk_starts:
        if s_ <= line_no:
            name = n
    return name

rows = list(csv.reader(open(dump, errors="replace")))
cur_file, hdr, cur_line = None, None, None
insts = []  # (addr, file, line, sass, samples, barrier, executed)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # duplicate "Source" header: second one is SASS
        sass_col = [i for i, h in enumerate(r) if h == "Source"][1]
        addr_col = r.index("Address")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur_line = int(r[0]); continue
    try:
        insts.append((int(r[addr_col], 16), cur_file, cur_line, r[sass_col].strip(), int(r[hdr["# Samples"]] or 0),
                      int(r[hdr["stall_barrier"]] or 0), int(r[hdr["Instructions Executed"]] or 0)))
    except ValueError:
        pass
insts.sort()
LEAF = {"fadd", "fsub", "fmul", "dadd", "dsub", "dmul", "fmod1_exact", "wrap01", "blep_eval", "philox4x32_10", "wire",
        "for_groups", "clamp1", "moog_coef", "math_op", "nonlinear", "?"}
agg = collections.OrderedDict()
region = "prologue"
for addr, f, line, sass, samp, bar, ex in insts:
    if f == "dsp.cuh":
        fn = fn_of(line)
        if fn not in LEAF:
            region = fn
    elif f == "voice_kernel.cuh":
        fn = vk_fn_of(line)
        if fn:
            region = fn
    elif f == "engine.cu":
        region = "engine.cu"
    a = agg.setdefault(region, [0, 0, 0])
    a[0] += ex; a[1] += samp - bar; a[2] += bar
tot_ex = sum(a[0] for a in agg.values()); tot_s = sum(a[1] + a[2] for a in agg.values())
print(f"total warp instructions {tot_ex}, samples {tot_s}")
print(f"{'region':22s} {'warp instr':>12s} {'%':>6s} {'busy samples':>13s} {'%':>6s} {'barrier samples':>16s}")
for k, a in agg.items():
    print(f"{k:22s} {a[0]:12d} {100*a[0]/max(tot_ex,1):6.1f} {a[1]:13d} {100*a[1]/max(tot_s,1):6.1f} {a[2]:16d}")
