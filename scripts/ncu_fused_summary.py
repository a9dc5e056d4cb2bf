#!/usr/bin/env python
"""Attribute an `ncu --page source --csv --print-source sass,cuda` dump of a fused kernel (srk_fused_kernel) to the
op templates of fused_ops.cuh: SASS is laid out op by op in plan order, so every instruction belongs to the op (and,
inside Osc / Adsr, the path) whose body its source line falls in; leaf helpers (fadd, wrap01, ...) stay with the op
that called them.  Prints per region: warp instructions executed, share, stall samples.
usage: ncu_fused_summary.py dump.csv [n_warp_samples]   (n_warp_samples = voices / 32 * samples: instr per sample)"""
import collections, csv, os, re, sys

dump = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else None
ops = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "s-rack_b200", "csrc", "fused_ops.cuh")
starts = []
for i, line in enumerate(open(ops), 1):
    m = re.match(r"struct (\w+) \{", line) or re.match(r"(?:static __device__ __noinline__|FZ_DEV) \S+ (\w+)\(", line) \
        or re.match(r"\s+FZ_DEV \S+ (\w+)\(", line) or re.match(r"\s+// region: (\w+)", line)
    if m:
        starts.append((i, m.group(1)))
LEAF = {"asf", "asu", "opaque", "ld_state", "st_state", "ld_param", "smem_u32"}  # stay with the op around them
OWNER = {"wrap01": "Osc", "fmod1_exact": "Osc", "ddiv_by_const": "Osc.blep", "blep_eval": "Osc.blep", "blep_eval_const": "Osc.blep",
         "clamp1": "Moog.run", "moog_coef": "Moog.coef", "philox4x32_10": "Noise.run", "nonlinear": "Math.run",
         "exp2f_glibc": "Sample.run", "tma_store_3d": "Out.flush", "tma_commit": "Out.flush", "tma_wait_read": "Out.begin_tile",
         "tma_wait_all": "Out.finish", "fence_async_smem": "Out.flush", "ctx_init": "prologue"}
STRUCTS = {"Osc", "Noise", "Moog", "Adsr", "Vca", "Mixer", "Math", "GridSeq", "PatSeq", "Sample", "Rings", "Out", "Ctx"}


def region_of(line_no):
    struct, fn = "?", ""
    for s, n in starts:
        if s > line_no:
            break
        if n in STRUCTS:
            struct, fn = n, ""
        else:
            fn = n
    return struct, fn


rows = list(csv.reader(open(dump, errors="replace")))
cur_file, hdr, cur_line = None, None, None
insts = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        sass_col = [i for i, h in enumerate(r) if h == "Source"][1]
        addr_col = r.index("Address")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur_line = int(r[0]); continue
    try:
        stalls = {h[6:]: int(r[i] or 0) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h}
        insts.append((int(r[addr_col], 16), cur_file, cur_line, r[sass_col].strip(), int(r[hdr["# Samples"]] or 0),
                      int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["stall_wait"]] or 0), int(r[hdr["stall_no_inst"]] or 0), stalls))
    except ValueError:
        pass
# an instruction inlined from the header is listed under the header's line AND under the generated file's call line: keep
# the header's entry
best = {}
for it in insts:
    if it[0] not in best or (it[1] == "fused_ops.cuh" and best[it[0]][1] != "fused_ops.cuh"):
        best[it[0]] = it
insts = sorted(best.values())
agg = collections.OrderedDict()
region = "prologue"
seen = collections.Counter()
why = collections.defaultdict(collections.Counter)
for addr, f, line, sass, samp, ex, wait, noinst, stalls in insts:
    if f == "fused_ops.cuh":
        struct, fn = region_of(line)
        if fn in OWNER:
            region = OWNER[fn]
        elif fn not in LEAF and struct != "?":
            region = struct + ("." + fn if fn and fn not in ("load", "store") else (".load/store" if fn else ""))
    elif f == "srk_fused.cu":
        pass  # the generated wiring: stays with the neighbouring op
    a = agg.setdefault(region, [0, 0, 0, 0, 0])
    a[0] += ex; a[1] += samp; a[2] += wait; a[3] += noinst; a[4] += 1
    why[region].update(stalls)
tot_ex = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot_ex}, samples {tot_s}" + (f", {tot_ex / per:.1f} warp instructions per warp-sample" if per else ""))
print(f"{'region':34s} {'SASS':>6s} {'warp instr':>13s} {'%':>6s}" + (f" {'/sample':>8s}" if per else "") + f" {'samples':>9s} {'%':>6s} {'wait%':>6s} {'noinst%':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] * 1000 < tot_ex:
        continue
    print(f"{k:34s} {a[4]:6d} {a[0]:13d} {100 * a[0] / max(tot_ex, 1):6.1f}" + (f" {a[0] / per:8.2f}" if per else "") +
          f" {a[1]:9d} {100 * a[1] / max(tot_s, 1):6.1f} {100 * a[2] / max(a[1], 1):6.1f} {100 * a[3] / max(a[1], 1):7.1f}  " +
          " ".join(f"{k}:{100 * v / max(a[1], 1):.0f}" for k, v in why[k].most_common(4)))
