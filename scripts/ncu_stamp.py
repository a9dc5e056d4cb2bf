#!/usr/bin/env python
"""Files one `ncu --set full` capture of a voice-kernel launch into profiles/traffic.json under the identity of the
kernel image it was taken on (srk_kernel_id), so that bench.py quotes DRAM traffic / instruction counts only for the
kernel that actually runs (measurement item 7).

usage (CPU box, after the .ncu-rep came back in gpurun_out/):
    ncu -i gpurun_out/prof_X.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_stamp.py /tmp/raw.csv <kernel_id> <config> <voices> <source note>
The kernel id of a launch: `python -c "import srack_b200 as s; p = s.Patch(); s.patches.cfg2(p, 4096); p.plan();
print(p.kernel_id(4096))"` (no GPU needed)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "profiles", "traffic.json")
M = {"dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "inst": "smsp__inst_executed.sum",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "duration": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
     "block": "launch__block_size"}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def num(d, key):
    if key not in d:
        return None
    unit, val = d[key]
    v = float(val.replace(",", ""))
    return v * UNIT_SCALE.get(unit, 1.0)


def main():
    raw, kernel_id, config, voices, source = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, zip(units, rows[2])))  # first captured launch
    entry = {"config": config, "voices": voices, "kernel_name": d.get("Kernel Name", ("", "?"))[1],
             "dram_bytes_per_launch": num(d, M["dram_read"]) + num(d, M["dram_write"]),
             "warp_instructions_per_launch": num(d, M["inst"]), "ncu_duration_ms": num(d, M["duration"]),
             "registers_per_thread": num(d, M["regs"]), "grid": num(d, M["grid"]), "block": num(d, M["block"]),
             "source": source,
             "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()}
    for k in ("issue_active_pct", "warps_active_pct", "fp64_pipe_pct", "fma_pipe_pct", "alu_pipe_pct", "dram_throughput_pct"):
        v = num(d, M[k])
        if v is not None:
            entry[k] = v
    try:
        doc = json.load(open(PATH))
    except Exception:
        doc = {}
    if "kernels" not in doc:
        doc = {"_what": "ncu --set full captures of single voice-kernel launches, keyed by srk_kernel_id (hash of the kernel's source): "
                        "bench.py quotes an entry only when the kernel it runs has exactly this id and voice count",
               "kernels": {}}
    doc["kernels"][kernel_id] = entry
    json.dump(doc, open(PATH, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
