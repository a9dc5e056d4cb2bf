#!/usr/bin/env python
"""Key numbers of one kernel from `ncu -i X.ncu-rep --page raw --csv` (read on the CPU box).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_raw_summary.py raw.csv"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1])
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k][1]:>16s} {d[k][0]}")
