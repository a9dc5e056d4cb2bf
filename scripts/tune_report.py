#!/usr/bin/env python
"""GPU: what the measured schedule choice picks, and the kernel time it gives, for a list of cfg:voices launches
(stems + mix in HBM, 48000 samples).   python scripts/tune_report.py cfg2:4096 cfg4:4096 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import srack_b200 as srk

N = 48000
builders = {n: c[0] for n, c in srk.patches.CONFIGS.items()}
builders.update({g.__name__: g for g in srk.patches.CFG5_GRAPHS})
for spec in sys.argv[1:]:
    name, V = spec.split(":")
    V = int(V)
    p = srk.Patch(device=0)
    builders[name](p, V)
    p.plan()
    stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0")
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    ms = []
    for i in range(4):
        p.render_into(V, N, 0, stems.data_ptr(), mix.data_ptr(), device_out=True)
        torch.cuda.synchronize()
        ms.append(p.last_render_ms()[0])
    info = p.program_info(V)
    print(f"{spec:18s} kernel {min(ms[1:]):8.3f} ms  fused={info['fused']} warps={info['n_warps']} group={info['fused_group']} regs={info['fused_regs']} "
          f"smem={info['smem_bytes']}\n    {p.schedule_report()}", flush=True)
    del stems, mix, p
    torch.cuda.empty_cache()
