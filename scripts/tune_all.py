#!/usr/bin/env python
"""GPU: measure the schedule choice for the launch shapes the BASELINE configs use and collect the decisions, so that
they can be shipped in s-rack_b200/tuned/ (consulted after the machine's own cache: a fresh checkout then launches the
kernels the committed ncu captures describe, without measuring first).  Any change to fused_ops.cuh / libm_glibc.cuh /
fused_gen.cpp changes every kernel id and with it every decision's key: re-run this and replace the directory.

    python scripts/tune_all.py gpurun_out/tuned_new        # one B200, about a minute

Shapes: cfg2 @ 4096 (BASELINE configs[1]), cfg2 / cfg3 / cfg3b @ 65536, the eight cfg5 graphs (cfg4 among them)
@ 32768 (one graph per GPU, and cfg4's shard of configs[3]) and @ 16384 (the halves cfg5_balanced deals out)."""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import srack_b200 as srk

N = 48000
out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "tuned_new")
cache = os.path.join(ROOT, "s-rack_b200", "kernel_cache")
for f in glob.glob(os.path.join(cache, "*.tune")):  # measure afresh
    os.remove(f)
builders = {n: c[0] for n, c in srk.patches.CONFIGS.items()}
builders.update({g.__name__: g for g in srk.patches.CFG5_GRAPHS})
shapes = [("cfg2", 4096), ("cfg2", 65536), ("cfg3", 65536), ("cfg3b", 65536)]
shapes += [(g.__name__, v) for v in (32768, 16384) for g in srk.patches.CFG5_GRAPHS]
for name, V in shapes:
    p = srk.Patch(device=0)
    builders[name](p, V)
    p.plan()
    stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0")
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    ms = []
    for i in range(3):
        p.render_into(V, N, 0, stems.data_ptr(), mix.data_ptr(), device_out=True)
        torch.cuda.synchronize()
        ms.append(p.last_render_ms()[0])
    print(f"{name}:{V:<6d} kernel {min(ms[1:]):8.3f} ms  {p.kernel_id(V)}\n    {p.schedule_report()}", flush=True)
    del stems, mix, p
    torch.cuda.empty_cache()
os.makedirs(out_dir, exist_ok=True)
n = 0
for f in glob.glob(os.path.join(cache, "*.tune")):
    shutil.copy(f, out_dir)
    n += 1
print(f"{n} decisions -> {out_dir}")
