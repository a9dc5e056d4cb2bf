#!/usr/bin/env python
"""Kernel-time sweep over schedules (SRK_WARPS / SRK_STEP) for the BASELINE configs; mix + stems in HBM.
Usage: python scripts/sweep.py [cfg:V:warps:step[:groups[:opbarrier]] ...]   (0 = library default;
groups / opbarrier: SRK_SOLO_GROUPS / SRK_SOLO_OP_BARRIER of the one-warp schedule)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import srack_b200 as srk

N = 48000
DEFAULT = ["cfg2:4096:0:0", "cfg2:4096:16:16", "cfg2:4096:16:8", "cfg2:4096:1:8", "cfg2:4096:1:32", "cfg2:4096:4:32",
           "cfg2:65536:0:0", "cfg2:65536:1:16", "cfg2:65536:1:32", "cfg2:65536:4:16", "cfg2:65536:2:32",
           "cfg3:65536:0:0", "cfg3:65536:1:16", "cfg3:65536:4:16", "cfg3b:65536:0:0",
           "cfg4:32768:0:0", "cfg4:32768:1:8", "cfg4:32768:4:16", "cfg4:32768:8:16", "cfg4:32768:16:16"]


def run(spec):
    name, V, warps, step, groups, opbar = (spec.split(":") + ["0", "0"])[:6]
    V, warps, step, groups, opbar = int(V), int(warps), int(step), int(groups), int(opbar)
    for k, v in (("SRK_WARPS", warps), ("SRK_STEP", step), ("SRK_SOLO_GROUPS", groups), ("SRK_SOLO_OP_BARRIER", opbar)):
        if v:
            os.environ[k] = str(v)
        else:
            os.environ.pop(k, None)
    p = srk.Patch(device=0)
    srk.patches.CONFIGS[name][0](p, V)
    p.plan()
    info = p.program_info(V)
    want_stems = V * N * 8 <= 60 << 30
    stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0") if want_stems else None
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    ms = []
    for i in range(4):
        p.render_into(V, N, 0, stems.data_ptr() if want_stems else None, mix.data_ptr(), device_out=True)
        torch.cuda.synchronize()
        ms.append(p.last_render_ms()[0])
    best = min(ms[1:])
    print(f"{spec:22s} warps={info['n_warps']:2d} stages={info['n_stages']} K={info['step_samples']:2d} "
          f"smem={info['smem_bytes']:6d} kernel {best:9.3f} ms  {V * N / best / 1e6:9.1f} Mvs/s"
          f"{'' if want_stems else '  (mix only)'}", flush=True)
    del stems, mix, p
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for spec in (sys.argv[1:] or DEFAULT):
        try:
            run(spec)
        except Exception as e:  # keep sweeping
            print(f"{spec}: FAILED {e}", flush=True)
