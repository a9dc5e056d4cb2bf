#!/usr/bin/env python
"""GPU: the fused kernels against the interpreter kernels (bit for bit) on every BASELINE graph, then timings.
    python scripts/fused_check.py [--time]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srack_b200 as srk  # noqa: E402


def render(builder, V, N, fused, B=1024, stems=True, chunks=None, stages=None):
    os.environ["SRK_FUSED"] = "1" if fused else "0"
    if stages:
        os.environ["SRK_FUSED_STAGES"] = str(stages)
    else:
        os.environ.pop("SRK_FUSED_STAGES", None)
    p = srk.Patch(srk.AudioConfig(48000, B, 2))
    builder(p, V)
    p.plan()
    if chunks:
        st, mx = [], []
        for n in chunks:
            a, b = p.render(V, n, stems=stems, mix=True)
            st.append(a)
            mx.append(b)
        return (np.concatenate(st, axis=1) if stems else None), np.concatenate(mx, axis=1), p
    a, b = p.render(V, N, stems=stems, mix=True)
    return a, b, p


def main():
    graphs = dict(srk.patches.CONFIGS)
    names = list(graphs) + ["sequenced", "sampler"] + [g.__name__ for g in srk.patches.CFG5_GRAPHS]
    builders = {n: graphs[n][0] for n in graphs}
    builders["sequenced"] = srk.patches.sequenced
    builders["sampler"] = srk.patches.sampler
    for g in srk.patches.CFG5_GRAPHS:
        builders[g.__name__] = g
    bad = 0
    for name in dict.fromkeys(names):
        for V, N, B in ((70, 5000, 1024), (37, 1001, 7), (64, 777, 1)):
            t0 = time.time()
            s0, m0, _ = render(builders[name], V, N, False, B)
            s1, m1, p = render(builders[name], V, N, True, B)
            info = p.program_info(V)
            same = np.array_equal(s0.view(np.uint32), s1.view(np.uint32))
            mix_err = float(np.abs(m0.astype(np.float64) - m1).max())
            # chunked fused render == one-shot fused render, bit for bit (stems and mix)
            s2, m2, _ = render(builders[name], V, N, True, B, chunks=[N // 3, 5, N - N // 3 - 5])
            chunk_ok = np.array_equal(s1.view(np.uint32), s2.view(np.uint32)) and np.array_equal(m1.view(np.uint32), m2.view(np.uint32))
            # every stage count gives the same bits (1: one warp per group; 2, 4: pipelined slices of the patch)
            stage_ok = True
            for st in (1, 2, 3):
                s3, m3, _ = render(builders[name], V, N, True, B, stages=st)
                stage_ok &= np.array_equal(s1.view(np.uint32), s3.view(np.uint32)) and np.array_equal(m1.view(np.uint32), m3.view(np.uint32))
            ok = same and mix_err < 1e-3 and chunk_ok and info["fused"] == 1 and stage_ok
            bad += not ok
            print(f"{'ok ' if ok else 'BAD'} {name:16s} V={V} N={N} B={B} fused={info['fused']} regs={info['fused_regs']} "
                  f"local={info['fused_local_bytes']} stems_equal={same} "
                  f"frac_equal={float((s0.view(np.uint32) == s1.view(np.uint32)).mean()):.6f} mix_err={mix_err:.2e} chunk_invariant={chunk_ok} "
                  f"stages={info['n_warps']} stage_invariant={stage_ok} "
                  f"({time.time() - t0:.1f}s)", flush=True)
    if "--time" in sys.argv:
        import torch
        for name, V in (("cfg2", 4096), ("cfg2", 8192), ("cfg2", 16384), ("cfg2", 32768), ("cfg2", 65536), ("cfg3", 4096), ("cfg3", 65536),
                        ("cfg3b", 65536), ("cfg4", 4096), ("cfg4", 32768), ("cfg1", 65536)):
            N = 48000
            stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda")
            mix = torch.empty((2, N), dtype=torch.float32, device="cuda")
            for fused, st in ((0, None), (1, 1), (1, None)):
                os.environ["SRK_FUSED"] = str(fused)
                if st:
                    os.environ["SRK_FUSED_STAGES"] = str(st)
                else:
                    os.environ.pop("SRK_FUSED_STAGES", None)
                p = srk.Patch()
                graphs[name][0](p, V)
                p.plan()
                ks = []
                for i in range(4):
                    p.render_into(V, N, 0, stems.data_ptr(), mix.data_ptr(), device_out=True)
                    ks.append(p.last_render_ms()[0])
                info = p.program_info(V)
                print(f"time {name} V={V} fused={info['fused']} regs={info['fused_regs']} warps={info['n_warps']} kernel_ms={min(ks[1:]):.3f} "
                      f"-> {V * N / min(ks[1:]) / 1e6:.1f} G voice-samples/s", flush=True)
            del stems, mix
    print("FAILED" if bad else "ALL OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
