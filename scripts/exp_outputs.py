import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
import srack_b200 as srk
N, V = 48000, 4096
for name in ("cfg1", "cfg2"):
    for mode in ("both", "stems", "mix", "none"):
        p = srk.Patch(device=0)
        srk.patches.CONFIGS[name][0](p, V)
        p.plan()
        stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0")
        mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
        ms = []
        for i in range(4):
            p.render_into(V, N, 0, stems.data_ptr() if mode in ("both", "stems") else None,
                          mix.data_ptr() if mode in ("both", "mix") else None, device_out=True)
            torch.cuda.synchronize()
            ms.append(p.last_render_ms()[0])
        print(name, mode, "kernel %.3f ms" % min(ms[1:]), flush=True)
