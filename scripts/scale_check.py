#!/usr/bin/env python
"""Scale check on one B200: cfg2 at 262144 voices WITH stems (100.7 GB, > 2^32 elements per channel) against the
oracle on sampled voices, and 1 048 576 voices mix-only against the sum of four 262144-voice shards."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import srack_b200 as srk
from oracle import orc

N = 48000
V = 262144
p = srk.Patch(device=0)
srk.patches.cfg2(p, V)
p.plan()
stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0")
mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
p.render_into(V, N, 0, stems.data_ptr(), mix.data_ptr(), device_out=True)
torch.cuda.synchronize()
print("262144 voices with stems: kernel %.1f ms, stems %.1f GB" % (p.last_render_ms()[0], stems.numel() * 4 / 1e9), flush=True)
pick = [0, 1, 31, 32, 131071, 131072, 200001, V - 33, V - 1]
got = stems[:, :, pick].cpu().numpy()
op = orc.OraclePatch(48000, 1024, 2)
srk.patches.cfg2(op, V)
worst = 0.0
for j, v in enumerate(pick):
    ref, _ = op.render(1, N, voice_offset=v)
    same = (got[:, :, j].view(np.uint32) == ref[:, :, 0].view(np.uint32)).mean()
    worst = max(worst, float(np.abs(got[:, :, j] - ref[:, :, 0]).max()))
    assert same == 1.0, (v, same)
print("sampled voices bit-identical to the oracle:", pick, "max err", worst, flush=True)
s = torch.empty((2, N), dtype=torch.float64, device="cuda:0")
for n0 in range(0, N, 500):  # (an f64 copy of the whole array would be 200 GB)
    s[:, n0:n0 + 500] = stems[:, n0:n0 + 500, :].sum(dim=2, dtype=torch.float64)
err = (s - mix.double()).abs().max().item()
bound = 1e-5 * max(np.sqrt(V), float(s.abs().max()))  # 1e-5 * max(|mix|, sqrt(V))
print("mix vs f64 sum of stems: max err %.3g (bound %.3g, |mix| up to %.3g)" % (err, bound, float(s.abs().max())))
assert err <= bound
del stems, s
torch.cuda.empty_cache()

V4 = 4 * V
q = srk.Patch(device=0)
srk.patches.cfg2(q, V4)
q.plan()
big = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
q.render_into(V4, N, 0, None, big.data_ptr(), device_out=True)
torch.cuda.synchronize()
print("1048576 voices mix only: kernel %.1f ms = %.3g voice-samples/s" % (q.last_render_ms()[0], V4 * N / (q.last_render_ms()[0] * 1e-3)), flush=True)
acc = torch.zeros((2, N), dtype=torch.float64, device="cuda:0")
part = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
for r in range(4):
    q.render_into(V, N, r * V, None, part.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    acc += part.double()
err = (acc - big.double()).abs().max().item()
bound = 1e-5 * max(np.sqrt(V4), float(acc.abs().max()))
print("1048576-voice mix vs sum of four shards: max err %.3g (bound %.3g, |mix| up to %.3g)" % (err, bound, float(acc.abs().max())))
assert err <= bound
print("scale check ok")
