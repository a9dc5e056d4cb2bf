"""Scale checks on one B200 (-m gpu): BASELINE's largest voice count on a single device WITH stems (262144 voices
= 100.7 GB, more than 2^32 elements per channel, so every index path is exercised beyond 32 bits) against the
oracle on sampled voices, and a million voices mix-only against the sum of four shards.  Skipped when the device
does not have the memory free."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 48000


def test_262144_voices_with_stems_and_a_million_mix_only(srk, orc, cuda_device):
    import torch

    free, _ = torch.cuda.mem_get_info(0)
    if free < 120 << 30:
        pytest.skip("needs 120 GB of free device memory")
    V = 262144
    p = srk.Patch(device=0)
    srk.patches.cfg2(p, V)
    p.plan()
    stems = torch.empty((2, N, V), dtype=torch.float32, device="cuda:0")
    mix = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    p.render_into(V, N, 0, stems.data_ptr(), mix.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    pick = [0, 1, 31, 32, 131071, 131072, 200001, V - 33, V - 1]
    got = stems[:, :, pick].cpu().numpy()
    op = orc.OraclePatch(48000, 1024, 2)
    srk.patches.cfg2(op, V)
    for j, v in enumerate(pick):
        ref, _ = op.render(1, N, voice_offset=v)
        assert (got[:, :, j].view(np.uint32) == ref[:, :, 0].view(np.uint32)).all(), v  # cfg2 is bit-exact
    s = torch.empty((2, N), dtype=torch.float64, device="cuda:0")
    for n0 in range(0, N, 500):  # (an f64 copy of the whole array would be 200 GB)
        s[:, n0:n0 + 500] = stems[:, n0:n0 + 500, :].sum(dim=2, dtype=torch.float64)
    bound = 1e-5 * max(np.sqrt(V), float(s.abs().max()))  # 1e-5 * max(|mix|, sqrt(V))
    assert float((s - mix.double()).abs().max()) <= bound
    del stems, s
    torch.cuda.empty_cache()

    V4 = 4 * V
    q = srk.Patch(device=0)
    srk.patches.cfg2(q, V4)
    q.plan()
    big = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    q.render_into(V4, N, 0, None, big.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    acc = torch.zeros((2, N), dtype=torch.float64, device="cuda:0")
    part = torch.empty((2, N), dtype=torch.float32, device="cuda:0")
    for r in range(4):
        q.render_into(V, N, r * V, None, part.data_ptr(), device_out=True)
        torch.cuda.synchronize()
        acc += part.double()
    bound = 1e-5 * max(np.sqrt(V4), float(acc.abs().max()))
    assert float((acc - big.double()).abs().max()) <= bound
