"""The N>1 path on CPU: two gloo ranks each render their contiguous voice range (with the CPU
oracle standing in for the GPU kernel) and reduce the mix onto rank 0 with the product's
shard.py; the result must equal the single-process render of all voices."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_voice_range_partitions_exactly(srk):
    for total in (0, 1, 7, 4096, 262144, 262145):
        for world in (1, 2, 3, 8):
            ranges = [srk.shard.voice_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0
            assert sum(c for _, c in ranges) == total
            for (o0, c0), (o1, _) in zip(ranges, ranges[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
    with pytest.raises(ValueError):
        srk.shard.voice_range(8, 2, 2)


def _worker(rank, world, port, n_voices, n_samples, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import srack_b200 as srk
    from oracle import orc

    def render_fn(off, cnt):
        p = orc.OraclePatch(48000, 256, 2)
        srk.patches.cfg4(p, n_voices)  # per-voice arrays are global; the rank renders a slice
        _, mix = p.render(cnt, n_samples, voice_offset=off, stems=False, mix=True)
        return torch.from_numpy(mix.astype(np.float32))

    mix = srk.shard.render_sharded(render_fn, n_voices, rank, world)
    if rank == 0:
        np.save(out_path, mix.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_single_process(srk, orc, tmp_path):
    import torch.multiprocessing as mp

    n_voices, n_samples = 7, 13312  # odd voice count: ranks get 4 + 3; the 2 Hz gate opens at sample 12000
    out = str(tmp_path / "mix.npy")
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, n_voices, n_samples, out), nprocs=2, join=True)
    got = np.load(out)
    p = orc.OraclePatch(48000, 256, 2)
    srk.patches.cfg4(p, n_voices)
    _, ref = p.render(n_voices, n_samples, stems=False, mix=True)
    assert got.shape == (2, n_samples)
    assert np.allclose(got, ref, rtol=0, atol=1e-5 * np.sqrt(n_voices))
    assert np.abs(got).max() > 0.01
