"""One rank of tests/test_gpu_multi.py (launched by torch.distributed.run, one process per GPU, NCCL): renders its
contiguous voice range of cfg4, sums the mixes with the one NCCL reduce of the product path (srk.shard.render_sharded),
and rank 0 compares the result with the same voices rendered in one go on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srack_b200 as srk  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V, N = int(sys.argv[1]), int(sys.argv[2])
    name = sys.argv[3] if len(sys.argv) > 3 else "cfg4"
    builder = srk.patches.CONFIGS[name][0]
    stream = torch.cuda.current_stream().cuda_stream
    keep = {}

    def render(off, cnt):
        p = srk.Patch(device=local)
        builder(p, V)
        p.plan()
        mix = torch.zeros((2, N), dtype=torch.float32, device=dev)
        stems = torch.empty((2, N, max(cnt, 1)), dtype=torch.float32, device=dev)
        p.render_into(cnt, N, off, stems.data_ptr() if cnt else None, mix.data_ptr(), device_out=True, stream=stream)
        keep["stems"], keep["range"] = stems, (off, cnt)
        return mix

    reduced = srk.shard.render_sharded(render, V, rank, world)
    # every rank's stems to rank 0 (test only), to check shard placement voice by voice
    sizes = [srk.shard.voice_range(V, r, world)[1] for r in range(world)]
    parts = [torch.empty((2, N, max(c, 1)), dtype=torch.float32, device=dev) for c in sizes] if rank == 0 else None
    if len(set(sizes)) == 1:
        dist.gather(keep["stems"], parts, dst=0)
    ok = True
    if rank == 0:
        p = srk.Patch(device=local)
        builder(p, V)
        p.plan()
        st, mx = p.render(V, N, stems=True, mix=True)
        got = reduced.cpu().numpy()
        bound = 1e-5 * np.maximum(np.abs(mx.astype(np.float64)), np.sqrt(V))
        err = np.abs(got.astype(np.float64) - mx.astype(np.float64))
        ok = bool((err <= bound).all()) and bool(np.abs(got).max() > 0)
        print(f"multi-gpu mix: world {world}, {name} {V} voices x {N}: max err {err.max():.3g}, worst err/bound {(err / bound).max():.3g}")
        if len(set(sizes)) == 1:
            all_stems = torch.cat(parts, dim=2).cpu().numpy()
            same = bool((all_stems.view(np.uint32) == st.view(np.uint32)).all())
            print("multi-gpu stems bit-identical to the single-GPU render:", same)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
