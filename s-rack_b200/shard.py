"""Voice sharding across GPUs: voices are independent patch instances, so
rank r renders a contiguous range of the global voice axis with no communication, and the
only collective is one sum of the [channels][n_samples] mix onto rank 0 (NCCL over NVLink on
GPUs; the same code runs over gloo on CPU tensors in the tests)."""


def voice_range(n_voices_total, rank, world_size):
    """Contiguous, balanced partition -> (voice_offset, n_voices) of `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(n_voices_total, world_size)
    off = rank * base + min(rank, rem)
    return off, base + (1 if rank < rem else 0)


def reduce_mix(mix, dst=0, group=None, all_ranks=False):
    """Sum the per-rank mixes (torch tensor [C][N], in place). Result valid on `dst`
    (every rank when all_ranks=True).  No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return mix
    if all_ranks:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(mix, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return mix


def render_sharded(render_fn, n_voices_total, rank, world_size, group=None, all_ranks=False):
    """`render_fn(voice_offset, n_voices) -> torch mix tensor [C][N]` for this rank's voices
    (n_voices may be 0 -> it must return zeros).  Returns the reduced mix."""
    off, cnt = voice_range(n_voices_total, rank, world_size)
    return reduce_mix(render_fn(off, cnt), group=group, all_ranks=all_ranks)
