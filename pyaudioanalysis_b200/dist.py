"""Clip sharding across the GPUs of one box (one process per GPU, torch.distributed).

Clips are independent (no state crosses feature_extraction calls), so the path shards with no
data-path collective; the only exchange is the optional final gather of the per-rank
[clips, F, T] blocks to one rank (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_clips, rank, world):
    """Contiguous block of clips owned by ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(n_clips), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_blocks(local, n_clips, dst=0, group=None):
    """Gather per-rank blocks [n_local, ...] into [n_clips, ...] on ``dst`` (None elsewhere).

    Blocks may differ by one clip; they are padded to the largest block for the collective and
    trimmed afterwards.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_clips, r, world) for r in range(world)]
    big = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < big:
        pad = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def feature_extraction_sharded(all_clips_fn, n_clips, sampling_rate, window, step, deltas=True, gather_to=0,
                               compute=None, group=None):
    """Each rank extracts features for its block of clips; optionally gather on ``gather_to``.

    ``all_clips_fn(lo, hi)`` returns this rank's clips [hi-lo, N] on its device; ``compute`` defaults
    to the GPU path (``feature_extraction_batch``) and is injectable so the sharding logic can be
    tested on CPU with gloo.
    """
    if compute is None:
        from .batch import feature_extraction_batch
        compute = lambda x: feature_extraction_batch(x, sampling_rate, window, step, deltas=deltas)   # noqa: E731
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(n_clips, rank, world)
    local = compute(all_clips_fn(lo, hi))
    if gather_to is None:
        return local
    return gather_blocks(local, n_clips, dst=gather_to, group=group)
