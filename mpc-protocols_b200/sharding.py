"""Multi-GPU sharding of the batch axis.

Every kernel of the hot path is a map over independent secrets / chunks / codewords, so a batch is split into contiguous
ranges, one per rank (one process per GPU, constant tables replicated), with NO collective on the data path.  The only
exchange is the optional gather of result shards (NCCL all_gather over NVLink on GPUs; the same code runs on gloo/CPU
tensors, which is how the host logic is tested without GPUs).
"""
from __future__ import annotations


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the larger shards."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]


def gather_shards(local, total: int, group=None):
    """All-gather per-rank result shards (tensor [local_count, ...]) into the full [total, ...] tensor on every rank.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local shard has the wrong length")
    mx = max(sizes)
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
