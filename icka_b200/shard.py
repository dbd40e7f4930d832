"""Batch sharding across ranks: sentences are independent, so a batch is cut into contiguous
per-rank shards exactly like the reference's ``DistributedSampler`` / ``DataParallel`` scatter
(My_cross_attention.py:707, 777-779).  Inference needs no data-path collective; the helpers below are
the only distributed logic: shard bounds, and gathering ragged per-rank tag lists back in order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n`` sentences owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, world: int, rank: int) -> dict:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return {k: v[lo:hi] for k, v in batch.items()}


def gather_tag_lists(local: Sequence[Sequence[int]], group=None) -> List[List[int]]:
    """All ranks obtain the full, ordered list of decoded tag sequences (ragged -> all_gather_object)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [list(x) for x in local]
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, [list(x) for x in local], group=group)
    return [seq for part in parts for seq in part]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing reduction bench.py uses: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
