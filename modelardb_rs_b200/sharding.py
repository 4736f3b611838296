"""Multi-GPU layout of the hot path: units (time series) are sharded over ranks, one process per GPU.

The path has no data-path exchange: a unit is compressed, gridded and aggregated entirely on the rank that owns
it (the reference likewise treats every series independently: compression.rs:95-104).  The only thing that
travels is the small per-series result of GROUP BY queries: every rank holds the aggregates of its own, disjoint
groups, so an all-gather in rank order yields the global result in unit order.  NCCL on GPUs, gloo in the CPU
tests.
"""
from __future__ import annotations

from typing import Tuple


def shard_units(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range [lo, hi) of units owned by `rank` (the first n_units % world ranks get one more)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank / world out of range")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_group_aggregates(count, mn, mx, sm, n_units: int, group=None):
    """All-gathers the per-group (COUNT i64, MIN f32, MAX f32, SUM f64) tensors of every rank's own groups into the
    global arrays in unit order.  Shards may differ in size by one, so every rank pads to the largest shard."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    widest = -(-n_units // world)
    out = []
    for t in (count, mn, mx, sm):
        padded = torch.zeros(widest, dtype=t.dtype, device=t.device)
        padded[: t.numel()] = t
        gathered = torch.empty(world * widest, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        parts = []
        for r in range(world):
            lo, hi = shard_units(n_units, r, world)
            parts.append(gathered[r * widest: r * widest + (hi - lo)])
        out.append(torch.cat(parts))
    return tuple(out)
