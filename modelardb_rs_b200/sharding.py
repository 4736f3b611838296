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


def combine_global_aggregates(count, mn, mx, sm, group=None):
    """Ungrouped aggregates over rows sharded across ranks: COUNT and SUM add, MIN and MAX reduce with the
    NaN-ignoring fold of the accumulators, whose identities are f32::MAX / f32::MIN (model_simple_aggregates.rs:97, 117).
    Inputs are this rank's one-element tensors (as returned by aggregate(segments, None)); every rank gets the result.
    The f64 sums are gathered and added in rank order, so the result does not depend on the reduction tree."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)

    def gather(t):
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out

    counts, mins, maxs, sums = gather(count), gather(mn), gather(mx), gather(sm)
    total_count = counts.sum().reshape(1)
    not_nan_min, not_nan_max = mins[~torch.isnan(mins)], maxs[~torch.isnan(maxs)]
    total_min = not_nan_min.min().reshape(1) if not_nan_min.numel() else mins[-1:].clone()
    total_max = not_nan_max.max().reshape(1) if not_nan_max.numel() else maxs[-1:].clone()
    total_sum = torch.zeros(1, dtype=sm.dtype, device=sm.device)
    for r in range(world):  # rank order = row order
        total_sum = total_sum + sums[r:r + 1]
    return total_count, total_min, total_max, total_sum
