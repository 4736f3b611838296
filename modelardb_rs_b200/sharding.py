"""Multi-GPU layout of the hot path: units (time series) are sharded over ranks, one process per GPU.

The path has no data-path exchange: a unit is compressed, gridded and aggregated entirely on the rank that owns
it (the reference likewise treats every series independently: compression.rs:95-104).  The only thing that
travels is the small per-series result of GROUP BY queries: every rank holds the aggregates of its own, disjoint
groups, so an all-gather in rank order yields the global result in unit order.  NCCL on GPUs, gloo in the CPU
tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native
from ._native import DEVICE, HOST


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


# ------------------------------------------------------------------------------------------------ in-library NCCL path
class Communicator:
    """mdbcu_comm (include/modelardb_cuda.h, "multi-GPU"): the library's own NCCL communicator, one per context / GPU.
    This is what a Rust host uses -- no torch.distributed involved; the functions above are the same exchange written
    with torch.distributed (gloo in the CPU tests) and define the expected result."""

    def __init__(self, handle, ctx, world: int, rank: int):
        self._h, self.ctx, self.world, self.rank = handle, ctx, world, rank

    @staticmethod
    def unique_id() -> bytes:
        """Rank 0 calls this and hands the 128 bytes to the other ranks by any means."""
        buf = (C.c_uint8 * 128)()
        _native.check(_native.lib().mdbcu_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def create(cls, ctx, world: int, rank: int, unique_id: bytes) -> "Communicator":
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        out = C.c_void_p()
        _native.check(_native.lib().mdbcu_comm_create(ctx._h, world, rank, buf, C.byref(out)))
        return cls(out, ctx, world, rank)

    def close(self):
        if self._h:
            _native.lib().mdbcu_comm_destroy(self._h)
            self._h = None

    def shard(self, n_units: int) -> Tuple[int, int]:
        lo, hi = C.c_uint64(), C.c_uint64()
        _native.check(_native.lib().mdbcu_shard_units(n_units, self.world, self.rank, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def aggregate_sharded(self, segments, group_off, n_local: int, n_total: int):
        """GROUP BY unit over a table sharded by mdbcu_shard_units: this rank's segments and group offsets in, all
        n_total groups out (count i64, min f32, max f32, sum f64) in unit order, on every rank.  One ncclAllGather."""
        from .compression import _order_streams, _ptr
        v, space = segments.view(), segments.space
        if isinstance(group_off, tuple):
            group_off = group_off[0]
        if space == HOST:
            if group_off is not None and not isinstance(group_off, int):
                group_off = np.ascontiguousarray(group_off, np.uint64)
            out = (np.zeros(n_total, np.int64), np.zeros(n_total, np.float32), np.zeros(n_total, np.float32), np.zeros(n_total, np.float64))
        else:
            import torch
            dev = f"cuda:{self.ctx.device}"
            out = (torch.empty(n_total, dtype=torch.int64, device=dev), torch.empty(n_total, dtype=torch.float32, device=dev),
                   torch.empty(n_total, dtype=torch.float32, device=dev), torch.empty(n_total, dtype=torch.float64, device=dev))
        gp = group_off if isinstance(group_off, int) else _ptr(group_off)
        _order_streams(space)
        _native.check(_native.lib().mdbcu_aggregate_sharded(self._h, space, C.byref(v), gp, n_local, n_total, *[_ptr(o) for o in out]))
        return out

    def aggregate_all_sharded(self, segments):
        """The ungrouped aggregate over rows sharded across the ranks (folded in rank order); one-element arrays."""
        from .compression import _order_streams, _ptr
        v, space = segments.view(), segments.space
        if space == HOST:
            out = (np.zeros(1, np.int64), np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float64))
        else:
            import torch
            dev = f"cuda:{self.ctx.device}"
            out = (torch.empty(1, dtype=torch.int64, device=dev), torch.empty(1, dtype=torch.float32, device=dev),
                   torch.empty(1, dtype=torch.float32, device=dev), torch.empty(1, dtype=torch.float64, device=dev))
        _order_streams(space)
        _native.check(_native.lib().mdbcu_aggregate_all_sharded(self._h, space, C.byref(v), *[_ptr(o) for o in out]))
        return out
