"""Host-side mirrors of the two query operators that sit on the hot path in the reference, driving the CUDA
kernels through the C-ABI (modelardb_rs_b200.compression):

  * GridStream            crates/modelardb_storage/src/query/grid_exec.rs:197-430   
  * Model*Accumulator     crates/modelardb_storage/src/optimizer/model_simple_aggregates.rs:336-618   (row a23)

Same names, same state machines and the same results as the Rust operators (COUNT / MIN / MAX and every reconstructed point
bit for bit; the f64 SUM / AVG within 1e-12 relative, because the row sums are added in a fixed tree instead of one after the
other: include/modelardb_cuda.h, mdbcu_aggregate); the per-row loops of the reference
(`modelardb_compression::grid` / `sum` / `len` once per segment) are replaced by ONE batched call per segment batch.
The reference's toolchain is not in this image, so these are Python where the reference is Rust; INTEGRATION.md
shows the Rust call sites that would bind the same C-ABI entry points.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import compression as mc

F32_MAX = np.float32(3.4028234663852886e38)
F32_MIN = np.float32(-3.4028234663852886e38)


class GridStreamMetrics:
    """grid_exec.rs:433-520: rows created in total and per model type, with residuals, with regular timestamps."""

    def __init__(self):
        self.rows_created = 0
        self.rows_by_model_type = {mc.PMC_MEAN_ID: 0, mc.SWING_ID: 0, mc.MACAQUE_V_ID: 0}
        self.segments_with_residuals = 0
        self.segments_with_regular_timestamps = 0

    def add_batch(self, segments: mc.HostSegments, point_off: np.ndarray):
        lens = np.diff(point_off).astype(np.int64)
        self.rows_created += int(lens.sum())
        for type_id in self.rows_by_model_type:
            self.rows_by_model_type[type_id] += int(lens[segments.model_type_id == type_id].sum())
        self.segments_with_residuals += int(np.count_nonzero(np.diff(segments.residuals_off)))
        ts_len = np.diff(segments.timestamps_off)
        first_byte = segments.timestamps_data[np.minimum(segments.timestamps_off[:-1], max(len(segments.timestamps_data) - 1, 0)).astype(np.int64)] \
            if len(segments.timestamps_data) else np.zeros(len(segments), np.uint8)
        regular = (ts_len == 0) | ((first_byte & 128) == 0)  # are_compressed_timestamps_regular, timestamps.rs:199-202
        self.segments_with_regular_timestamps += int(np.count_nonzero(regular))


class _TagRuns:
    """The tag columns of the points of the current batch as runs: run r covers points [off[r], off[r+1]) and carries
    values[c][r] for tag column c.  One segment row is one run (its tag values repeated len() times in the reference,
    grid_exec.rs:340-356), so the replication the reference does per created row becomes arithmetic on offsets."""

    def __init__(self, n_columns: int):
        self.off = np.zeros(1, np.int64)
        self.values: List[np.ndarray] = [np.zeros(0, object) for _ in range(n_columns)]

    def tail(self, offset: int) -> "_TagRuns":
        """The runs of points [offset, end), rebased to 0."""
        out = _TagRuns(len(self.values))
        r0 = int(np.searchsorted(self.off, offset, "right")) - 1
        if offset >= self.off[-1]:
            return out
        out.off = np.concatenate([[0], self.off[r0 + 1:] - offset])
        out.values = [v[r0:] for v in self.values]
        return out

    def extend(self, lens: np.ndarray, values: Sequence[np.ndarray]):
        keep = lens > 0
        self.off = np.concatenate([self.off, self.off[-1] + np.cumsum(lens[keep])])
        self.values = [np.concatenate([old, np.asarray(new, object)[keep]]) for old, new in zip(self.values, values)]

    def filter(self, keep: np.ndarray):
        """Only the points with keep[i] survive."""
        if len(self.off) == 1:
            return
        lens = np.add.reduceat(keep.astype(np.int64), self.off[:-1])  # runs are non-empty, so the starts increase
        alive = lens > 0
        self.off = np.concatenate([[0], np.cumsum(lens[alive])])
        self.values = [v[alive] for v in self.values]

    def slice(self, lo: int, hi: int) -> List[Tuple[np.ndarray, np.ndarray]]:
        """(values, run lengths) per tag column for points [lo, hi)."""
        if hi <= lo:
            return [(np.zeros(0, object), np.zeros(0, np.int64)) for _ in self.values]
        r0 = int(np.searchsorted(self.off, lo, "right")) - 1
        r1 = int(np.searchsorted(self.off, hi, "left"))
        lens = np.diff(np.clip(self.off[r0:r1 + 1], lo, hi))
        return [(v[r0:r1], lens) for v in self.values]


class GridStream:
    """Reconstructs data points from batches of segments and hands them out in batches of `batch_size` rows:
    (timestamps int64[], values float32[], tag columns...), sorted as the input is (grid_exec.rs:187-195).

    input: an iterable of (segments, tags) where `segments` is a HostSegments / CompressedSegments batch and `tags` a
    sequence of per-row tag arrays (one array per tag column, possibly none).
    time_range / device_time_clip: the time predicate pushed down -- whole segments outside (lo, hi) are skipped, and with
    device_time_clip the points outside it are dropped inside the grid kernels' call (mdbcu_grid_range).
    predicate: optional function (timestamps, values) -> bool mask applied to every reconstructed batch, the
    reference's `maybe_predicate` (all points are reconstructed, then pruned: grid_exec.rs:368-386).
    tag_runs: hand out every tag column as (values, run_lengths) instead of one string per created row (SURVEY 8(f2):
    tags as run-lengths); np.repeat(values, run_lengths) is the column the reference builds.
    limit: like the reference's, batch_size becomes min(limit, batch_size) (grid_exec.rs:239-246).  The reference leaves
    stopping to the LimitExec above it; here the stream also ends after `limit` rows and -- the push-down of SURVEY
    8(f2) -- reconstructs only the leading segments of an input batch that are needed to reach it (when there is no
    predicate, whose selectivity is unknown).  The first `limit` rows are the reference's first `limit` rows.
    """

    def __init__(self, input: Iterable[Tuple[object, Sequence[np.ndarray]]], batch_size: int, n_tag_columns: int = 0,
                 predicate: Optional[Callable[[np.ndarray, np.ndarray], np.ndarray]] = None, ctx: Optional[mc.Context] = None,
                 time_range: Optional[Tuple[Optional[int], Optional[int]]] = None, tag_runs: bool = False,
                 limit: Optional[int] = None, device_time_clip: bool = False):
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        if limit is not None:
            if limit <= 0:
                raise ValueError("limit must be positive")
            batch_size = min(limit, batch_size)
        self.limit = limit
        self._handed_out = 0
        self._input: Iterator = iter(input)
        self._input_done = False
        self.batch_size = batch_size
        self.predicate = predicate
        # (lo, hi), either may be None: segments that end before lo or start after hi are not reconstructed at all --
        # the push-down the reference applies to its Parquet scan (time_series_table.rs:290-373), SURVEY 8(f2).  The
        # caller's predicate must imply the range; points of the surviving segments are still pruned by it.
        self.time_range = time_range
        # With device_time_clip the range is ALSO evaluated point by point inside the grid call (mdbcu_grid_range): of the
        # surviving segments only the points with lo <= timestamp <= hi are reconstructed into the output and copied back,
        # instead of all of them (grid_exec.rs:366-387 prunes after reconstruction).  Since the predicate implies the range
        # the stream's output is unchanged.
        self.device_time_clip = device_time_clip
        self.tag_runs = tag_runs
        self.segments_skipped = 0
        self.ctx = ctx
        self.metrics = GridStreamMetrics()
        self._timestamps = np.zeros(0, np.int64)
        self._values = np.zeros(0, np.float32)
        self._tags = _TagRuns(n_tag_columns)
        self._offset = 0  # current_batch_offset

    def __iter__(self):
        return self

    def _remaining(self) -> int:
        return len(self._timestamps) - self._offset

    def _grid_and_append_to_leftovers_in_current_batch(self, segments, tags: Sequence[np.ndarray]):
        # grid_exec.rs:261-391 -- one batched kernel call instead of one grid() per row
        host = segments.to_host() if isinstance(segments, mc.CompressedSegments) else segments
        if len(tags) != len(self._tags.values):
            raise ValueError("every batch must carry the same tag columns")
        if self.time_range is not None:
            lo, hi = self.time_range
            keep = np.ones(len(host), bool)
            if lo is not None:
                keep &= host.end_time >= lo
            if hi is not None:
                keep &= host.start_time <= hi
            if not keep.all():
                self.segments_skipped += int(len(host) - keep.sum())
                host = host.take(keep)
                tags = [np.asarray(t, object)[keep] for t in tags]
        if len(host) == 0:
            point_off, ts, vals = np.zeros(1, np.uint64), np.zeros(0, np.int64), np.zeros(0, np.float32)
        elif self.device_time_clip and self.time_range is not None:
            lo, hi = self.time_range
            ts, vals, point_off = mc.grid_range(host, -(2 ** 63) if lo is None else lo, 2 ** 63 - 1 if hi is None else hi, self.ctx,
                                                with_point_off=True)
        else:
            point_off, _ = mc.grid_count(host, self.ctx)
            if self.limit is not None and self.predicate is None:
                # rows still owed beyond the leftovers; the first segments that cover them are enough
                owed = self.limit - self._handed_out - self._remaining()
                needed = int(np.searchsorted(point_off[1:], max(owed, 0), "left")) + 1 if owed > 0 else 0
                if needed < len(host):
                    self.segments_skipped += len(host) - needed
                    host = host.slice(0, needed)
                    tags = [np.asarray(t, object)[:needed] for t in tags]
                    point_off = point_off[:needed + 1]
            if len(host) == 0:
                ts, vals = np.zeros(0, np.int64), np.zeros(0, np.float32)
            else:
                ts, vals = mc.grid(host, ctx=self.ctx)
        self.metrics.add_batch(host, point_off)
        lens = np.diff(point_off).astype(np.int64)
        runs = self._tags.tail(self._offset)
        runs.extend(lens, tags)  # each tag value once per created row, as one run
        # (the leftovers were filtered when they were created; the predicate is a per-row test, so filtering them again
        # together with the new points changes nothing)
        ts = np.concatenate([self._timestamps[self._offset:], ts])
        vals = np.concatenate([self._values[self._offset:], vals])
        if self.predicate is not None:
            keep = np.asarray(self.predicate(ts, vals), bool)
            ts, vals = ts[keep], vals[keep]
            runs.filter(keep)
        self._timestamps, self._values, self._tags = ts, vals, runs
        self._offset = 0

    def __next__(self):
        # grid_exec.rs:394-430
        if self.limit is not None and self._handed_out >= self.limit:
            raise StopIteration
        if self._remaining() < self.batch_size and not self._input_done:
            try:
                segments, tags = next(self._input)
                self._grid_and_append_to_leftovers_in_current_batch(segments, tags)
            except StopIteration:
                self._input_done = True
        if self._input_done and self._remaining() == 0:
            raise StopIteration
        length = min(self.batch_size, self._remaining())
        if self.limit is not None:
            length = min(length, self.limit - self._handed_out)
        lo, hi = self._offset, self._offset + length
        self._offset = hi
        self._handed_out += length
        runs = self._tags.slice(lo, hi)
        tags = runs if self.tag_runs else [np.repeat(values, lens) for values, lens in runs]
        return (self._timestamps[lo:hi], self._values[lo:hi], *tags)


class _ModelAccumulator:
    """Accumulator protocol of the reference: update_batch folds a batch of segments into the state, state() returns
    it and resets; merge_batch / evaluate are never called on the model accumulators (`unreachable!()`)."""

    def __init__(self, ctx: Optional[mc.Context] = None):
        self.ctx = ctx

    def merge_batch(self, _states):
        raise RuntimeError("unreachable: model accumulators are only used in Partial aggregates")

    def evaluate(self):
        raise RuntimeError("unreachable: model accumulators are only used in Partial aggregates")

    def _aggregate(self, segments):
        count, mn, mx, sm = mc.aggregate(segments, None, self.ctx)
        return int(count[0]), np.float32(mn[0]), np.float32(mx[0]), float(sm[0])


class ModelCountAccumulator(_ModelAccumulator):  # model_simple_aggregates.rs:337-388
    def __init__(self, ctx=None):
        super().__init__(ctx)
        self.count = 0

    def update_batch(self, segments):
        self.count += self._aggregate(segments)[0]

    def state(self):
        state, self.count = [self.count], 0
        return state


class ModelMinAccumulator(_ModelAccumulator):  # model_simple_aggregates.rs:391-431 (starts at f32::MAX, NaN ignored)
    def __init__(self, ctx=None):
        super().__init__(ctx)
        self.min = F32_MAX

    def update_batch(self, segments):
        batch_min = self._aggregate(segments)[1]
        self.min = batch_min if np.isnan(self.min) else (batch_min if batch_min < self.min else self.min)

    def state(self):
        state, self.min = [self.min], F32_MAX
        return state


class ModelMaxAccumulator(_ModelAccumulator):  # model_simple_aggregates.rs:434-470 (starts at f32::MIN)
    def __init__(self, ctx=None):
        super().__init__(ctx)
        self.max = F32_MIN

    def update_batch(self, segments):
        batch_max = self._aggregate(segments)[2]
        self.max = batch_max if np.isnan(self.max) else (batch_max if batch_max > self.max else self.max)

    def state(self):
        state, self.max = [self.max], F32_MIN
        return state


class ModelSumAccumulator(_ModelAccumulator):  # model_simple_aggregates.rs:473-527 (per-row f32 sums added in f64)
    def __init__(self, ctx=None):
        super().__init__(ctx)
        self.sum = 0.0

    def update_batch(self, segments):
        self.sum += self._aggregate(segments)[3]

    def state(self):
        state, self.sum = [self.sum], 0.0
        return state


class ModelAvgAccumulator(_ModelAccumulator):  # model_simple_aggregates.rs:530-618: state is (count u64, sum f64)
    def __init__(self, ctx=None):
        super().__init__(ctx)
        self.sum = 0.0
        self.count = 0

    def update_batch(self, segments):
        count, _, _, sm = self._aggregate(segments)
        self.sum += sm
        self.count += count

    def state(self):
        state = [self.count, self.sum]
        self.sum, self.count = 0.0, 0
        return state


def grouped_model_aggregates(segments, tag_columns: Sequence[np.ndarray], ctx: Optional[mc.Context] = None):
    """COUNT / MIN / MAX / SUM per distinct combination of tag values, directly from segments (the
    aggregate rule extended to GROUP BY <tag columns>; the reference's rule only rewrites ungrouped aggregates,
    model_simple_aggregates.rs:203-334, and answers grouped ones by reconstructing every point).

    Rows with equal tags are contiguous in what compress and the storage layer produce (one series after the other), so
    each run of equal tags is one group of ONE mdbcu_aggregate call; runs that repeat an earlier key (several files of
    the same series) are merged on the host in row order with the accumulators' folds.  Returns (keys, count i64[],
    min f32[], max f32[], sum f64[]) with keys as tuples in order of first appearance; AVG = sum / count."""
    host = segments.to_host() if isinstance(segments, mc.CompressedSegments) else segments
    n = len(host)
    columns = [np.asarray(c, object) for c in tag_columns]
    if any(len(c) != n for c in columns):
        raise ValueError("a tag column needs one value per segment")
    if n == 0:
        return [], np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float64)
    change = np.zeros(n, bool)
    change[0] = True
    for c in columns:
        change[1:] |= c[1:] != c[:-1]
    starts = np.flatnonzero(change)
    group_off = np.concatenate([starts, [n]]).astype(np.uint64)
    count, mn, mx, sm = mc.aggregate(host, group_off, ctx)
    keys: List[tuple] = []
    index = {}
    slot = np.empty(len(starts), np.int64)
    for r, row in enumerate(starts):
        key = tuple(c[row] for c in columns)
        if key not in index:
            index[key] = len(keys)
            keys.append(key)
        slot[r] = index[key]
    g = len(keys)
    out_count, out_sum = np.zeros(g, np.int64), np.zeros(g, np.float64)
    out_min, out_max = np.full(g, F32_MAX, np.float32), np.full(g, F32_MIN, np.float32)
    if g == len(starts):  # the usual case: every key is one run
        return keys, count.astype(np.int64), mn, mx, sm
    for r in range(len(starts)):
        k = slot[r]
        out_count[k] += count[r]
        out_sum[k] += sm[r]
        if mn[r] < out_min[k]:
            out_min[k] = mn[r]
        if mx[r] > out_max[k]:
            out_max[k] = mx[r]
    return keys, out_count, out_min, out_max, out_sum
