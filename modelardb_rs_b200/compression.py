"""Host-side mirror of the reference's `modelardb_compression` API over the C-ABI.

Names, argument meaning and error behaviour follow crates/modelardb_compression/src/lib.rs:26-34:
`try_compress_univariate_time_series`, `grid`, `sum`, `len`, `is_value_within_error_bound`'s ErrorBound
type, `MODEL_TYPE_COUNT` / `MODEL_TYPE_NAMES`.  Every function here calls libmodelardb_cuda.so; nothing
is computed in Python.  Inputs may be numpy arrays (host space: the library stages them through the
GPU) or torch CUDA tensors (device space: nothing crosses PCIe).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native
from ._native import DEVICE, HOST, ModelarDbCudaError, SegmentsView

PMC_MEAN_ID, SWING_ID, MACAQUE_V_ID = 0, 1, 2           # models/mod.rs:36-38
MODEL_TYPE_COUNT = 3                                     # models/mod.rs:41
MODEL_TYPE_NAMES = ["pmc_mean", "swing", "macaque_v"]    # models/mod.rs:44
UNCOMPRESSED_DATA_BUFFER_CAPACITY = 64 * 1024            # modelardb_server/src/storage/mod.rs:58

_COLUMNS = ("model_type_id", "start_time", "end_time", "min_value", "max_value", "timestamps_off",
            "timestamps_data", "values_off", "values_data", "residuals_off", "residuals_data")
_DTYPES = dict(model_type_id=np.int8, start_time=np.int64, end_time=np.int64, min_value=np.float32,
               max_value=np.float32, timestamps_off=np.uint64, timestamps_data=np.uint8, values_off=np.uint64,
               values_data=np.uint8, residuals_off=np.uint64, residuals_data=np.uint8)


@dataclass(frozen=True)
class ErrorBound:
    """modelardb_types/src/types.rs:299-335."""
    kind: int
    value: float = 0.0

    @staticmethod
    def lossless() -> "ErrorBound":
        return ErrorBound(0, 0.0)

    @staticmethod
    def try_new_absolute(value: float) -> "ErrorBound":
        if not math.isfinite(value) or value <= 0.0:
            raise ValueError("An absolute error bound must be a positive finite value.")
        return ErrorBound(1, float(value))

    @staticmethod
    def try_new_relative(percentage: float) -> "ErrorBound":
        if not (0.0 < percentage <= 100.0):
            raise ValueError("A relative error bound must be a positive value that is at most 100.0%.")
        return ErrorBound(2, float(percentage))


Lossless = ErrorBound.lossless()


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> int:
    if x is None:
        return 0
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


class Context:
    """One GPU, one stream (mdbcu_context)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _native.check(_native.lib().mdbcu_context_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            _native.lib().mdbcu_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return _native.lib().mdbcu_context_stream(self._h) or 0

    def set_stream(self, cuda_stream: int):
        _native.check(_native.lib().mdbcu_context_set_stream(self._h, C.c_void_p(cuda_stream)))

    @property
    def launch_count(self) -> int:
        return _native.lib().mdbcu_context_launch_count(self._h)

    def set_chunk_len(self, chunk_len: int):
        """Chunk length of the parallel segmentation (0 = automatic); results never depend on it."""
        _native.check(_native.lib().mdbcu_context_set_chunk_len(self._h, chunk_len))

    def set_option(self, name: str, value: int):
        """Tuning knob by name (mdbcu_context_set_option); results never depend on it."""
        _native.check(_native.lib().mdbcu_context_set_option(self._h, name.encode(), int(value)))

    def set_lane_warmup(self, points: int):
        """Points a speculative lane chain starts before its chunk (engine 4); results never depend on it."""
        _native.check(_native.lib().mdbcu_context_set_lane_warmup(self._h, points))

    def set_fit_engine(self, engine: int):
        """0 automatic (5; 3 when no unit has a lossy bound), 1 one thread per chain in rounds, 2 one warp per
        chain in rounds, 3 one warp per chain with the asynchronous scheduler, 4 one lane per chain for the bulk of the
        chains and 3 for the stitching, 5 = 3 with the screened fit (csrc/mdb_fit_screen.cuh); results are identical."""
        _native.check(_native.lib().mdbcu_context_set_fit_engine(self._h, engine))

    @property
    def last_compress_rounds(self) -> int:
        return _native.lib().mdbcu_context_last_compress_rounds(self._h)

    def set_profiling(self, enabled: bool):
        _native.check(_native.lib().mdbcu_context_set_profiling(self._h, 1 if enabled else 0))

    def kernel_stats(self) -> dict:
        """{kernel name: (total device ms, launches)} since profiling was enabled."""
        out = {}
        i = 0
        L = _native.lib()
        while True:
            name, ms, n = C.c_char_p(), C.c_double(), C.c_uint64()
            if L.mdbcu_context_kernel_stat(self._h, i, C.byref(name), C.byref(ms), C.byref(n)) != 0:
                break
            out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out


_default_contexts = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_contexts:
        _default_contexts[device] = Context(device)
    return _default_contexts[device]


class HostSegments:
    """A batch of compressed segments in host memory: the columns of QUERY_COMPRESSED_SCHEMA
    (modelardb_types/src/schemas.rs:40-52) as numpy arrays, binary columns as offsets + data."""

    def __init__(self, unit_seg_off=None, **cols):
        for c in _COLUMNS:
            setattr(self, c, np.ascontiguousarray(cols[c], dtype=_DTYPES[c]))
        self.unit_seg_off = unit_seg_off

    def __len__(self):
        return len(self.model_type_id)

    def view(self) -> SegmentsView:
        v = SegmentsView()
        v.n_segments = len(self)
        for c in _COLUMNS:
            setattr(v, c, getattr(self, c).ctypes.data)
        return v

    space = HOST

    def row(self, i: int) -> dict:
        def sl(off, data):
            return data[int(off[i]): int(off[i + 1])].tobytes()
        return dict(model_type_id=int(self.model_type_id[i]), start_time=int(self.start_time[i]),
                    end_time=int(self.end_time[i]), timestamps=sl(self.timestamps_off, self.timestamps_data),
                    min_value=self.min_value[i], max_value=self.max_value[i],
                    values=sl(self.values_off, self.values_data), residuals=sl(self.residuals_off, self.residuals_data))

    def segment_bytes(self) -> int:
        return int(29 * len(self) + len(self.timestamps_data) + len(self.values_data) + len(self.residuals_data))

    def slice(self, lo: int, hi: int) -> "HostSegments":
        """Rows [lo, hi) as an independent batch (offsets rebased)."""
        cols = {c: getattr(self, c)[lo:hi] for c in ("model_type_id", "start_time", "end_time", "min_value", "max_value")}
        for name in ("timestamps", "values", "residuals"):
            off = getattr(self, name + "_off")
            a, b = int(off[lo]), int(off[hi])
            cols[name + "_off"] = off[lo:hi + 1] - off[lo]
            cols[name + "_data"] = getattr(self, name + "_data")[a:b]
        return HostSegments(**cols)

    def take(self, rows) -> "HostSegments":
        """The given rows (a boolean mask or increasing indices) as an independent batch, in their original order."""
        rows = np.asarray(rows)
        idx = np.flatnonzero(rows) if rows.dtype == bool else rows.astype(np.int64)
        cols = {c: getattr(self, c)[idx] for c in ("model_type_id", "start_time", "end_time", "min_value", "max_value")}
        for name in ("timestamps", "values", "residuals"):
            off = getattr(self, name + "_off").astype(np.int64)
            data = getattr(self, name + "_data")
            lens = off[idx + 1] - off[idx]
            new_off = np.concatenate([[0], np.cumsum(lens)])
            # byte k of the output comes from off[row] + (k - new_off[row]) of the input
            src = np.repeat(off[idx] - new_off[:-1], lens) + np.arange(int(new_off[-1]))
            cols[name + "_off"] = new_off.astype(np.uint64)
            cols[name + "_data"] = data[src] if len(src) else np.zeros(0, np.uint8)
        return HostSegments(**cols)


class DeviceSegments:
    """A caller-assembled batch of compressed segments in device memory: the same columns as HostSegments as
    contiguous CUDA tensors (int8 / int64 / float32 / uint8; the offset columns as int64).  Batches that compress
    produced are CompressedSegments; this is for segment columns that reached the device some other way."""

    space = DEVICE

    def __init__(self, **cols):
        for c in _COLUMNS:
            t = cols[c]
            if not (_is_torch(t) and t.is_cuda and t.is_contiguous()):
                raise ValueError(f"{c}: a contiguous CUDA tensor is required")
            setattr(self, c, t)

    def __len__(self):
        return len(self.model_type_id)

    def view(self) -> SegmentsView:
        v = SegmentsView()
        v.n_segments = len(self)
        for c in _COLUMNS:
            setattr(v, c, getattr(self, c).data_ptr())
        return v


class CompressedSegments:
    """An owned device-resident batch produced by compress (mdbcu_segments)."""

    space = DEVICE

    def __init__(self, handle: C.c_void_p, ctx: Context, n_units: int):
        self._h = handle
        self.ctx = ctx
        self.n_units = n_units

    def __len__(self):
        return _native.lib().mdbcu_segments_len(self._h)

    def free(self):
        if self._h:
            _native.lib().mdbcu_segments_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _get(self, space: int):
        v = SegmentsView()
        uso = C.c_void_p()
        _native.check(_native.lib().mdbcu_segments_get(self._h, space, C.byref(v), C.byref(uso)))
        return v, uso.value

    def view(self) -> SegmentsView:
        return self._get(DEVICE)[0]

    def unit_seg_off_device_ptr(self) -> int:
        return self._get(DEVICE)[1] or 0

    def to_host(self, copy: bool = True) -> HostSegments:
        """The columns in host memory.  copy=False returns views of the library's own (pinned) host copy: they
        stay valid only while this object is alive and un-freed, and passing them back to grid / aggregate
        skips the bounce through the staging ring."""
        v, uso = self._get(HOST)
        n = v.n_segments

        def arr(ptr, count, dt):
            if count == 0 or not ptr:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,))
            return a.copy() if copy else a

        cols = {}
        for c in ("model_type_id", "start_time", "end_time", "min_value", "max_value"):
            cols[c] = arr(getattr(v, c), n, _DTYPES[c])
        for name in ("timestamps", "values", "residuals"):
            off = arr(getattr(v, name + "_off"), n + 1, np.uint64)
            cols[name + "_off"] = off
            cols[name + "_data"] = arr(getattr(v, name + "_data"), int(off[-1]) if len(off) else 0, np.uint8)
        host = HostSegments(unit_seg_off=arr(uso, self.n_units + 1, np.uint64), **cols)
        if not copy:
            host._owner = self
        return host


def _space_of(*arrays) -> int:
    kinds = {_is_torch(a) for a in arrays if a is not None}
    if len(kinds) > 1:
        raise ValueError("mixing numpy (host) and torch (device) arrays in one call")
    if kinds == {True}:
        for a in arrays:
            if a is not None and not a.is_cuda:
                raise ValueError("torch tensors must live on the GPU; pass numpy arrays for host memory")
            if a is not None and not a.is_contiguous():
                raise ValueError("tensors must be contiguous")
        # the context launches on its own non-blocking stream: whatever produced these tensors on
        # torch's current stream must have finished before the library reads them
        import torch
        torch.cuda.current_stream().synchronize()
        return DEVICE
    return HOST


def _bounds(error_bound, n_units: int):
    if isinstance(error_bound, ErrorBound):
        error_bound = [error_bound] * n_units
    if len(error_bound) != n_units:
        raise ValueError("one error bound per unit is required")
    return (np.array([e.kind for e in error_bound], np.uint8), np.array([e.value for e in error_bound], np.float32))


def compress(timestamps, values, unit_off=None, error_bound=Lossless, ctx: Optional[Context] = None) -> CompressedSegments:
    """Batch form of try_compress_univariate_time_series (compression.rs:191-275): unit u is the slice
    [unit_off[u], unit_off[u+1]) and is compressed independently."""
    ctx = ctx or default_context()
    if len(timestamps) != len(values):
        # compression.rs:202-206
        raise ModelarDbCudaError("Uncompressed timestamps and uncompressed values have different lengths.")
    space = _space_of(timestamps, values)
    if space == HOST:
        timestamps = np.ascontiguousarray(timestamps, np.int64)
        values = np.ascontiguousarray(values, np.float32)
        if unit_off is None:
            unit_off = np.array([0, len(timestamps)], np.uint64)
        unit_off = np.ascontiguousarray(unit_off, np.uint64)
        n_units = len(unit_off) - 1
        kinds, ebv = _bounds(error_bound, n_units)
    else:
        import torch
        if timestamps.dtype != torch.int64 or values.dtype != torch.float32:
            raise ValueError("device inputs must be int64 timestamps and float32 values")
        if unit_off is None:
            unit_off = torch.tensor([0, len(timestamps)], dtype=torch.int64, device=timestamps.device)
        n_units = len(unit_off) - 1
        if isinstance(error_bound, tuple) and _is_torch(error_bound[0]):
            kinds, ebv = error_bound  # already on the device
        else:
            k, v = _bounds(error_bound, n_units)
            kinds = torch.from_numpy(k).to(timestamps.device)
            ebv = torch.from_numpy(v).to(timestamps.device)
    out = C.c_void_p()
    _native.check(_native.lib().mdbcu_compress(ctx._h, space, _ptr(timestamps), _ptr(values), _ptr(unit_off), n_units,
                                               _ptr(kinds), _ptr(ebv), C.byref(out)))
    return CompressedSegments(out, ctx, n_units)


def try_compress_univariate_time_series(uncompressed_timestamps, uncompressed_values, error_bound: ErrorBound = Lossless,
                                        ctx: Optional[Context] = None) -> HostSegments:
    """compression.rs:191-275 for one series, returning the segment batch in host memory."""
    seg = compress(uncompressed_timestamps, uncompressed_values, None, error_bound, ctx)
    try:
        return seg.to_host()
    finally:
        seg.free()


def sort_rows(series_code, timestamps, ctx: Optional[Context] = None):
    """The permutation `lexsort_to_indices` yields for (tag tuple code, timestamp) (compression.rs:110-141), computed by a
    stable radix sort on the device (mdbcu_sort_rows); rows with equal keys keep their input order."""
    ctx = ctx or default_context()
    code = np.ascontiguousarray(series_code, np.uint32)
    ts = np.ascontiguousarray(timestamps, np.int64)
    if len(code) != len(ts):
        raise ValueError("one code per row")
    order = np.empty(len(ts), np.uint32)
    _native.check(_native.lib().mdbcu_sort_rows(ctx._h, HOST, _ptr(code), _ptr(ts), len(ts), _ptr(order)))
    return order


def take_rows(order, timestamps, field_columns: Sequence[np.ndarray], ctx: Optional[Context] = None):
    """`take_arrays` of the same function on the device: (timestamps[order], [field[order] for every field column])."""
    ctx = ctx or default_context()
    order = np.ascontiguousarray(order, np.uint32)
    ts = np.ascontiguousarray(timestamps, np.int64)
    fields = [np.ascontiguousarray(f, np.float32) for f in field_columns]
    n = len(order)
    ts_out = np.empty(n, np.int64)
    outs = [np.empty(n, np.float32) for _ in fields]
    ins_p = (C.c_void_p * max(1, len(fields)))(*[f.ctypes.data for f in fields])
    outs_p = (C.c_void_p * max(1, len(fields)))(*[o.ctypes.data for o in outs])
    _native.check(_native.lib().mdbcu_take_rows(ctx._h, HOST, _ptr(order), n, _ptr(ts), _ptr(ts_out), ins_p, outs_p, len(fields)))
    return ts_out, outs


def plan_multivariate(timestamps, tag_columns: Sequence[Sequence[str]], field_columns: Sequence[np.ndarray], device_ctx: Optional[Context] = None):
    """The host part of try_compress_multivariate_time_series (compression.rs:42-141): sort the rows by all tags and then
    time (`sort_time_series_by_tags_and_time`, :110-141), split them into time series where any tag changes (:64-93),
    and lay the (series, field) pairs out as the units of ONE compress call, series-major and field-minor -- the
    order in which the reference appends its RecordBatches (:147-179).

    Returns (unit_timestamps, unit_values, unit_off, unit_series, unit_field, series_tags): unit u is
    unit_*[unit_off[u]:unit_off[u+1]], belongs to series unit_series[u] (tags series_tags[unit_series[u]]) and to field
    column unit_field[u].

    device_ctx: sort and gather on the GPU (mdbcu_sort_rows / mdbcu_take_rows: the tag strings stay here, every row gets the
    dense code of its tag tuple) instead of numpy's lexsort; the result is the same."""
    ts = np.ascontiguousarray(timestamps, np.int64)
    n = len(ts)
    fields = [np.ascontiguousarray(f, np.float32) for f in field_columns]
    if any(len(f) != n for f in fields) or any(len(t) != n for t in tag_columns):
        raise ValueError("all columns must have the same number of rows")
    n_fields = len(fields)
    if n == 0 or n_fields == 0:
        return ts[:0], np.zeros(0, np.float32), np.zeros(1, np.uint64), np.zeros(0, np.int64), np.zeros(0, np.int64), []
    # lexsort by (tag_0, tag_1, ..., timestamp), ascending; tags compare as UTF-8 strings like Arrow's StringView sort
    tag_codes, tag_values = [], []
    for t in tag_columns:
        values, codes = np.unique(np.asarray(t, dtype=str), return_inverse=True)
        tag_codes.append(codes)
        tag_values.append(values)
    if device_ctx is not None:
        if tag_codes:  # the rank of every row's tag tuple among the tuples (np.unique sorts rows lexicographically)
            _, tuple_code = np.unique(np.stack(tag_codes, axis=1), axis=0, return_inverse=True)
            tuple_code = np.asarray(tuple_code).reshape(-1)
        else:
            tuple_code = np.zeros(n, np.int64)
        order = sort_rows(tuple_code.astype(np.uint32), ts, device_ctx)
        ts, fields = take_rows(order, ts, fields, device_ctx)
        tag_codes = [c[order] for c in tag_codes]
        order = None  # (the columns are already in sorted order)
    else:
        order = np.lexsort([ts] + tag_codes[::-1])  # np.lexsort: the LAST key is the primary one
        ts = ts[order]
        tag_codes = [c[order] for c in tag_codes]
    new_series = np.zeros(n, bool)
    new_series[0] = True
    for c in tag_codes:
        new_series[1:] |= c[1:] != c[:-1]
    starts = np.flatnonzero(new_series)
    lens = np.diff(np.append(starts, n))
    n_series = len(starts)
    series_tags = [tuple(str(v[c[a]]) for v, c in zip(tag_values, tag_codes)) for a in starts]
    # units: series-major, field-minor
    unit_len = np.repeat(lens, n_fields)
    unit_off = np.concatenate([[0], np.cumsum(unit_len)]).astype(np.uint64)
    unit_series = np.repeat(np.arange(n_series), n_fields)
    unit_field = np.tile(np.arange(n_fields), n_series)
    total = int(unit_off[-1])
    row = np.repeat(np.repeat(starts, n_fields) - unit_off[:-1].astype(np.int64), unit_len) + np.arange(total)
    field_of_row = np.repeat(unit_field, unit_len)
    field_matrix = np.stack([f if order is None else f[order] for f in fields])
    return ts[row], field_matrix[field_of_row, row], unit_off, unit_series, unit_field, series_tags


def try_compress_multivariate_time_series(timestamps, tag_columns: Sequence[Sequence[str]], field_columns: Sequence[np.ndarray],
                                          error_bounds: Sequence[ErrorBound], ctx: Optional[Context] = None, device_sort: bool = True):
    """compression.rs:42-107 with the per-series loop replaced by one batched kernel call: returns, in the reference's
    order, one (tag_values, field_column_index, HostSegments) triple per (time series, field column)."""
    if len(error_bounds) != len(field_columns):
        raise ValueError("one error bound per field column")
    u_ts, u_val, unit_off, unit_series, unit_field, series_tags = plan_multivariate(
        timestamps, tag_columns, field_columns, device_ctx=(ctx or default_context()) if device_sort else None)
    n_units = len(unit_series)
    if n_units == 0:
        return []
    seg = compress(u_ts, u_val, unit_off, [error_bounds[f] for f in unit_field], ctx)
    try:
        host = seg.to_host()
    finally:
        seg.free()
    uso = host.unit_seg_off
    return [(series_tags[unit_series[u]], int(unit_field[u]), host.slice(int(uso[u]), int(uso[u + 1]))) for u in range(n_units)]


def split_into_buffers(n_points_per_series: Sequence[int], capacity: int = UNCOMPRESSED_DATA_BUFFER_CAPACITY) -> np.ndarray:
    """unit_off for the server ingestion path: every series is cut into buffers of at most `capacity`
    points before compression (modelardb_server/src/storage/mod.rs:58,
    uncompressed_data_manager.rs:530-581)."""
    offs = [0]
    pos = 0
    for n in n_points_per_series:
        end = pos + int(n)
        while pos < end:
            pos = min(end, pos + capacity)
            offs.append(pos)
    return np.array(offs, np.uint64)


def _view_of(segments):
    return segments.view(), segments.space


def _order_streams(space: int):
    """Device space: the context launches on its own non-blocking stream, while the outputs were just taken from
    torch's caching allocator (which may hand out a block that a kernel still queued on torch's stream uses) and
    caller tensors may still be being written there: let torch's current stream drain before the library runs."""
    if space == DEVICE:
        import torch
        torch.cuda.current_stream().synchronize()


def grid_count(segments, ctx: Optional[Context] = None):
    """len() per row (models/mod.rs:98-124) as an exclusive prefix sum: returns (point_off, total)."""
    ctx = ctx or getattr(segments, "ctx", None) or default_context()
    v, space = _view_of(segments)
    total = C.c_uint64()
    if space == HOST:
        off = np.zeros(v.n_segments + 1, np.uint64)
    else:
        import torch
        # torch.empty, never torch.zeros: a fill kernel on torch's stream is unordered w.r.t. the context's stream
        off = torch.empty(v.n_segments + 1, dtype=torch.int64, device=f"cuda:{ctx.device}")
    _order_streams(space)
    _native.check(_native.lib().mdbcu_grid_count(ctx._h, space, C.byref(v), _ptr(off), C.byref(total)))
    return off, total.value


def grid(segments, timestamps_out=None, values_out=None, ctx: Optional[Context] = None):
    """grid() over every row of the batch, rows in order (models/mod.rs:190-251, grid_exec.rs:323-337).
    Returns (timestamps, values) trimmed to the number of data points."""
    ctx = ctx or getattr(segments, "ctx", None) or default_context()
    v, space = _view_of(segments)
    if timestamps_out is None:
        _, total = grid_count(segments, ctx)
        if space == HOST:
            timestamps_out = np.empty(total, np.int64)
            values_out = np.empty(total, np.float32)
        else:
            import torch
            dev = f"cuda:{ctx.device}"
            timestamps_out = torch.empty(total, dtype=torch.int64, device=dev)
            values_out = torch.empty(total, dtype=torch.float32, device=dev)
    n = C.c_uint64()
    _order_streams(space)
    _native.check(_native.lib().mdbcu_grid(ctx._h, space, C.byref(v), _ptr(timestamps_out), _ptr(values_out),
                                           len(timestamps_out), C.byref(n)))
    return timestamps_out[: n.value], values_out[: n.value]


def grid_range(segments, t_lo: int, t_hi: int, ctx: Optional[Context] = None, with_point_off: bool = False):
    """grid() with the time predicate `t_lo <= timestamp <= t_hi` evaluated on the device (mdbcu_grid_range): rows outside
    the range are never reconstructed, of the others only the points inside it come back.  The reference prunes after
    reconstructing everything (grid_exec.rs:366-387).  Returns (timestamps, values[, point_off]); point_off is the
    exclusive prefix sum of the points returned per row of the batch."""
    ctx = ctx or getattr(segments, "ctx", None) or default_context()
    v, space = _view_of(segments)
    # capacity: the points of the rows that meet the range (len() per row is cheap: no point is reconstructed for it)
    row_off, total = grid_count(segments, ctx)
    if hasattr(segments, "start_time") and v.n_segments:
        keep = (segments.end_time >= t_lo) & (segments.start_time <= t_hi)
        bound = int((row_off[1:] - row_off[:-1])[keep].sum())
    else:  # (an owned batch: its columns are only reachable through the view)
        bound = int(total)
    if space == HOST:
        ts_out, val_out = np.empty(bound, np.int64), np.empty(bound, np.float32)
        point_off = np.empty(v.n_segments + 1, np.uint64) if with_point_off else None
    else:
        import torch
        dev = f"cuda:{ctx.device}"
        ts_out = torch.empty(bound, dtype=torch.int64, device=dev)
        val_out = torch.empty(bound, dtype=torch.float32, device=dev)
        point_off = torch.empty(v.n_segments + 1, dtype=torch.int64, device=dev) if with_point_off else None
    n = C.c_uint64()
    _order_streams(space)
    if bound == 0:  # nothing meets the range: the call still fills point_off
        _native.check(_native.lib().mdbcu_grid_range(ctx._h, space, C.byref(v), int(t_lo), int(t_hi), _ptr(point_off), None, None, 0, C.byref(n)))
    else:
        _native.check(_native.lib().mdbcu_grid_range(ctx._h, space, C.byref(v), int(t_lo), int(t_hi), _ptr(point_off), _ptr(ts_out), _ptr(val_out),
                                                     bound, C.byref(n)))
    if with_point_off:
        return ts_out[: n.value], val_out[: n.value], point_off
    return ts_out[: n.value], val_out[: n.value]


def segment_sums(segments, ctx: Optional[Context] = None):
    """sum() per row (models/mod.rs:129-184)."""
    ctx = ctx or getattr(segments, "ctx", None) or default_context()
    v, space = _view_of(segments)
    if space == HOST:
        out = np.empty(v.n_segments, np.float32)
    else:
        import torch
        out = torch.empty(v.n_segments, dtype=torch.float32, device=f"cuda:{ctx.device}")
    _order_streams(space)
    _native.check(_native.lib().mdbcu_segment_sums(ctx._h, space, C.byref(v), _ptr(out)))
    return out


def aggregate(segments, group_off=None, ctx: Optional[Context] = None):
    """COUNT / MIN / MAX / SUM per group of rows without materialising data points
    (model_simple_aggregates.rs:345-585). Returns (count i64, min f32, max f32, sum f64)."""
    ctx = ctx or getattr(segments, "ctx", None) or default_context()
    v, space = _view_of(segments)
    if isinstance(group_off, tuple):  # (raw pointer in the batch's memory space, n_groups)
        group_off, g = group_off
    else:
        g = 1 if group_off is None else len(group_off) - 1
    if space == HOST:
        if group_off is not None and not isinstance(group_off, int):
            group_off = np.ascontiguousarray(group_off, np.uint64)
        count, mn, mx, sm = np.zeros(g, np.int64), np.zeros(g, np.float32), np.zeros(g, np.float32), np.zeros(g, np.float64)
    else:
        import torch
        dev = f"cuda:{ctx.device}"
        count = torch.empty(g, dtype=torch.int64, device=dev)
        mn = torch.empty(g, dtype=torch.float32, device=dev)
        mx = torch.empty(g, dtype=torch.float32, device=dev)
        sm = torch.empty(g, dtype=torch.float64, device=dev)
    gp = group_off if isinstance(group_off, int) else _ptr(group_off)
    _order_streams(space)
    _native.check(_native.lib().mdbcu_aggregate(ctx._h, space, C.byref(v), gp, g if group_off is not None else 1,
                                                _ptr(count), _ptr(mn), _ptr(mx), _ptr(sm)))
    return count, mn, mx, sm


# Row-wise forms with the reference's names and argument order (models/mod.rs:98, :129-138, :190-201).
def _one_row(model_type_id, start_time, end_time, timestamps, min_value, max_value, values, residuals) -> HostSegments:
    u8 = lambda b: np.frombuffer(bytes(b), np.uint8).copy()
    return HostSegments(
        model_type_id=np.array([model_type_id], np.int8), start_time=np.array([start_time], np.int64),
        end_time=np.array([end_time], np.int64), min_value=np.array([min_value], np.float32),
        max_value=np.array([max_value], np.float32),
        timestamps_off=np.array([0, len(timestamps)], np.uint64), timestamps_data=u8(timestamps),
        values_off=np.array([0, len(values)], np.uint64), values_data=u8(values),
        residuals_off=np.array([0, len(residuals)], np.uint64), residuals_data=u8(residuals))


def len_(start_time: int, end_time: int, timestamps: bytes, ctx: Optional[Context] = None) -> int:
    """models/mod.rs:98-124."""
    row = _one_row(PMC_MEAN_ID, start_time, end_time, timestamps, 0.0, 0.0, b"", b"")
    return grid_count(row, ctx)[1]


def sum_(model_type_id, start_time, end_time, timestamps, min_value, max_value, values, residuals,
         ctx: Optional[Context] = None) -> np.float32:
    """models/mod.rs:129-184."""
    row = _one_row(model_type_id, start_time, end_time, timestamps, min_value, max_value, values, residuals)
    return segment_sums(row, ctx)[0]


def grid_row(model_type_id, start_time, end_time, timestamps, min_value, max_value, values, residuals,
             ctx: Optional[Context] = None):
    """models/mod.rs:190-251 for a single row: returns (timestamps, values)."""
    row = _one_row(model_type_id, start_time, end_time, timestamps, min_value, max_value, values, residuals)
    return grid(row, ctx=ctx)
