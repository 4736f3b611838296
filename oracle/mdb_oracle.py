"""ctypes binding of the CPU parity oracle (oracle/libmdb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under modelardb_rs_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmdb_oracle.so")

LOSSLESS, ABSOLUTE, RELATIVE = 0, 1, 2
PMC_MEAN, SWING, MACAQUE_V = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile if the .so is missing or stale."""
    src = [os.path.join(_HERE, f) for f in ("mdb_oracle.cc", "mdb_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmdb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Model(C.Structure):
    _fields_ = [
        ("model_type_id", C.c_int8),
        ("start_index", C.c_uint64),
        ("end_index", C.c_uint64),
        ("min_value", C.c_float),
        ("max_value", C.c_float),
        ("values", C.c_uint8 * 8),
        ("values_len", C.c_uint32),
        ("model_last_value", C.c_float),
        ("bytes_per_value", C.c_float),
        ("pmc_len", C.c_uint64),
        ("swing_len", C.c_uint64),
    ]


class _View(C.Structure):
    _fields_ = [
        ("n_segments", C.c_uint64),
        ("model_type_id", C.c_void_p),
        ("start_time", C.c_void_p),
        ("end_time", C.c_void_p),
        ("min_value", C.c_void_p),
        ("max_value", C.c_void_p),
        ("timestamps_off", C.c_void_p),
        ("timestamps_data", C.c_void_p),
        ("values_off", C.c_void_p),
        ("values_data", C.c_void_p),
        ("residuals_off", C.c_void_p),
        ("residuals_data", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, sz, u64, i64, f32, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int64, C.c_float, C.c_int
    L.mdbo_is_value_within_error_bound.argtypes = [i32, f32, f32, f32]
    L.mdbo_is_value_within_error_bound.restype = i32
    L.mdbo_maximum_allowed_deviation.argtypes = [i32, f32, C.c_double]
    L.mdbo_maximum_allowed_deviation.restype = C.c_double
    L.mdbo_compress_residual_timestamps.argtypes = [vp, sz, vp, sz]
    L.mdbo_compress_residual_timestamps.restype = sz
    L.mdbo_decompress_all_timestamps.argtypes = [i64, i64, vp, sz, vp, sz]
    L.mdbo_decompress_all_timestamps.restype = sz
    L.mdbo_len.argtypes = [i64, i64, vp, sz]
    L.mdbo_len.restype = sz
    L.mdbo_macaque_v_compress.argtypes = [i32, f32, vp, sz, i32, f32, vp, sz, vp, vp, vp, vp]
    L.mdbo_macaque_v_compress.restype = sz
    L.mdbo_macaque_v_grid.argtypes = [vp, sz, sz, i32, f32, vp]
    L.mdbo_macaque_v_grid.restype = None
    L.mdbo_macaque_v_sum.argtypes = [vp, sz, sz, i32, f32]
    L.mdbo_macaque_v_sum.restype = f32
    L.mdbo_rewrite_least_mantissa_bits.argtypes = [i32, f32, f32]
    L.mdbo_rewrite_least_mantissa_bits.restype = f32
    L.mdbo_rewrite_position_libm.argtypes = [f32]
    L.mdbo_rewrite_position_libm.restype = C.c_int32
    L.mdbo_rewrite_position_f64.argtypes = [f32]
    L.mdbo_rewrite_position_f64.restype = C.c_int32
    L.mdbo_rewrite_position_steps.argtypes = [i32, C.c_uint32, C.c_uint32, i32, vp, vp, sz]
    L.mdbo_rewrite_position_steps.restype = sz
    L.mdbo_fit_next_model.argtypes = [u64, i32, f32, vp, vp, u64, C.POINTER(_Model)]
    L.mdbo_fit_next_model.restype = None
    L.mdbo_pmc_fit_prefix.argtypes = [i32, f32, vp, u64, vp]
    L.mdbo_pmc_fit_prefix.restype = u64
    L.mdbo_swing_fit_prefix.argtypes = [i32, f32, vp, vp, u64, vp, vp]
    L.mdbo_swing_fit_prefix.restype = u64
    L.mdbo_swing_bounds.argtypes = [i32, f32, vp, vp, u64, vp]
    L.mdbo_swing_bounds.restype = None
    L.mdbo_compress.argtypes = [vp, vp, vp, u64, vp, vp, i32, vp]
    L.mdbo_compress.restype = vp
    L.mdbo_segments_view_get.argtypes = [vp, C.POINTER(_View)]
    L.mdbo_segments_view_get.restype = None
    L.mdbo_segments_free.argtypes = [vp]
    L.mdbo_segments_free.restype = None
    L.mdbo_model_finish.argtypes = [C.POINTER(_Model), i32, f32, u64, vp, vp]
    L.mdbo_model_finish.restype = vp
    L.mdbo_macaque_v_segment.argtypes = [i32, f32, u64, u64, vp, vp]
    L.mdbo_macaque_v_segment.restype = vp
    L.mdbo_decode_values_for_pmc_mean.argtypes = [f32, f32, vp, sz]
    L.mdbo_decode_values_for_pmc_mean.restype = f32
    L.mdbo_decode_values_for_swing.argtypes = [f32, f32, vp, sz, vp, vp]
    L.mdbo_decode_values_for_swing.restype = i32
    L.mdbo_grid_count.argtypes = [C.POINTER(_View), vp, i32]
    L.mdbo_grid_count.restype = u64
    L.mdbo_grid.argtypes = [C.POINTER(_View), vp, vp, u64, i32]
    L.mdbo_grid.restype = u64
    L.mdbo_segment_sums.argtypes = [C.POINTER(_View), vp, i32]
    L.mdbo_segment_sums.restype = None
    L.mdbo_aggregate.argtypes = [C.POINTER(_View), vp, u64, vp, vp, vp, vp, i32]
    L.mdbo_aggregate.restype = None
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        return np.frombuffer(bytes(a), dtype=np.uint8).copy()
    return np.ascontiguousarray(a, dtype=np.uint8)


def error_bound(kind, value=0.0):
    """Accept (kind, value), 'lossless', ('abs', x), ('rel', x)."""
    if isinstance(kind, str):
        kind = {"lossless": LOSSLESS, "abs": ABSOLUTE, "absolute": ABSOLUTE, "rel": RELATIVE, "relative": RELATIVE}[kind]
    return int(kind), float(value)


# --------------------------------------------------------------------------- scalar pieces

def is_value_within_error_bound(eb, real, approx) -> bool:
    k, v = eb
    return bool(lib().mdbo_is_value_within_error_bound(k, v, float(np.float32(real)), float(np.float32(approx))))


def maximum_allowed_deviation(eb, value: float) -> float:
    k, v = eb
    return lib().mdbo_maximum_allowed_deviation(k, v, float(value))


# --------------------------------------------------------------------------- MacaqueTS

def compress_residual_timestamps(ts) -> bytes:
    ts = _i64(ts)
    out = np.empty(16 + 9 * len(ts), dtype=np.uint8)
    n = lib().mdbo_compress_residual_timestamps(_p(ts), len(ts), _p(out), len(out))
    assert n != C.c_size_t(-1).value
    return out[:n].tobytes()


def decompress_all_timestamps(start, end, b: bytes, cap=None) -> np.ndarray:
    bb = _u8(b)
    if cap is None:
        cap = max(8, lib().mdbo_len(start, end, _p(bb), len(bb)) + 8)
    out = np.empty(cap, dtype=np.int64)
    n = lib().mdbo_decompress_all_timestamps(start, end, _p(bb), len(bb), _p(out), cap)
    assert n != C.c_size_t(-1).value
    return out[:n].copy()


def length(start, end, b: bytes) -> int:
    bb = _u8(b)
    return lib().mdbo_len(start, end, _p(bb), len(bb))


# --------------------------------------------------------------------------- MacaqueV

@dataclass
class MacaqueVResult:
    data: bytes
    min_value: np.float32
    max_value: np.float32
    last_leading_zero_bits: int
    last_trailing_zero_bits: int
    last_value: np.float32


def macaque_v_compress(eb, values, seed=None) -> MacaqueVResult:
    k, v = eb
    vals = _f32(values)
    out = np.empty(8 + 6 * len(vals), dtype=np.uint8)
    mn, mx, lv = (np.zeros(1, np.float32) for _ in range(3))
    st = np.zeros(2, np.uint8)
    n = lib().mdbo_macaque_v_compress(
        k, v, _p(vals), len(vals), 0 if seed is None else 1, 0.0 if seed is None else float(np.float32(seed)),
        _p(out), len(out), _p(mn), _p(mx), _p(st), _p(lv))
    assert n != C.c_size_t(-1).value
    return MacaqueVResult(out[:n].tobytes(), mn[0], mx[0], int(st[0]), int(st[1]), lv[0])


def macaque_v_grid(b: bytes, n_values: int, seed=None) -> np.ndarray:
    bb = _u8(b)
    out = np.empty(n_values, dtype=np.float32)
    lib().mdbo_macaque_v_grid(_p(bb), len(bb), n_values, 0 if seed is None else 1,
                              0.0 if seed is None else float(np.float32(seed)), _p(out))
    return out


def macaque_v_sum(b: bytes, n_values: int, seed=None) -> np.float32:
    bb = _u8(b)
    return np.float32(lib().mdbo_macaque_v_sum(_p(bb), len(bb), n_values, 0 if seed is None else 1,
                                               0.0 if seed is None else float(np.float32(seed))))


def rewrite_least_mantissa_bits(eb, value) -> np.float32:
    k, v = eb
    return np.float32(lib().mdbo_rewrite_least_mantissa_bits(k, v, float(np.float32(value))))


def rewrite_position_steps(which: int, first_bits: int = 0, last_bits: int = 0x7F800000, n_threads: int = 0):
    """`23 - floor(|log2(x)|) as i32` (macaque_v.rs:185) as a step function of the f32 bit pattern over
    [first_bits, last_bits]: (bits, position) at every pattern where it changes.  which = 0: libm log2f (the
    reference), 1: f64 log2 rounded once (the form the GPU uses)."""
    n_threads = n_threads or (os.cpu_count() or 1)
    cap = 4096
    bits = np.empty(cap, np.uint32)
    pos = np.empty(cap, np.int32)
    n = lib().mdbo_rewrite_position_steps(which, first_bits, last_bits, n_threads, bits.ctypes.data, pos.ctypes.data, cap)
    assert n <= cap
    return bits[:n].copy(), pos[:n].copy()


# --------------------------------------------------------------------------- models

@dataclass
class Model:
    model_type_id: int
    start_index: int
    end_index: int
    min_value: np.float32
    max_value: np.float32
    values: bytes
    model_last_value: np.float32
    bytes_per_value: np.float32
    pmc_len: int
    swing_len: int
    _c: object = None


def fit_next_model(start_index, eb, ts, values) -> Model:
    k, v = eb
    ts, vals = _i64(ts), _f32(values)
    m = _Model()
    lib().mdbo_fit_next_model(start_index, k, v, _p(ts), _p(vals), len(ts), C.byref(m))
    return Model(m.model_type_id, m.start_index, m.end_index, np.float32(m.min_value), np.float32(m.max_value),
                 bytes(m.values[: m.values_len]), np.float32(m.model_last_value), np.float32(m.bytes_per_value),
                 m.pmc_len, m.swing_len, m)


def pmc_fit_prefix(eb, values):
    k, v = eb
    vals = _f32(values)
    mean = np.zeros(1, np.float32)
    n = lib().mdbo_pmc_fit_prefix(k, v, _p(vals), len(vals), _p(mean))
    return int(n), mean[0]


def swing_fit_prefix(eb, ts, values):
    k, v = eb
    ts, vals = _i64(ts), _f32(values)
    f, l = np.zeros(1, np.float32), np.zeros(1, np.float32)
    n = lib().mdbo_swing_fit_prefix(k, v, _p(ts), _p(vals), len(vals), _p(f), _p(l))
    return int(n), f[0], l[0]


def swing_bounds(eb, ts, values):
    k, v = eb
    ts, vals = _i64(ts), _f32(values)
    out = np.zeros(4, np.float64)
    lib().mdbo_swing_bounds(k, v, _p(ts), _p(vals), len(vals), _p(out))
    return tuple(out)


def decode_values_for_pmc_mean(mn, mx, values: bytes) -> np.float32:
    bb = _u8(values)
    return np.float32(lib().mdbo_decode_values_for_pmc_mean(float(mn), float(mx), _p(bb), len(bb)))


def decode_values_for_swing(mn, mx, values: bytes):
    bb = _u8(values)
    f, l = np.zeros(1, np.float32), np.zeros(1, np.float32)
    rc = lib().mdbo_decode_values_for_swing(float(mn), float(mx), _p(bb), len(bb), _p(f), _p(l))
    if rc != 0:
        raise ValueError("Unknown encoding of swing.")
    return f[0], l[0]


# --------------------------------------------------------------------------- segment batches

_COLS = ("model_type_id", "start_time", "end_time", "min_value", "max_value",
         "timestamps_off", "timestamps_data", "values_off", "values_data", "residuals_off", "residuals_data")


class Segments:
    """A batch of compressed segments as numpy arrays (Arrow LargeBinary-style offsets + data)."""

    def __init__(self, **cols):
        for c in _COLS:
            setattr(self, c, cols[c])
        self.unit_seg_off = cols.get("unit_seg_off")

    def __len__(self):
        return len(self.model_type_id)

    @staticmethod
    def _from_handle(h, unit_seg_off=None) -> "Segments":
        L = lib()
        v = _View()
        L.mdbo_segments_view_get(h, C.byref(v))
        n = v.n_segments

        def arr(ptr, count, dt):
            if count == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,)).copy()

        cols = dict(
            model_type_id=arr(v.model_type_id, n, np.int8),
            start_time=arr(v.start_time, n, np.int64),
            end_time=arr(v.end_time, n, np.int64),
            min_value=arr(v.min_value, n, np.float32),
            max_value=arr(v.max_value, n, np.float32),
            timestamps_off=arr(v.timestamps_off, n + 1, np.uint64),
            values_off=arr(v.values_off, n + 1, np.uint64),
            residuals_off=arr(v.residuals_off, n + 1, np.uint64),
        )
        cols["timestamps_data"] = arr(v.timestamps_data, int(cols["timestamps_off"][-1]), np.uint8)
        cols["values_data"] = arr(v.values_data, int(cols["values_off"][-1]), np.uint8)
        cols["residuals_data"] = arr(v.residuals_data, int(cols["residuals_off"][-1]), np.uint8)
        cols["unit_seg_off"] = unit_seg_off
        L.mdbo_segments_free(h)
        return Segments(**cols)

    def view(self) -> _View:
        v = _View()
        v.n_segments = len(self)
        for c in _COLS:
            setattr(v, c, getattr(self, c).ctypes.data)
        return v

    def row(self, i):
        def sl(off, data):
            return data[int(off[i]): int(off[i + 1])].tobytes()
        return dict(model_type_id=int(self.model_type_id[i]), start_time=int(self.start_time[i]),
                    end_time=int(self.end_time[i]), timestamps=sl(self.timestamps_off, self.timestamps_data),
                    min_value=self.min_value[i], max_value=self.max_value[i],
                    values=sl(self.values_off, self.values_data), residuals=sl(self.residuals_off, self.residuals_data))

    def segment_bytes(self) -> int:
        """B_seg of SURVEY 8(d): sum over rows of 29 + |timestamps| + |values| + |residuals|."""
        return int(29 * len(self) + len(self.timestamps_data) + len(self.values_data) + len(self.residuals_data))


def compress(ts, values, unit_off=None, eb=(LOSSLESS, 0.0), n_threads=1) -> Segments:
    """try_compress_univariate_time_series per unit; eb is one bound or a per-unit list."""
    ts, vals = _i64(ts), _f32(values)
    assert len(ts) == len(vals)
    if unit_off is None:
        unit_off = np.array([0, len(ts)], dtype=np.uint64)
    unit_off = np.ascontiguousarray(unit_off, dtype=np.uint64)
    n_units = len(unit_off) - 1
    if isinstance(eb, tuple):
        eb = [eb] * n_units
    kinds = np.array([e[0] for e in eb], dtype=np.uint8)
    evals = np.array([e[1] for e in eb], dtype=np.float32)
    seg_off = np.zeros(n_units + 1, dtype=np.uint64)
    h = lib().mdbo_compress(_p(ts), _p(vals), _p(unit_off), n_units, _p(kinds), _p(evals), n_threads, _p(seg_off))
    return Segments._from_handle(h, seg_off)


def model_finish(model: Model, eb, residuals_end_index, ts, values) -> Segments:
    k, v = eb
    ts, vals = _i64(ts), _f32(values)
    return Segments._from_handle(lib().mdbo_model_finish(C.byref(model._c), k, v, residuals_end_index, _p(ts), _p(vals)))


def macaque_v_segment(eb, start_index, end_index, ts, values) -> Segments:
    k, v = eb
    ts, vals = _i64(ts), _f32(values)
    return Segments._from_handle(lib().mdbo_macaque_v_segment(k, v, start_index, end_index, _p(ts), _p(vals)))


def grid_count(seg: Segments, n_threads=1) -> np.ndarray:
    off = np.zeros(len(seg) + 1, dtype=np.uint64)
    v = seg.view()
    lib().mdbo_grid_count(C.byref(v), _p(off), n_threads)
    return off


def grid(seg: Segments, n_threads=1):
    off = grid_count(seg, n_threads)
    total = int(off[-1])
    ts = np.empty(total, dtype=np.int64)
    val = np.empty(total, dtype=np.float32)
    v = seg.view()
    n = lib().mdbo_grid(C.byref(v), _p(ts), _p(val), total, n_threads)
    assert n == total
    return ts, val, off


def segment_sums(seg: Segments, n_threads=1) -> np.ndarray:
    out = np.empty(len(seg), dtype=np.float32)
    v = seg.view()
    lib().mdbo_segment_sums(C.byref(v), _p(out), n_threads)
    return out


def aggregate(seg: Segments, group_off=None, n_threads=1):
    """Returns (count i64[G], min f32[G], max f32[G], sum f64[G])."""
    g = 1 if group_off is None else len(group_off) - 1
    count = np.zeros(g, np.int64)
    mn, mx = np.zeros(g, np.float32), np.zeros(g, np.float32)
    sm = np.zeros(g, np.float64)
    v = seg.view()
    go = None if group_off is None else np.ascontiguousarray(group_off, dtype=np.uint64)
    lib().mdbo_aggregate(C.byref(v), None if go is None else _p(go), g, _p(count), _p(mn), _p(mx), _p(sm), n_threads)
    return count, mn, mx, sm
