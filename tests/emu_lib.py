"""Builds and binds tests/emu/emu.cc: the per-thread kernel bodies stepped on the host.

Debugging harness for the GPU-less build container (see emu.cc); the GPU tests are the parity tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import mdb_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emu", "emu.cc")
_SO = os.path.join(_HERE, "emu", "libmdb_emu.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "modelardb_rs_b200", "csrc")
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    deps = [_SRC, os.path.join(_HERE, "emu", "warp_emu.h"), os.path.join(_HERE, "emu", "mdb_host_shim.h")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-I", os.path.join(_HERE, "emu"), "-x", "c++", _SRC, "-o", _SO])
    L = C.CDLL(_SO)
    vp, u64 = C.c_void_p, C.c_uint64
    L.emu_compress.argtypes = [vp, vp, vp, u64, vp, vp, C.c_uint32, vp]
    L.emu_compress.restype = vp
    L.emu_compress_sched.argtypes = [vp, vp, vp, u64, vp, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    L.emu_compress_sched.restype = vp
    L.emu_check_eight_points.argtypes = [vp, vp, C.c_uint32, C.c_uint8, C.c_float]
    L.emu_check_eight_points.restype = u64
    L.emu_set_engine.argtypes = [C.c_int]
    L.emu_set_engine.restype = None
    L.emu_set_lanes.argtypes = [C.c_int]
    L.emu_set_lanes.restype = None
    L.emu_set_lane_warmup.argtypes = [C.c_uint32]
    L.emu_set_lane_warmup.restype = None
    L.emu_set_lane_rounds.argtypes = [C.c_uint32]
    L.emu_set_lane_rounds.restype = None
    L.emu_lane_reruns.restype = u64
    L.emu_lane_counters.argtypes = [vp, vp]
    L.emu_lane_counters.restype = None
    L.emu_division_mismatches.restype = u64
    L.emu_screen_counters.argtypes = [vp, C.c_int]
    L.emu_screen_counters.restype = None
    L.emu_fit_models.argtypes = [vp, vp, C.c_uint32, C.c_uint8, C.c_float, C.c_int, vp, vp, C.c_uint32, vp]
    L.emu_fit_models.restype = None
    L.emu_segments_len.argtypes = [vp]
    L.emu_segments_len.restype = u64
    L.emu_segments_view.argtypes = [vp, C.POINTER(O._View), C.POINTER(vp)]
    L.emu_segments_view.restype = None
    L.emu_segments_free.argtypes = [vp]
    L.emu_grid_count.argtypes = [C.POINTER(O._View), vp]
    L.emu_grid_count.restype = u64
    L.emu_grid.argtypes = [C.POINTER(O._View), vp, vp, u64]
    L.emu_grid.restype = u64
    L.emu_segment_sums.argtypes = [C.POINTER(O._View), vp, vp]
    L.emu_segment_sums.restype = C.c_int
    L.emu_aggregate.argtypes = [C.POINTER(O._View), vp, u64, vp, vp, vp, vp]
    L.emu_aggregate.restype = C.c_int
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_variants = {}


def variant(*defines):
    """A separate build of emu.cc with extra -D flags (e.g. "MDB_FIT_WIDE_ENABLED=1"); emu_fit_models and
    emu_warp_macaque_decode are the entry points used on variants."""
    key = tuple(defines)
    if key not in _variants:
        lib()  # (builds the default library first, so compile errors show up there)
        so = os.path.join(_HERE, "emu", "libmdb_emu_" + "_".join(d.replace("=", "-") for d in defines) + ".so")
        deps = [_SRC, os.path.join(_HERE, "emu", "warp_emu.h"), os.path.join(_HERE, "emu", "mdb_host_shim.h")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                                   "-I", os.path.join(_HERE, "emu")] + ["-D" + d for d in defines] + ["-x", "c++", _SRC, "-o", so])
        L = C.CDLL(so)
        L.emu_fit_models.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.emu_fit_models.restype = None
        L.emu_division_mismatches.restype = C.c_uint64
        _variants[key] = L
    return _variants[key]


FIT_RECORD = np.dtype([("start", np.uint32), ("end", np.uint32), ("min", np.uint32), ("max", np.uint32), ("last", np.uint32),
                       ("bpv", np.uint32), ("type", np.int32), ("vlen", np.int32), ("aborted", np.int32), ("irregular", np.int32)])


def fit_models(ts, values, eb, engine, starts, budget_ends, library=None):
    """fit_next_model at each start on the host: engine 1 = the one-thread code, 2 = the warp-cooperative code of
    mdb_fit_warp.cuh run on 32 fibers (tests/emu/warp_emu.h).  Same records as mdbcu_debug_fit_models."""
    ts = np.ascontiguousarray(ts, np.int64)
    vals = np.ascontiguousarray(values, np.float32)
    starts = np.ascontiguousarray(starts, np.uint32)
    budget_ends = np.ascontiguousarray(budget_ends, np.uint32)
    out = np.zeros(len(starts), FIT_RECORD)
    (library or lib()).emu_fit_models(_p(ts), _p(vals), len(ts), eb[0], eb[1], engine, _p(starts), _p(budget_ends), len(starts), _p(out))
    return out


def division_mismatches() -> int:
    """Emulated fast-path quotients (ddiv_fast*) that differed from the host's a / b so far."""
    return int(lib().emu_division_mismatches())


def screen_counters(clear=True):
    """Event counters of the screened engine (csrc/mdb_fit_screen.cuh): fits, fits the exact engine took, passes, exact point
    evaluations, exact candidates, quiet steps, second rounds (PMC-Mean followed exactly after its bound did not decide)."""
    out = (C.c_uint64 * 8)()
    lib().emu_screen_counters(out, 1 if clear else 0)
    return dict(zip(("fits", "exact_fits", "passes", "exact_points", "exact_candidates", "quiet_steps", "pmc_second_rounds"), list(out)[:7]))


def lane_counters():
    """(chunks run by emulated lanes so far, chunks a lane abandoned because of a non-finite value)."""
    a, b = C.c_uint64(), C.c_uint64()
    lib().emu_lane_counters(C.byref(a), C.byref(b))
    return a.value, b.value


def compress(ts, values, unit_off=None, eb=(0, 0.0), chunk_len=0, rounds=None, sched_seed=0, in_flight=0, engine=1, lanes=False,
             lane_warmup=0, lane_rounds=4) -> O.Segments:
    """sched_seed == 0: the round scheme; otherwise the asynchronous scheduler stepped in a seeded random order with
    `in_flight` concurrent workers (rounds then receives the largest number of chain runs of any unit).
    engine: 1 the one-thread fit, 2 the warp-cooperative fit on 32 fibers.
    lanes (asynchronous scheduler only): every chunk's chain is first run by the one-lane-per-chain engine
    (csrc/mdb_fit_lanes.cuh), the scheduler then only stitches, as mdbcu_compress does by default."""
    lib().emu_set_engine(engine)
    lib().emu_set_lanes(1 if lanes else 0)
    lib().emu_set_lane_warmup(lane_warmup)
    lib().emu_set_lane_rounds(lane_rounds)
    ts = np.ascontiguousarray(ts, np.int64)
    vals = np.ascontiguousarray(values, np.float32)
    if unit_off is None:
        unit_off = np.array([0, len(ts)], np.uint64)
    unit_off = np.ascontiguousarray(unit_off, np.uint64)
    n_units = len(unit_off) - 1
    if isinstance(eb, tuple):
        eb = [eb] * n_units
    kinds = np.array([e[0] for e in eb], np.uint8)
    evals = np.array([e[1] for e in eb], np.float32)
    L = lib()
    r = C.c_uint32(0)
    h = L.emu_compress_sched(_p(ts), _p(vals), _p(unit_off), n_units, _p(kinds), _p(evals), chunk_len, C.byref(r), sched_seed, in_flight)
    if not h:
        raise RuntimeError("emulated chunk-speculative compress did not converge / stalled / row count mismatch")
    if rounds is not None:
        rounds.append(r.value)
    v = O._View()
    uso = C.c_void_p()
    L.emu_segments_view(h, C.byref(v), C.byref(uso))
    n = v.n_segments

    def arr(ptr, count, dt):
        if count == 0 or not ptr:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,)).copy()

    cols = dict(model_type_id=arr(v.model_type_id, n, np.int8), start_time=arr(v.start_time, n, np.int64),
                end_time=arr(v.end_time, n, np.int64), min_value=arr(v.min_value, n, np.float32),
                max_value=arr(v.max_value, n, np.float32), timestamps_off=arr(v.timestamps_off, n + 1, np.uint64),
                values_off=arr(v.values_off, n + 1, np.uint64), residuals_off=arr(v.residuals_off, n + 1, np.uint64))
    cols["timestamps_data"] = arr(v.timestamps_data, int(cols["timestamps_off"][-1]), np.uint8)
    cols["values_data"] = arr(v.values_data, int(cols["values_off"][-1]), np.uint8)
    cols["residuals_data"] = arr(v.residuals_data, int(cols["residuals_off"][-1]), np.uint8)
    cols["unit_seg_off"] = arr(uso.value, n_units + 1, np.uint64)
    L.emu_segments_free(h)
    return O.Segments(**cols)


def grid(seg: O.Segments):
    off = np.zeros(len(seg) + 1, np.uint64)
    v = seg.view()
    total = lib().emu_grid_count(C.byref(v), _p(off))
    if total == 2**64 - 1:
        raise ValueError("malformed segment")
    ts = np.empty(total, np.int64)
    val = np.empty(total, np.float32)
    n = lib().emu_grid(C.byref(v), _p(ts), _p(val), total)
    assert n == total
    return ts, val, off


def segment_sums(seg: O.Segments):
    sums = np.empty(len(seg), np.float32)
    counts = np.empty(len(seg), np.uint64)
    v = seg.view()
    if lib().emu_segment_sums(C.byref(v), _p(sums), _p(counts)) != 0:
        raise ValueError("malformed segment")
    return sums, counts


def aggregate(seg: O.Segments, group_off=None):
    g = 1 if group_off is None else len(group_off) - 1
    count = np.zeros(g, np.int64)
    mn, mx = np.zeros(g, np.float32), np.zeros(g, np.float32)
    sm = np.zeros(g, np.float64)
    v = seg.view()
    go = None if group_off is None else np.ascontiguousarray(group_off, np.uint64)
    if lib().emu_aggregate(C.byref(v), None if go is None else _p(go), g, _p(count), _p(mn), _p(mx), _p(sm)) != 0:
        raise ValueError("malformed segment")
    return count, mn, mx, sm


def check_eight_points(ts, values, eb) -> int:
    """Number of start indices at which fit_reaches_eight_points disagrees with fit_next_model's outcome."""
    ts = np.ascontiguousarray(ts, np.int64)
    vals = np.ascontiguousarray(values, np.float32)
    return int(lib().emu_check_eight_points(_p(ts), _p(vals), len(ts), eb[0], eb[1]))


def warp_macaque_decode(data: bytes, count: int, seed=None, misalign: int = 0, library=None):
    """warp_macaque_v_decode (csrc/mdb_macaque_warp.cuh) run by 32 emulated lanes on one MacaqueV stream placed `misalign`
    bytes past a 16-byte boundary between poisoned bytes; returns (values, last value).  library: a variant() build."""
    raw = np.frombuffer(data, np.uint8)
    backing = np.full(len(raw) + 64, 0xC3, np.uint8)
    start = (-backing.ctypes.data) % 16 + misalign
    backing[start:start + len(raw)] = raw
    stream = backing[start:start + len(raw)]
    out = np.zeros(count, np.float32)
    last = np.zeros(1, np.float32)
    (library or lib()).emu_warp_macaque_decode(_p(stream), C.c_uint64(len(raw)), C.c_uint32(count), 0 if seed is None else 1,
                                  C.c_float(0.0 if seed is None else float(np.float32(seed))), _p(out), _p(last))
    return out, last[0]


def warp_macaque_encode(values, eb=(0, 0.0)):
    """warp_macaque_v_encode (csrc/mdb_macaque_warp.cuh) run by 32 emulated lanes: returns (bytes, min, max, the size the
    counting pass predicted)."""
    vals = np.ascontiguousarray(values, np.float32)
    out = np.full(6 * len(vals) + 8, 0xEE, np.uint8)
    mn, mx = np.zeros(1, np.float32), np.zeros(1, np.float32)
    counted = np.zeros(1, np.uint64)
    L = lib()
    L.emu_warp_macaque_encode.restype = C.c_uint64
    n = L.emu_warp_macaque_encode(C.c_uint8(eb[0]), C.c_float(eb[1]), _p(vals), C.c_uint32(len(vals)), _p(out), _p(mn), _p(mx), _p(counted))
    if n == 2**64 - 1:
        raise AssertionError("the counting and the writing pass disagree on min / max")
    return out[:n].tobytes(), mn[0], mx[0], int(counted[0])
