"""ctypes binding of libmodelardb_cuda.so (include/modelardb_cuda.h).

There is no fallback: if the library is missing, or there is no CUDA device, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MODELARDB_CUDA_LIB: another build of the same library (a tuning variant made by build.py), never another implementation
LIB_PATH = os.environ.get("MODELARDB_CUDA_LIB") or os.path.join(_HERE, "libmodelardb_cuda.so")

HOST, DEVICE = 0, 1

# Every symbol include/modelardb_cuda.h declares (tests check the .so exports exactly these).
SYMBOLS = (
    "mdbcu_last_error", "mdbcu_device_count", "mdbcu_version",
    "mdbcu_context_create", "mdbcu_context_destroy", "mdbcu_context_set_stream", "mdbcu_context_stream",
    "mdbcu_context_launch_count", "mdbcu_context_set_profiling", "mdbcu_context_kernel_stat",
    "mdbcu_context_set_chunk_len", "mdbcu_context_set_lane_warmup", "mdbcu_context_set_option", "mdbcu_context_last_compress_rounds", "mdbcu_context_set_fit_engine", "mdbcu_debug_fit_models", "mdbcu_debug_counters",
    "mdbcu_debug_rewrite_position_steps",
    "mdbcu_compress", "mdbcu_segments_len", "mdbcu_segments_get", "mdbcu_segments_free",
    "mdbcu_grid_count", "mdbcu_grid", "mdbcu_grid_range", "mdbcu_sort_rows", "mdbcu_take_rows", "mdbcu_segment_sums", "mdbcu_aggregate",
    "mdbcu_shard_units", "mdbcu_comm_unique_id", "mdbcu_comm_create", "mdbcu_comm_create_all", "mdbcu_comm_destroy",
    "mdbcu_comm_world", "mdbcu_comm_rank", "mdbcu_aggregate_sharded", "mdbcu_aggregate_all_sharded",
)


class ModelarDbCudaError(RuntimeError):
    """A call through the C-ABI returned MDBCU_FAILURE; the message is mdbcu_last_error()."""


class SegmentsView(C.Structure):
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
    if not os.path.exists(LIB_PATH):
        raise ModelarDbCudaError(
            f"{LIB_PATH} is missing: build it with `python -m modelardb_rs_b200.build` "
            "(there is no CPU implementation to fall back to)")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.mdbcu_last_error.restype = C.c_char_p
    L.mdbcu_version.restype = C.c_char_p
    L.mdbcu_device_count.restype = i32
    L.mdbcu_context_create.argtypes = [i32, C.POINTER(vp)]
    L.mdbcu_context_create.restype = i32
    L.mdbcu_context_destroy.argtypes = [vp]
    L.mdbcu_context_destroy.restype = None
    L.mdbcu_context_set_stream.argtypes = [vp, vp]
    L.mdbcu_context_set_stream.restype = i32
    L.mdbcu_context_stream.argtypes = [vp]
    L.mdbcu_context_stream.restype = vp
    L.mdbcu_context_launch_count.argtypes = [vp]
    L.mdbcu_context_launch_count.restype = u64
    L.mdbcu_context_set_profiling.argtypes = [vp, i32]
    L.mdbcu_context_set_profiling.restype = i32
    L.mdbcu_context_kernel_stat.argtypes = [vp, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(u64)]
    L.mdbcu_context_kernel_stat.restype = i32
    L.mdbcu_context_set_chunk_len.argtypes = [vp, C.c_uint32]
    L.mdbcu_context_set_chunk_len.restype = i32
    L.mdbcu_context_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.mdbcu_context_set_option.restype = i32
    L.mdbcu_context_set_lane_warmup.argtypes = [vp, C.c_uint32]
    L.mdbcu_context_set_lane_warmup.restype = i32
    L.mdbcu_context_last_compress_rounds.argtypes = [vp]
    L.mdbcu_context_last_compress_rounds.restype = C.c_uint32
    L.mdbcu_context_set_fit_engine.argtypes = [vp, i32]
    L.mdbcu_context_set_fit_engine.restype = i32
    L.mdbcu_debug_fit_models.argtypes = [vp, vp, vp, C.c_uint32, i32, C.c_float, i32, vp, vp, C.c_uint32, vp]
    L.mdbcu_debug_fit_models.restype = i32
    L.mdbcu_debug_counters.argtypes = [vp, vp]
    L.mdbcu_debug_counters.restype = i32
    L.mdbcu_debug_rewrite_position_steps.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, vp]
    L.mdbcu_debug_rewrite_position_steps.restype = i32
    L.mdbcu_compress.argtypes = [vp, i32, vp, vp, vp, u64, vp, vp, C.POINTER(vp)]
    L.mdbcu_compress.restype = i32
    L.mdbcu_segments_len.argtypes = [vp]
    L.mdbcu_segments_len.restype = u64
    L.mdbcu_segments_get.argtypes = [vp, i32, C.POINTER(SegmentsView), C.POINTER(vp)]
    L.mdbcu_segments_get.restype = i32
    L.mdbcu_segments_free.argtypes = [vp]
    L.mdbcu_segments_free.restype = None
    L.mdbcu_grid_count.argtypes = [vp, i32, C.POINTER(SegmentsView), vp, C.POINTER(u64)]
    L.mdbcu_grid_count.restype = i32
    L.mdbcu_grid.argtypes = [vp, i32, C.POINTER(SegmentsView), vp, vp, u64, C.POINTER(u64)]
    L.mdbcu_grid.restype = i32
    L.mdbcu_grid_range.argtypes = [vp, i32, C.POINTER(SegmentsView), C.c_int64, C.c_int64, vp, vp, vp, u64, C.POINTER(u64)]
    L.mdbcu_grid_range.restype = i32
    L.mdbcu_sort_rows.argtypes = [vp, i32, vp, vp, u64, vp]
    L.mdbcu_sort_rows.restype = i32
    L.mdbcu_take_rows.argtypes = [vp, i32, vp, u64, vp, vp, vp, vp, C.c_uint32]
    L.mdbcu_take_rows.restype = i32
    L.mdbcu_segment_sums.argtypes = [vp, i32, C.POINTER(SegmentsView), vp]
    L.mdbcu_segment_sums.restype = i32
    L.mdbcu_aggregate.argtypes = [vp, i32, C.POINTER(SegmentsView), vp, u64, vp, vp, vp, vp]
    L.mdbcu_aggregate.restype = i32
    L.mdbcu_shard_units.argtypes = [u64, i32, i32, C.POINTER(u64), C.POINTER(u64)]
    L.mdbcu_shard_units.restype = i32
    L.mdbcu_comm_unique_id.argtypes = [vp]
    L.mdbcu_comm_unique_id.restype = i32
    L.mdbcu_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.mdbcu_comm_create.restype = i32
    L.mdbcu_comm_create_all.argtypes = [vp, i32, vp]
    L.mdbcu_comm_create_all.restype = i32
    L.mdbcu_comm_destroy.argtypes = [vp]
    L.mdbcu_comm_destroy.restype = None
    L.mdbcu_comm_world.argtypes = [vp]
    L.mdbcu_comm_world.restype = i32
    L.mdbcu_comm_rank.argtypes = [vp]
    L.mdbcu_comm_rank.restype = i32
    L.mdbcu_aggregate_sharded.argtypes = [vp, i32, C.POINTER(SegmentsView), vp, u64, u64, vp, vp, vp, vp]
    L.mdbcu_aggregate_sharded.restype = i32
    L.mdbcu_aggregate_all_sharded.argtypes = [vp, i32, C.POINTER(SegmentsView), vp, vp, vp, vp]
    L.mdbcu_aggregate_all_sharded.restype = i32
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise ModelarDbCudaError(lib().mdbcu_last_error().decode("utf-8", "replace"))
