//! `modelardb_cuda`: the FFI crate that binds `libmodelardb_cuda.so` (the B200 implementation of ModelarDB's
//! data-parallel hot path) for ModelarDB-RS.  `ffi` is the `extern "C"` block, one declaration per entry point of
//! `include/modelardb_cuda.h`; the rest of the crate is the thin safe layer the three call sites use:
//!
//! * `compress_units`   replaces the per-series calls of `try_compress_univariate_time_series`
//!   (`crates/modelardb_compression/src/compression.rs:191-275`) with ONE batch call;
//! * `grid_batch`       replaces the per-row loop over `modelardb_compression::grid`
//!   (`crates/modelardb_storage/src/query/grid_exec.rs:323-337`);
//! * `aggregate_batch`  replaces the per-row loops of the five accumulators' `update_batch`
//!   (`crates/modelardb_storage/src/optimizer/model_simple_aggregates.rs:345-585`).
//!
//! `rust/patches/*.diff` apply them.  This crate has NOT been compiled: the repository it ships in has no Rust
//! toolchain (see README.md next to Cargo.toml); the C side is exercised through the same symbols by the Python
//! ctypes binding and a C++ program in that repository's tests.

use std::ffi::{c_char, c_int, c_void, CStr};
use std::sync::Arc;

use arrow::array::{Array, ArrayRef, BinaryViewArray, BinaryViewBuilder, Float32Array, Int8Array};
use arrow::buffer::ScalarBuffer;
use modelardb_types::types::{ErrorBound, TimestampArray, ValueArray};

pub mod ffi {
    use super::*;

    #[repr(C)]
    pub struct MdbcuContext {
        _private: [u8; 0],
    }
    #[repr(C)]
    pub struct MdbcuSegments {
        _private: [u8; 0],
    }
    #[repr(C)]
    pub struct MdbcuComm {
        _private: [u8; 0],
    }

    /// `mdbcu_segments_view`: the columns of QUERY_COMPRESSED_SCHEMA (`crates/modelardb_types/src/schemas.rs:40-52`)
    /// as plain arrays; the three `BinaryView` columns as `u64` offsets (n_segments + 1) + bytes.
    #[repr(C)]
    pub struct MdbcuSegmentsView {
        pub n_segments: u64,
        pub model_type_id: *const i8,
        pub start_time: *const i64,
        pub end_time: *const i64,
        pub min_value: *const f32,
        pub max_value: *const f32,
        pub timestamps_off: *const u64,
        pub timestamps_data: *const u8,
        pub values_off: *const u64,
        pub values_data: *const u8,
        pub residuals_off: *const u64,
        pub residuals_data: *const u8,
    }

    pub const MDBCU_SUCCESS: c_int = 0;
    pub const MDBCU_HOST: c_int = 0;
    pub const MDBCU_DEVICE: c_int = 1;

    #[link(name = "modelardb_cuda")]
    extern "C" {
        // errors, library, contexts
        pub fn mdbcu_last_error() -> *const c_char;
        pub fn mdbcu_device_count() -> c_int;
        pub fn mdbcu_version() -> *const c_char;
        pub fn mdbcu_context_create(device: c_int, out: *mut *mut MdbcuContext) -> c_int;
        pub fn mdbcu_context_destroy(ctx: *mut MdbcuContext);
        pub fn mdbcu_context_set_stream(ctx: *mut MdbcuContext, cuda_stream: *mut c_void) -> c_int;
        pub fn mdbcu_context_stream(ctx: *mut MdbcuContext) -> *mut c_void;
        pub fn mdbcu_context_launch_count(ctx: *const MdbcuContext) -> u64;
        pub fn mdbcu_context_set_profiling(ctx: *mut MdbcuContext, enabled: c_int) -> c_int;
        pub fn mdbcu_context_kernel_stat(ctx: *mut MdbcuContext, index: u32, name: *mut *const c_char,
                                         total_ms: *mut f64, launches: *mut u64) -> c_int;
        // K1 compress
        pub fn mdbcu_compress(ctx: *mut MdbcuContext, space: c_int, timestamps: *const i64, values: *const f32,
                              unit_off: *const u64, n_units: u64, eb_kind: *const u8, eb_value: *const f32,
                              out: *mut *mut MdbcuSegments) -> c_int;
        pub fn mdbcu_segments_len(segments: *const MdbcuSegments) -> u64;
        pub fn mdbcu_segments_get(segments: *mut MdbcuSegments, space: c_int, view: *mut MdbcuSegmentsView,
                                  unit_seg_off: *mut *const u64) -> c_int;
        pub fn mdbcu_segments_free(segments: *mut MdbcuSegments);
        // K2 grid
        pub fn mdbcu_grid_count(ctx: *mut MdbcuContext, space: c_int, segments: *const MdbcuSegmentsView,
                                point_off: *mut u64, total: *mut u64) -> c_int;
        pub fn mdbcu_grid(ctx: *mut MdbcuContext, space: c_int, segments: *const MdbcuSegmentsView,
                          timestamps_out: *mut i64, values_out: *mut f32, capacity: u64, n_points: *mut u64) -> c_int;
        pub fn mdbcu_grid_range(ctx: *mut MdbcuContext, space: c_int, segments: *const MdbcuSegmentsView, t_lo: i64, t_hi: i64,
                                point_off: *mut u64, timestamps_out: *mut i64, values_out: *mut f32, capacity: u64,
                                n_points: *mut u64) -> c_int;
        pub fn mdbcu_sort_rows(ctx: *mut MdbcuContext, space: c_int, series_code: *const u32, timestamps: *const i64, n: u64,
                               order_out: *mut u32) -> c_int;
        pub fn mdbcu_take_rows(ctx: *mut MdbcuContext, space: c_int, order: *const u32, n: u64, timestamps_in: *const i64,
                               timestamps_out: *mut i64, fields_in: *const *const f32, fields_out: *const *mut f32,
                               n_fields: u32) -> c_int;
        // K3 aggregates
        pub fn mdbcu_segment_sums(ctx: *mut MdbcuContext, space: c_int, segments: *const MdbcuSegmentsView,
                                  sums_out: *mut f32) -> c_int;
        pub fn mdbcu_aggregate(ctx: *mut MdbcuContext, space: c_int, segments: *const MdbcuSegmentsView,
                               group_off: *const u64, n_groups: u64, count: *mut i64, min: *mut f32,
                               max: *mut f32, sum: *mut f64) -> c_int;
        // multi-GPU: one communicator per context; only aggregate results travel (one packed ncclAllGather)
        pub fn mdbcu_shard_units(n_units: u64, world: c_int, rank: c_int, lo: *mut u64, hi: *mut u64) -> c_int;
        pub fn mdbcu_comm_unique_id(id128: *mut u8) -> c_int;
        pub fn mdbcu_comm_create(ctx: *mut MdbcuContext, world: c_int, rank: c_int, id128: *const u8,
                                 out: *mut *mut MdbcuComm) -> c_int;
        pub fn mdbcu_comm_create_all(ctxs: *const *mut MdbcuContext, n: c_int, out: *mut *mut MdbcuComm) -> c_int;
        pub fn mdbcu_comm_destroy(comm: *mut MdbcuComm);
        pub fn mdbcu_comm_world(comm: *const MdbcuComm) -> c_int;
        pub fn mdbcu_comm_rank(comm: *const MdbcuComm) -> c_int;
        pub fn mdbcu_aggregate_sharded(comm: *mut MdbcuComm, space: c_int, segments: *const MdbcuSegmentsView,
                                       group_off: *const u64, n_local: u64, n_total: u64, count: *mut i64,
                                       min: *mut f32, max: *mut f32, sum: *mut f64) -> c_int;
        pub fn mdbcu_aggregate_all_sharded(comm: *mut MdbcuComm, space: c_int, segments: *const MdbcuSegmentsView,
                                           count: *mut i64, min: *mut f32, max: *mut f32, sum: *mut f64) -> c_int;
        // tuning and diagnostics; a binding can leave these out (results never depend on them)
        pub fn mdbcu_context_set_option(ctx: *mut MdbcuContext, name: *const c_char, value: i64) -> c_int;
        pub fn mdbcu_context_set_chunk_len(ctx: *mut MdbcuContext, chunk_len: u32) -> c_int;
        pub fn mdbcu_context_set_lane_warmup(ctx: *mut MdbcuContext, points: u32) -> c_int;
        pub fn mdbcu_context_last_compress_rounds(ctx: *const MdbcuContext) -> u32;
        pub fn mdbcu_context_set_fit_engine(ctx: *mut MdbcuContext, engine: c_int) -> c_int;
        pub fn mdbcu_debug_fit_models(ctx: *mut MdbcuContext, timestamps: *const i64, values: *const f32, n: u32,
                                      eb_kind: c_int, eb_value: f32, engine: c_int, starts: *const u32,
                                      budget_ends: *const u32, n_starts: u32, out: *mut c_void) -> c_int;
        pub fn mdbcu_debug_counters(ctx: *mut MdbcuContext, out8: *mut u64) -> c_int;
        pub fn mdbcu_debug_rewrite_position_steps(ctx: *mut MdbcuContext, first_bits: u32, last_bits: u32,
                                                  bits_out: *mut u32, pos_out: *mut i32, cap: u32,
                                                  n_steps: *mut u32) -> c_int;
    }
}

/// What a failed call reports: the library's thread-local message (same convention as modelardb_embedded's C-API,
/// `crates/modelardb_embedded/src/capi.rs:1148-1157`: 0 = ok, else read the last error).
#[derive(Debug)]
pub struct CudaError(pub String);

impl std::fmt::Display for CudaError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "modelardb_cuda: {}", self.0)
    }
}
impl std::error::Error for CudaError {}

pub type Result<T> = std::result::Result<T, CudaError>;

fn check(rc: c_int) -> Result<()> {
    if rc == ffi::MDBCU_SUCCESS {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::mdbcu_last_error()) }.to_string_lossy().into_owned();
    Err(CudaError(msg))
}

/// One CUDA stream on one device.  Calls on a context are blocking and must not overlap: one context per host
/// thread (the compressor thread, each tokio worker that polls a `GridStream`, each accumulator's partition).
pub struct Context {
    raw: *mut ffi::MdbcuContext,
}

// The handle may move between threads; the library serialises nothing, the owner does (one thread at a time).
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::mdbcu_context_create(device, &mut raw) })?;
        Ok(Self { raw })
    }

    pub fn device_count() -> i32 {
        unsafe { ffi::mdbcu_device_count() }
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::mdbcu_context_destroy(self.raw) }
    }
}

/// `ErrorBound` -> (kind, value) as the library takes it (`crates/modelardb_types/src/types.rs:299-335`).
pub fn error_bound_parts(error_bound: ErrorBound) -> (u8, f32) {
    match error_bound {
        ErrorBound::Lossless => (0, 0.0),
        ErrorBound::Absolute(value) => (1, value),
        ErrorBound::Relative(percentage) => (2, percentage),
    }
}

/// The columns of one unit's compressed segments, ready for `CompressedSegmentBatchBuilder`
/// (`crates/modelardb_compression/src/types.rs:411-517`): `error` is filled with NaN and `field_column` / tags with
/// the unit's constants by the caller, exactly as the reference does (`types.rs:492-516`).
pub struct UnitSegments {
    pub model_type_ids: Int8Array,
    pub start_times: TimestampArray,
    pub end_times: TimestampArray,
    pub timestamps: BinaryViewArray,
    pub min_values: ValueArray,
    pub max_values: ValueArray,
    pub values: BinaryViewArray,
    pub residuals: BinaryViewArray,
}

fn binary_views(off: &[u64], data: *const u8, rows: std::ops::Range<usize>) -> BinaryViewArray {
    let mut builder = BinaryViewBuilder::with_capacity(rows.len());
    for row in rows {
        let (a, b) = (off[row] as usize, off[row + 1] as usize);
        builder.append_value(unsafe { std::slice::from_raw_parts(data.add(a), b - a) });
    }
    builder.finish()
}

/// Batch form of `try_compress_univariate_time_series` (`compression.rs:191-275`): unit `u` is
/// `timestamps[unit_off[u]..unit_off[u + 1]]` / `values[..]`, compressed within `error_bounds[u]`.  Whole series for the
/// bulk / embedded path, <= 64 Ki-point buffers for the server path (`uncompressed_data_manager.rs:530-581`): all
/// units of one call share one launch sequence on the GPU.  Returns one `UnitSegments` per unit, in order.
pub fn compress_units(
    ctx: &mut Context,
    timestamps: &[i64],
    values: &[f32],
    unit_off: &[u64],
    error_bounds: &[ErrorBound],
) -> Result<Vec<UnitSegments>> {
    let n_units = unit_off.len().saturating_sub(1);
    if timestamps.len() != values.len() {
        // compression.rs:202-206
        return Err(CudaError("Uncompressed timestamps and uncompressed values have different lengths.".to_owned()));
    }
    if error_bounds.len() != n_units {
        return Err(CudaError("one error bound per unit is required".to_owned()));
    }
    let (kinds, bounds): (Vec<u8>, Vec<f32>) = error_bounds.iter().map(|eb| error_bound_parts(*eb)).unzip();
    let mut segments = std::ptr::null_mut();
    check(unsafe {
        ffi::mdbcu_compress(ctx.raw, ffi::MDBCU_HOST, timestamps.as_ptr(), values.as_ptr(), unit_off.as_ptr(),
                            n_units as u64, kinds.as_ptr(), bounds.as_ptr(), &mut segments)
    })?;
    // The view points into ONE pinned block owned by `segments`: copy out before freeing it.
    let mut view = std::mem::MaybeUninit::<ffi::MdbcuSegmentsView>::uninit();
    let mut unit_seg_off = std::ptr::null();
    let got = check(unsafe { ffi::mdbcu_segments_get(segments, ffi::MDBCU_HOST, view.as_mut_ptr(), &mut unit_seg_off) });
    if let Err(error) = got {
        unsafe { ffi::mdbcu_segments_free(segments) };
        return Err(error);
    }
    let view = unsafe { view.assume_init() };
    let rows = view.n_segments as usize;
    let (uso, ts_off, val_off, res_off) = unsafe {
        (
            std::slice::from_raw_parts(unit_seg_off, n_units + 1),
            std::slice::from_raw_parts(view.timestamps_off, rows + 1),
            std::slice::from_raw_parts(view.values_off, rows + 1),
            std::slice::from_raw_parts(view.residuals_off, rows + 1),
        )
    };
    let column = |ptr: *const u8, size: usize, lo: usize, hi: usize| unsafe {
        std::slice::from_raw_parts(ptr.add(lo * size), (hi - lo) * size).to_vec()
    };
    let mut units = Vec::with_capacity(n_units);
    for u in 0..n_units {
        let (lo, hi) = (uso[u] as usize, uso[u + 1] as usize);
        let i8s: Vec<i8> = column(view.model_type_id as *const u8, 1, lo, hi).into_iter().map(|b| b as i8).collect();
        let as_i64 = |ptr: *const i64| unsafe { std::slice::from_raw_parts(ptr.add(lo), hi - lo).to_vec() };
        let as_f32 = |ptr: *const f32| unsafe { std::slice::from_raw_parts(ptr.add(lo), hi - lo).to_vec() };
        units.push(UnitSegments {
            model_type_ids: Int8Array::from(i8s),
            start_times: TimestampArray::new(ScalarBuffer::from(as_i64(view.start_time)), None),
            end_times: TimestampArray::new(ScalarBuffer::from(as_i64(view.end_time)), None),
            timestamps: binary_views(ts_off, view.timestamps_data, lo..hi),
            min_values: Float32Array::from(as_f32(view.min_value)),
            max_values: Float32Array::from(as_f32(view.max_value)),
            values: binary_views(val_off, view.values_data, lo..hi),
            residuals: binary_views(res_off, view.residuals_data, lo..hi),
        });
    }
    unsafe { ffi::mdbcu_segments_free(segments) };
    Ok(units)
}

/// A batch of segment rows as the library takes it: offsets + data for the three binary columns, built from the
/// `BinaryViewArray`s of a `RecordBatch` in one pass (views are not contiguous in general).
pub struct SegmentColumns<'a> {
    model_type_ids: &'a Int8Array,
    start_times: &'a TimestampArray,
    end_times: &'a TimestampArray,
    min_values: &'a ValueArray,
    max_values: &'a ValueArray,
    binary: [(Vec<u64>, Vec<u8>); 3],
}

impl<'a> SegmentColumns<'a> {
    /// `arrays` in the order the accumulators and `GridStream` already use: model_type_id, start_time, end_time,
    /// timestamps, min_value, max_value, values, residuals (`model_simple_aggregates.rs:347-354`, `grid_exec.rs:282-299`).
    pub fn try_new(arrays: &'a [ArrayRef]) -> Result<Self> {
        fn cast<'b, T: 'static>(arrays: &'b [ArrayRef], index: usize) -> Result<&'b T> {
            arrays
                .get(index)
                .and_then(|a| a.as_any().downcast_ref::<T>())
                .ok_or_else(|| CudaError(format!("column {index} has an unexpected type")))
        }
        let flatten = |array: &BinaryViewArray| {
            let mut off = Vec::with_capacity(array.len() + 1);
            let mut data = Vec::new();
            off.push(0u64);
            for row in 0..array.len() {
                data.extend_from_slice(array.value(row));
                off.push(data.len() as u64);
            }
            (off, data)
        };
        Ok(Self {
            model_type_ids: cast::<Int8Array>(arrays, 0)?,
            start_times: cast::<TimestampArray>(arrays, 1)?,
            end_times: cast::<TimestampArray>(arrays, 2)?,
            min_values: cast::<ValueArray>(arrays, 4)?,
            max_values: cast::<ValueArray>(arrays, 5)?,
            binary: [
                flatten(cast::<BinaryViewArray>(arrays, 3)?),
                flatten(cast::<BinaryViewArray>(arrays, 6)?),
                flatten(cast::<BinaryViewArray>(arrays, 7)?),
            ],
        })
    }

    pub fn len(&self) -> usize {
        self.model_type_ids.len()
    }

    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }

    fn view(&self) -> ffi::MdbcuSegmentsView {
        ffi::MdbcuSegmentsView {
            n_segments: self.len() as u64,
            model_type_id: self.model_type_ids.values().as_ptr(),
            start_time: self.start_times.values().as_ptr(),
            end_time: self.end_times.values().as_ptr(),
            min_value: self.min_values.values().as_ptr(),
            max_value: self.max_values.values().as_ptr(),
            timestamps_off: self.binary[0].0.as_ptr(),
            timestamps_data: self.binary[0].1.as_ptr(),
            values_off: self.binary[1].0.as_ptr(),
            values_data: self.binary[1].1.as_ptr(),
            residuals_off: self.binary[2].0.as_ptr(),
            residuals_data: self.binary[2].1.as_ptr(),
        }
    }
}

/// `grid` over every row of the batch, rows in order (`models/mod.rs:190-251` looped by `grid_exec.rs:323-337`).
/// Returns (timestamps, values, point_off): row `i` created `point_off[i + 1] - point_off[i]` data points, which is
/// what the tag columns are repeated by (`grid_exec.rs:339-346`).
pub fn grid_batch(ctx: &mut Context, segments: &SegmentColumns) -> Result<(Vec<i64>, Vec<f32>, Vec<u64>)> {
    let view = segments.view();
    let mut point_off = vec![0u64; segments.len() + 1];
    let mut total = 0u64;
    check(unsafe { ffi::mdbcu_grid_count(ctx.raw, ffi::MDBCU_HOST, &view, point_off.as_mut_ptr(), &mut total) })?;
    let mut timestamps = vec![0i64; total as usize];
    let mut values = vec![0f32; total as usize];
    let mut n = 0u64;
    check(unsafe {
        ffi::mdbcu_grid(ctx.raw, ffi::MDBCU_HOST, &view, timestamps.as_mut_ptr(), values.as_mut_ptr(), total, &mut n)
    })?;
    Ok((timestamps, values, point_off))
}

/// `grid` with the predicate `t_lo <= timestamp AND timestamp <= t_hi` evaluated inside the call: `GridStream` prunes after
/// reconstructing every data point (`grid_exec.rs:366-387`), here rows outside the range are never reconstructed and only the
/// points inside it come back.  Returns (timestamps, values, exclusive prefix sum of the points returned per row).
pub fn grid_range(ctx: &mut Context, segments: &SegmentColumns, t_lo: i64, t_hi: i64) -> Result<(Vec<i64>, Vec<f32>, Vec<u64>)> {
    let view = segments.view();
    let mut n = 0u64;
    check(unsafe {
        ffi::mdbcu_grid_range(ctx.raw, ffi::MDBCU_HOST, &view, t_lo, t_hi, std::ptr::null_mut(), std::ptr::null_mut(),
                              std::ptr::null_mut(), 0, &mut n) // a count
    })?;
    let mut timestamps = vec![0i64; n as usize];
    let mut values = vec![0f32; n as usize];
    let mut point_off = vec![0u64; segments.len() + 1];
    check(unsafe {
        ffi::mdbcu_grid_range(ctx.raw, ffi::MDBCU_HOST, &view, t_lo, t_hi, point_off.as_mut_ptr(), timestamps.as_mut_ptr(),
                              values.as_mut_ptr(), n, &mut n)
    })?;
    Ok((timestamps, values, point_off))
}

/// COUNT / MIN / MAX / SUM over every row of the batch without reconstructing a data point: what one `update_batch`
/// of the five accumulators folds into its state (`model_simple_aggregates.rs:345-585`).
pub struct BatchAggregates {
    pub count: i64,
    pub min: f32,
    pub max: f32,
    pub sum: f64,
}

pub fn aggregate_batch(ctx: &mut Context, segments: &SegmentColumns) -> Result<BatchAggregates> {
    let view = segments.view();
    let mut out = BatchAggregates { count: 0, min: f32::MAX, max: f32::MIN, sum: 0.0 };
    check(unsafe {
        ffi::mdbcu_aggregate(ctx.raw, ffi::MDBCU_HOST, &view, std::ptr::null(), 1, &mut out.count, &mut out.min,
                             &mut out.max, &mut out.sum)
    })?;
    Ok(out)
}

/// Per-row `sum` (`models/mod.rs:129-184`) for callers that need the f32 row sums themselves.
pub fn segment_sums(ctx: &mut Context, segments: &SegmentColumns) -> Result<Vec<f32>> {
    let view = segments.view();
    let mut sums = vec![0f32; segments.len()];
    check(unsafe { ffi::mdbcu_segment_sums(ctx.raw, ffi::MDBCU_HOST, &view, sums.as_mut_ptr()) })?;
    Ok(sums)
}

thread_local! {
    /// One context per host thread (device 0 unless MODELARDB_CUDA_DEVICE says otherwise), created on first use.
    static THREAD_CONTEXT: std::cell::RefCell<Option<Context>> = const { std::cell::RefCell::new(None) };
}

/// Runs `f` with this thread's context.
pub fn with_thread_context<T>(f: impl FnOnce(&mut Context) -> Result<T>) -> Result<T> {
    THREAD_CONTEXT.with(|cell| {
        let mut slot = cell.borrow_mut();
        if slot.is_none() {
            let device = std::env::var("MODELARDB_CUDA_DEVICE").ok().and_then(|d| d.parse().ok()).unwrap_or(0);
            *slot = Some(Context::new(device)?);
        }
        f(slot.as_mut().expect("context was just created"))
    })
}

/// Keeps `Arc` in scope for callers that build `ArrayRef`s from `UnitSegments`.
pub fn array_ref<A: Array + 'static>(array: A) -> ArrayRef {
    Arc::new(array)
}
