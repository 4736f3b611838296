/* modelardb_cuda.h -- C-ABI of the B200-native ModelarDB hot path (libmodelardb_cuda.so).
 *
 * This is the boundary a `modelardb_cuda` FFI crate binds (see INTEGRATION.md for the Rust side).
 * Each entry point is the BATCH form of one function of the reference's `modelardb_compression`
 * API (crates/modelardb_compression/src/lib.rs:26-34) or of one operator loop that calls it:
 *
 *   mdbcu_compress      <- try_compress_univariate_time_series   compression.rs:191-275, called per
 *                          (series, field) by try_split_and_compress_univariate_time_series
 *                          (compression.rs:147-179) and by the server's compressor thread
 *                          (crates/modelardb_server/src/storage/uncompressed_data_manager.rs:563-581)
 *   mdbcu_grid_count    <- len                                    models/mod.rs:98-124
 *   mdbcu_grid          <- grid, looped over the rows of a batch  models/mod.rs:190-251,
 *                          crates/modelardb_storage/src/query/grid_exec.rs:323-337
 *   mdbcu_segment_sums  <- sum                                    models/mod.rs:129-184
 *   mdbcu_aggregate     <- Model{Count,Min,Max,Sum,Avg}Accumulator::update_batch
 *                          crates/modelardb_storage/src/optimizer/model_simple_aggregates.rs:345-585
 *
 * Conventions follow the reference's own C-API (crates/modelardb_embedded/src/capi.rs:56-80,
 * 1148-1157; bindings/c/modelardb_embedded.h:73-199): every call returns 0 on success and 1 on
 * failure, the message of the last failure on the calling thread is returned by
 * mdbcu_last_error(), outputs are written through caller-supplied pointers, and memory produced
 * by the library is released by a library call.  Where the reference would panic on a malformed
 * segment row (models/mod.rs:170, :237; types.rs:315-319, :391, :405) these calls fail instead.
 *
 * Calls are blocking.  Every array argument lives in the memory space named by `space`:
 * MDBCU_HOST (pageable or pinned host memory; the library stages it through the device) or
 * MDBCU_DEVICE (device memory of the context's GPU; nothing crosses PCIe).  Scalars returned
 * through pointers (totals) are always host memory.
 *
 * There is no CPU implementation behind this header: without a CUDA device every compute call
 * fails with "no CUDA device".
 */
#ifndef MODELARDB_CUDA_H
#define MODELARDB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDBCU_SUCCESS 0
#define MDBCU_FAILURE 1

/* ErrorBound (crates/modelardb_types/src/types.rs:299-335). */
#define MDBCU_LOSSLESS 0
#define MDBCU_ABSOLUTE 1 /* value > 0, finite */
#define MDBCU_RELATIVE 2 /* 0 < percent <= 100 */

/* Model type ids (crates/modelardb_compression/src/models/mod.rs:36-38). */
#define MDBCU_PMC_MEAN 0
#define MDBCU_SWING 1
#define MDBCU_MACAQUE_V 2

typedef enum { MDBCU_HOST = 0, MDBCU_DEVICE = 1 } mdbcu_space;

typedef struct mdbcu_context mdbcu_context;   /* one GPU, one stream, reusable scratch */
typedef struct mdbcu_segments mdbcu_segments; /* an owned batch of compressed segments */

/* A batch of compressed segments: the columns of QUERY_COMPRESSED_SCHEMA
 * (crates/modelardb_types/src/schemas.rs:40-52) as plain arrays.  The three BinaryView columns are
 * Arrow LargeBinary style: row i is data[off[i] .. off[i+1]].  The `error` column is constant NaN
 * (compression.rs:398, types.rs:265) and `field_column` / tags are constants per compress call
 * (types.rs:492-516), so neither is carried. */
typedef struct {
    uint64_t n_segments;
    const int8_t *model_type_id;
    const int64_t *start_time;
    const int64_t *end_time;
    const float *min_value;
    const float *max_value;
    const uint64_t *timestamps_off; /* n_segments + 1 */
    const uint8_t *timestamps_data;
    const uint64_t *values_off;     /* n_segments + 1 */
    const uint8_t *values_data;
    const uint64_t *residuals_off;  /* n_segments + 1 */
    const uint8_t *residuals_data;
} mdbcu_segments_view;

/* ---- library state ------------------------------------------------------------------------- */

/* Message of the last failed call on this thread; valid until the next call on this thread. */
const char *mdbcu_last_error(void);
/* Number of visible CUDA devices (0 without a driver/GPU; never fails). */
int mdbcu_device_count(void);
/* Library version as "major.minor.patch". */
const char *mdbcu_version(void);

/* A context is one GPU and one CUDA stream plus the library's staging memory for it (a pinned bounce ring
 * for pageable caller memory, cached pinned blocks for host copies of segments, a mapped mailbox for
 * scalar read-backs).  Calls on one context are blocking and must not overlap: use one context per
 * host thread; contexts on different threads run concurrently on the device.  Free every
 * mdbcu_segments created on a context before destroying it.
 * Tuning knob read when a context is created: MDBCU_LANE_ROWS_MIN (environment) = the number of long MacaqueV rows in a
 * batch from which grid / aggregate decode one row per thread instead of one row per warp (default 24 576; results
 * are identical either way). */
int mdbcu_context_create(int device, mdbcu_context **out);
void mdbcu_context_destroy(mdbcu_context *ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int mdbcu_context_set_stream(mdbcu_context *ctx, void *cuda_stream);
/* The cudaStream_t the context launches on (for CUDA-event timing by the caller). */
void *mdbcu_context_stream(mdbcu_context *ctx);
/* Number of kernels this context has launched since it was created. */
uint64_t mdbcu_context_launch_count(const mdbcu_context *ctx);
/* Per-kernel device time: while enabled every launch is bracketed by CUDA events on the context's
 * stream (enabling also clears the table).  This is the analogue of the EXPLAIN ANALYZE counters
 * GridExec keeps (crates/modelardb_storage/src/query/grid_exec.rs:441-519). */
int mdbcu_context_set_profiling(mdbcu_context *ctx, int enabled);
/* Entry `index` of the table: kernel name, summed device milliseconds, number of launches.
 * Fails when index is past the end (iterate from 0 until failure). */
int mdbcu_context_kernel_stat(mdbcu_context *ctx, uint32_t index, const char **name, double *total_ms,
                              uint64_t *launches);

/* ---- K1: compress -------------------------------------------------------------------------- */

/* Compress n_units independent sorted (timestamps, values) slices; unit u is
 * [unit_off[u], unit_off[u+1]) of `timestamps` / `values` and is compressed within the bound
 * (eb_kind[u], eb_value[u]).  Each unit is one try_compress_univariate_time_series call: whole
 * series for the bulk / embedded path, <= 65 536-point buffers for the server path.  An empty unit
 * yields no segments (compression.rs:208-211).  Segment rows are ordered by unit, then time.
 * Fails on invalid bounds (types.rs:312-334) or non-monotone unit_off. */
int mdbcu_compress(mdbcu_context *ctx, mdbcu_space space, const int64_t *timestamps,
                   const float *values, const uint64_t *unit_off, uint64_t n_units,
                   const uint8_t *eb_kind, const float *eb_value, mdbcu_segments **out);

/* Tuning / diagnostics of the parallel segmentation.  Every unit is cut into chunks that run their
 * greedy chains concurrently and are stitched by a fixpoint that reproduces the sequential chain
 * exactly (csrc/mdb_compress.cuh); results never depend on the chunk length.  0 = automatic. */
int mdbcu_context_set_chunk_len(mdbcu_context *ctx, uint32_t chunk_len);
/* Tuning knobs by name, for tests and benchmarks (results never depend on them): "chunk_len", "fit_engine",
 * "lane_warmup", "lane_rounds_by_lanes", "lane_rows_min" (long MacaqueV rows per batch from which one thread decodes a
 * row), "block_row_min" (MacaqueV values from which a row is decoded by a whole block), "grid_tma_stores" (the tile kernel that stages tiles in shared memory and stores them with the TMA engine; measured
 * slower than per-thread stores on B200, which stay the default).  Unknown names fail. */
int mdbcu_context_set_option(mdbcu_context *ctx, const char *name, int64_t value);
/* How many points before its chunk a speculative one-lane-per-chain chain starts (engine 4; csrc/mdb_fit_lanes.cuh).
 * Tuning only: results never depend on it. */
int mdbcu_context_set_lane_warmup(mdbcu_context *ctx, uint32_t points);
/* Number of chain rounds the last mdbcu_compress on this context needed (1 = no re-run at all; always
 * 1 with the asynchronous scheduler, which has no rounds). */
uint32_t mdbcu_context_last_compress_rounds(const mdbcu_context *ctx);
/* Which fit_next_model engine runs the chains and how they are scheduled: 0 automatic (5; 3 when no unit has
 * a lossy bound), 1 one
 * thread per chain in global rounds, 2 one warp per chain (32 lanes fit 128 points per step,
 * csrc/mdb_fit_warp.cuh) in global rounds, 3 one warp per chain served from a device-side work queue
 * by persistent warps, each unit advancing its own exact frontier (csrc/mdb_compress.cuh,
 * sched_advance), 4 one LANE per chain for the bulk of the chains (every thread walks its own chunk point by point,
 * csrc/mdb_fit_lanes.cuh; regular units with finite values) followed by 3 for the exact stitching and for everything
 * the lanes leave alone, 5 = 3 with the screened fit (csrc/mdb_fit_screen.cuh: Swing's comparisons decided in f32
 * where they are certain and by the reference's own f64 operations where they are not; units with regular timestamps
 * and a lossy bound, everything else falls to the exact engine inside the same kernel).  Results are identical. */
int mdbcu_context_set_fit_engine(mdbcu_context *ctx, int engine);

/* Diagnostics: fit_next_model (compression.rs:280-301) at each of `starts` with the chosen engine
 * (1 one thread, 2 one warp, 5 one warp with the screened fit); out receives n_starts records of 40 bytes {u32 start, u32 end, f32 min,
 * f32 max, f32 last, f32 bytes_per_value, i32 model_type_id, i32 values_len, i32 aborted, i32 irregular}.
 * Host space only.  Used by the tests to compare the two engines model by model. */
int mdbcu_debug_fit_models(mdbcu_context *ctx, const int64_t *timestamps, const float *values, uint32_t n,
                           int eb_kind, float eb_value, int engine, const uint32_t *starts,
                           const uint32_t *budget_ends, uint32_t n_starts, void *out);

/* Diagnostics: reads and clears eight event counters of the warp fit engine (fits, fits handed to the
 * one-thread code, steps, quiet steps, speculation passes, mismatches, in-order PMC sums, unused).
 * They only count in a library built with -DMDB_FIT_COUNTERS. */
int mdbcu_debug_counters(mdbcu_context *ctx, uint64_t *out8);

/* Diagnostics: `23 - floor(|log2(x)|) as i32` of rewrite_least_mantissa_bits (macaque_v.rs:185) evaluated on the
 * device for every f32 bit pattern in [first_bits, last_bits], reported as a step function: the patterns at which
 * the position changes (first_bits always listed), unordered; *n_steps may exceed cap.  Host space only.  The tests
 * compare it with libm's log2f, which the reference uses, over all non-negative patterns. */
int mdbcu_debug_rewrite_position_steps(mdbcu_context *ctx, uint32_t first_bits, uint32_t last_bits,
                                       uint32_t *bits_out, int32_t *pos_out, uint32_t cap, uint32_t *n_steps);

uint64_t mdbcu_segments_len(const mdbcu_segments *segments);
/* Columns of an owned batch in `space` (a host copy is made on first request).  unit_seg_off
 * (nullable) receives a pointer to n_units + 1 row offsets: unit u produced rows
 * [unit_seg_off[u], unit_seg_off[u+1]).  Pointers stay valid until mdbcu_segments_free. */
int mdbcu_segments_get(mdbcu_segments *segments, mdbcu_space space, mdbcu_segments_view *view,
                       const uint64_t **unit_seg_off);
void mdbcu_segments_free(mdbcu_segments *segments);

/* ---- K2: grid ------------------------------------------------------------------------------ */

/* point_off (n_segments + 1, nullable) receives the exclusive prefix sum of len() per row;
 * *total (host, nullable) the number of data points in the batch. */
int mdbcu_grid_count(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments,
                     uint64_t *point_off, uint64_t *total);
/* Reconstruct every data point of every row, rows in order (grid_exec.rs:323-337 appends the same
 * way).  Fails if the batch holds more than `capacity` points; *n_points (host) receives the count. */
int mdbcu_grid(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments,
               int64_t *timestamps_out, float *values_out, uint64_t capacity, uint64_t *n_points);

/* grid() with the query's time predicate pushed in.  The reference selects segments by start_time / end_time
 * (crates/modelardb_storage/src/query/time_series_table.rs:290-373), reconstructs every data point of them and then
 * prunes by the predicate (grid_exec.rs:366-387).  Here rows that end before t_lo or start after t_hi are never
 * reconstructed, and of the others only the data points with t_lo <= timestamp <= t_hi are written, rows in order:
 * the same points GridStream yields for the predicate `t_lo <= timestamp AND timestamp <= t_hi`.
 * point_off (n_segments + 1, nullable) receives the exclusive prefix sum of the points written per row (0 for pruned
 * rows); *n_points (host) their number.  With both outputs null the call only counts; otherwise it fails if more than
 * `capacity` points lie in the range. */
int mdbcu_grid_range(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments, int64_t t_lo,
                     int64_t t_hi, uint64_t *point_off, int64_t *timestamps_out, float *values_out,
                     uint64_t capacity, uint64_t *n_points);

/* ---- the front end of try_compress_multivariate_time_series --------------------------------- */

/* The row order `sort_time_series_by_tags_and_time` establishes (compression.rs:110-141: lexsort_to_indices over the tag
 * columns, then the timestamp column, ascending): series_code[i] is the dense code of row i's tag tuple, codes ordered
 * like the tuples (the tag strings stay with the host); order_out[k] = index of the row that comes k-th.  A stable LSD
 * radix sort on the device over the bytes of (timestamp - min timestamp) and of the code that actually differ; rows
 * with equal tags and timestamp keep their input order (the reference leaves it unspecified). */
int mdbcu_sort_rows(mdbcu_context *ctx, mdbcu_space space, const uint32_t *series_code, const int64_t *timestamps,
                    uint64_t n, uint32_t *order_out);
/* compute::take_arrays of the same function: out[k] = in[order[k]] for the timestamp column (nullable) and n_fields field
 * columns. */
int mdbcu_take_rows(mdbcu_context *ctx, mdbcu_space space, const uint32_t *order, uint64_t n,
                    const int64_t *timestamps_in, int64_t *timestamps_out, const float *const *fields_in,
                    float *const *fields_out, uint32_t n_fields);

/* ---- K3: aggregates ------------------------------------------------------------------------ */

/* Per-row `sum` exactly as models/mod.rs:129-184 computes it (f32). */
int mdbcu_segment_sums(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments,
                       float *sums_out);
/* Fold rows [group_off[g], group_off[g+1]) into group g without materialising data points:
 * count += len (i64), min/max fold the metadata columns from f32::MAX / f32::MIN with NaN-ignoring
 * min/max, sum += (f64) per-row f32 sum.  group_off == NULL means one group over all rows
 * (what the reference's rule rewrites); GROUP BY series passes unit_seg_off.  AVG = sum / count.
 * COUNT, MIN and MAX are exact.  The reference adds the row sums to `self.sum` one row after the other
 * (model_simple_aggregates.rs:481-511); here they are added in a fixed tree whose shape depends only on the batch (rows per
 * group), never on the device, so SUM is reproducible everywhere and within 1e-12 relative of the row-by-row fold -- far
 * inside the 0.001 % the reference's own tests allow between segment and data point aggregates
 * (modelardb_server/tests/integration_test.rs:1184-1246). */
int mdbcu_aggregate(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments,
                    const uint64_t *group_off, uint64_t n_groups, int64_t *count, float *min,
                    float *max, double *sum);

/* ---- multi-GPU ----------------------------------------------------------------------------- */

/* The path shards by unit (time series): a rank compresses, grids and aggregates the units it owns and nothing of
 * that crosses GPUs (the reference treats every series independently: compression.rs:95-104, grid_exec.rs:323-356).
 * Only the small result of an aggregate query travels, over NCCL / NVLink: what DataFusion's final aggregation merges
 * from the partial states of model_simple_aggregates.rs:367-606.  One communicator per context (= per GPU); NCCL is
 * loaded at run time, so a host that never creates a communicator does not need it. */
typedef struct mdbcu_comm mdbcu_comm;

/* Contiguous, balanced range [*lo, *hi) of the units rank `rank` of `world` owns (the first n_units % world ranks
 * get one more). */
int mdbcu_shard_units(uint64_t n_units, int world, int rank, uint64_t *lo, uint64_t *hi);
/* One process (or thread) per GPU: rank 0 obtains 128 bytes, hands them to the others by any means, and every rank
 * joins with its context. */
int mdbcu_comm_unique_id(uint8_t *id128);
int mdbcu_comm_create(mdbcu_context *ctx, int world, int rank, const uint8_t *id128, mdbcu_comm **out);
/* One process driving n GPUs: communicators for n contexts (one per device) at once; out receives n handles.  Calls
 * on different communicators must come from different host threads (each call blocks until its stream is done). */
int mdbcu_comm_create_all(mdbcu_context *const *ctxs, int n, mdbcu_comm **out);
void mdbcu_comm_destroy(mdbcu_comm *comm);
int mdbcu_comm_world(const mdbcu_comm *comm);
int mdbcu_comm_rank(const mdbcu_comm *comm);
/* mdbcu_aggregate GROUP BY unit over a table whose n_total units are sharded with mdbcu_shard_units: `segments` and
 * group_off (n_local + 1 entries) describe THIS rank's units; count / min / max / sum receive all n_total groups in
 * unit order on every rank.  The ranks' 24-byte records are exchanged by ONE ncclAllGather.  Collective: every rank
 * of the communicator must call it. */
int mdbcu_aggregate_sharded(mdbcu_comm *comm, mdbcu_space space, const mdbcu_segments_view *segments,
                            const uint64_t *group_off, uint64_t n_local, uint64_t n_total, int64_t *count,
                            float *min, float *max, double *sum);
/* The ungrouped aggregate over rows sharded across the ranks: one record per rank, folded in rank order (= row
 * order, so the f64 sum does not depend on a reduction tree); single values, identical on every rank.  Collective. */
int mdbcu_aggregate_all_sharded(mdbcu_comm *comm, mdbcu_space space, const mdbcu_segments_view *segments,
                                int64_t *count, float *min, float *max, double *sum);

#ifdef __cplusplus
}
#endif
#endif /* MODELARDB_CUDA_H */
