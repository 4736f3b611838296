/* mdb_oracle.h -- C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's
 * `modelardb_compression` crate (crates/modelardb_compression/src/...) and of the
 * model accumulators (crates/modelardb_storage/src/optimizer/model_simple_aggregates.rs).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product path (modelardb_rs_b200/) never does.
 *
 * Parity pinning: the reference is Rust and cannot be built in this image (no
 * cargo/rustc), so the oracle is pinned against every golden vector / known-answer
 * test the reference's own unit tests hold for this path (tests/test_oracle_golden.py
 * lists them with file:line).  Unpinned corners (stated, not hidden):
 *   - signed-zero ties in f32::min/max (x86-64 LLVM lowering assumed, see rust_min),
 *   - f32::log2 within 1 ulp of an integer (glibc log2f assumed, see floor_abs_log2).
 */
#ifndef MDB_ORACLE_H
#define MDB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error bound kinds (modelardb_types/src/types.rs:299-335). */
enum { MDBO_LOSSLESS = 0, MDBO_ABSOLUTE = 1, MDBO_RELATIVE = 2 };

/* Model type ids (models/mod.rs:36-38). */
enum { MDBO_PMC_MEAN = 0, MDBO_SWING = 1, MDBO_MACAQUE_V = 2 };

/* ---- scalar pieces (models/mod.rs:53-95) ---- */
int mdbo_is_value_within_error_bound(int kind, float eb, float real_value, float approximate_value);
double mdbo_maximum_allowed_deviation(int kind, float eb, double value);

/* ---- MacaqueTS (models/timestamps.rs) ---- */
/* Returns the number of bytes written (<= cap), or (size_t)-1 if cap is too small. */
size_t mdbo_compress_residual_timestamps(const int64_t *ts, size_t n, uint8_t *out, size_t cap);
size_t mdbo_decompress_all_timestamps(int64_t start_time, int64_t end_time, const uint8_t *bytes,
                                      size_t n_bytes, int64_t *out, size_t cap);
size_t mdbo_len(int64_t start_time, int64_t end_time, const uint8_t *bytes, size_t n_bytes);

/* ---- MacaqueV (models/macaque_v.rs) ---- */
/* has_seed != 0 -> compress_values_without_first(values, seed). state_out (nullable) receives
 * {last_leading_zero_bits, last_trailing_zero_bits}; last_value_out (nullable) the last stored value. */
size_t mdbo_macaque_v_compress(int kind, float eb, const float *values, size_t n, int has_seed,
                               float seed, uint8_t *out, size_t cap, float *min_out, float *max_out,
                               uint8_t *state_out, float *last_value_out);
void mdbo_macaque_v_grid(const uint8_t *bytes, size_t n_bytes, size_t n_values, int has_seed,
                         float seed, float *out);
float mdbo_macaque_v_sum(const uint8_t *bytes, size_t n_bytes, size_t n_values, int has_seed,
                         float seed);
/* rewrite_least_mantissa_bits (macaque_v.rs:168-196), exposed for the log2 boundary tests. */
float mdbo_rewrite_least_mantissa_bits(int kind, float eb, float value);
/* 23 - floor(|log2f(x)|) as i32, the reference way (libm log2f) and the f64 way used on the GPU. */
int32_t mdbo_rewrite_position_libm(float factorized_epsilon);
int32_t mdbo_rewrite_position_f64(float factorized_epsilon);
/* The position as a step function of the bit pattern over [first_bits, last_bits]: the patterns at which it changes
 * (first_bits always listed), ascending; which = 0 libm, 1 f64.  Returns the number of steps (may exceed cap). */
size_t mdbo_rewrite_position_steps(int which, uint32_t first_bits, uint32_t last_bits, int n_threads,
                                   uint32_t *bits_out, int32_t *pos_out, size_t cap);

/* ---- model fitting (types.rs:40-145, pmc_mean.rs, swing.rs) ---- */
typedef struct {
    int8_t model_type_id;
    uint64_t start_index;
    uint64_t end_index;
    float min_value;
    float max_value;
    uint8_t values[8];
    uint32_t values_len;
    float model_last_value;
    float bytes_per_value;
    uint64_t pmc_len;   /* diagnostics: lengths reached by both models */
    uint64_t swing_len;
} mdbo_model;

void mdbo_fit_next_model(uint64_t start_index, int kind, float eb, const int64_t *ts,
                         const float *values, uint64_t n, mdbo_model *out);

/* PMC-Mean / Swing incremental fitters for the fit / no-fit matrices
 * (pmc_mean.rs:119-335, swing.rs:366-798).  Returns how many leading points were accepted. */
uint64_t mdbo_pmc_fit_prefix(int kind, float eb, const float *values, uint64_t n, float *mean_out);
uint64_t mdbo_swing_fit_prefix(int kind, float eb, const int64_t *ts, const float *values,
                               uint64_t n, float *first_out, float *last_out);

/* lower/upper bound (slope, intercept) after fitting a prefix: {ls, li, us, ui} (swing.rs:45-56). */
void mdbo_swing_bounds(int kind, float eb, const int64_t *ts, const float *values, uint64_t n, double *out4);

/* ---- segment batches (types.rs:411-517, schemas.rs:40-52) ---- */
typedef struct mdbo_segments mdbo_segments;

typedef struct {
    uint64_t n_segments;
    const int8_t *model_type_id;
    const int64_t *start_time;
    const int64_t *end_time;
    const float *min_value;
    const float *max_value;
    const uint64_t *timestamps_off; /* n_segments + 1 */
    const uint8_t *timestamps_data;
    const uint64_t *values_off;
    const uint8_t *values_data;
    const uint64_t *residuals_off;
    const uint8_t *residuals_data;
} mdbo_segments_view;

/* try_compress_univariate_time_series (compression.rs:191-275) over n_units independent
 * (timestamps, values) slices [unit_off[u], unit_off[u+1]); unit_seg_off_out (n_units+1, nullable)
 * receives the first segment row of each unit.  n_threads > 1 partitions units over threads. */
mdbo_segments *mdbo_compress(const int64_t *ts, const float *values, const uint64_t *unit_off,
                             uint64_t n_units, const uint8_t *eb_kind, const float *eb_value,
                             int n_threads, uint64_t *unit_seg_off_out);
void mdbo_segments_view_get(const mdbo_segments *s, mdbo_segments_view *out);
void mdbo_segments_free(mdbo_segments *s);

/* CompressedSegmentBuilder::finish (types.rs:197-267): one segment from a fitted model plus
 * residuals up to residuals_end_index. */
mdbo_segments *mdbo_model_finish(const mdbo_model *model, int kind, float eb,
                                 uint64_t residuals_end_index, const int64_t *ts,
                                 const float *values);
/* compress_and_store_residuals_in_a_separate_segment (compression.rs:367-400). */
mdbo_segments *mdbo_macaque_v_segment(int kind, float eb, uint64_t start_index, uint64_t end_index,
                                      const int64_t *ts, const float *values);

/* decode_values_for_pmc_mean / decode_values_for_swing (types.rs:307-321, 374-407). */
float mdbo_decode_values_for_pmc_mean(float min_value, float max_value, const uint8_t *values,
                                      size_t n);
int mdbo_decode_values_for_swing(float min_value, float max_value, const uint8_t *values, size_t n,
                                 float *first_out, float *last_out);

/* ---- grid / sum / len over a batch (models/mod.rs:98-251) ---- */
/* point_off_out: n_segments+1 (nullable). Returns total points, or (uint64_t)-1 on a malformed row. */
uint64_t mdbo_grid_count(const mdbo_segments_view *v, uint64_t *point_off_out, int n_threads);
uint64_t mdbo_grid(const mdbo_segments_view *v, int64_t *ts_out, float *val_out, uint64_t capacity,
                   int n_threads);
/* per-segment `sum` (f32) -- models/mod.rs:129-184 */
void mdbo_segment_sums(const mdbo_segments_view *v, float *sum_out, int n_threads);

/* Model accumulators (model_simple_aggregates.rs:336-618), folded row by row in order within each
 * group [group_off[g], group_off[g+1]); group_off == NULL means one group over all rows. */
void mdbo_aggregate(const mdbo_segments_view *v, const uint64_t *group_off, uint64_t n_groups,
                    int64_t *count, float *min, float *max, double *sum, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
