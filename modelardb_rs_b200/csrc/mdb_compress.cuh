// mdb_compress.cuh -- per-thread bodies of the compress kernels, K1.
//
// try_compress_univariate_time_series (compression.rs:191-275) is a greedy, strictly ordered
// segmenter: segment k+1 starts where segment k ended, and a rejected start index is re-fitted from
// the next index.  One unit (one call of the reference function) is therefore one chain; units are
// independent.  The work is split in two so that the byte columns can be allocated exactly:
//   pass 1  spec_chain ...:      runs the chain (in parallel over chunks, see below) and turns the
//                                accepted models into one fixed-size SegRecord per segment row (model,
//                                boundaries, metadata, and the byte LENGTH of each of the three binary
//                                columns, obtained by running the encoders on a counter);
//   pass 2  compress_emit_segment: one thread per segment row re-runs the encoders on a writer at the
//                                offsets given by an exclusive scan of those lengths.
#pragma once

#include "mdb_device.cuh"

namespace mdb {

constexpr uint32_t RESIDUAL_VALUES_MAX_LENGTH = 255; // compression.rs:38

// Upper bound on the number of segment rows of a unit of n points: a stored model covers >= 8 points
// (29/len <= 4, compression.rs:238), a separate MacaqueV row follows a model only with > 255 residuals
// (compression.rs:320), plus one leading MacaqueV row.
MDB_DEV uint64_t max_segments_of_unit(uint64_t n) { return n == 0 ? 0 : n / 8 + n / 264 + 2; }

struct SegRecord {           // 48 bytes
    uint32_t start_index;    // first point of the row, relative to the unit
    uint32_t model_end_index;// last point represented by the model (== res_end_index for MacaqueV rows)
    uint32_t res_end_index;  // last point of the row (model + residuals)
    uint32_t ts_len;         // byte length of the `timestamps` column
    uint32_t val_len;        // byte length of the `values` column
    uint32_t res_len;        // byte length of the `residuals` column
    float min_value, max_value;
    float model_last_value;  // seed of the residual encoder (types.rs:270-278)
    uint8_t values[8];       // `values` of PMC-Mean / Swing rows (types.rs:283-370)
    int8_t model_type_id;
    uint8_t regular;         // timestamps of the row are regular
    uint8_t wide;            // long MacaqueV row: val_len / min / max and the value bytes come from the warp kernels
    uint8_t pad;
};

// MacaqueV rows of at least this many values are encoded by a whole warp (k_records_macaque_warp,
// k_emit_macaque_warp) when the caller passes it as defer_min; 0 never defers (host emulation, tests).
constexpr uint32_t WIDE_ENCODE_MIN = 256;
static_assert(sizeof(SegRecord) == 48, "SegRecord layout");

// ------------------------------------------------------------------------------------------------
// model fitting
// ------------------------------------------------------------------------------------------------

struct PMCMean { // pmc_mean.rs:31-93
    float min_value, max_value;
    double sum_of_values;
    uint32_t length;
    MDB_DEV void init() {
        min_value = max_value = __uint_as_float(0x7fc00000u);
        sum_of_values = 0.0;
        length = 0;
    }
    MDB_DEV bool fit_value(const ErrorBound &eb, float value) { // pmc_mean.rs:58-75
        float next_min = rust_minf(min_value, value);
        float next_max = rust_maxf(max_value, value);
        double next_sum = __dadd_rn(sum_of_values, (double)value);
        uint32_t next_length = length + 1;
        float average = __double2float_rn(__ddiv_rn(next_sum, (double)next_length));
        if (is_value_within_error_bound(eb, next_min, average) && is_value_within_error_bound(eb, next_max, average)) {
            min_value = next_min; max_value = next_max; sum_of_values = next_sum; length = next_length;
            return true;
        }
        return false;
    }
    MDB_DEV float model() const { return canonical_nan(__double2float_rn(__ddiv_rn(sum_of_values, (double)length))); }
};

struct Swing { // swing.rs:34-259
    int64_t start_time, end_time;
    double first_value;
    double upper_slope, upper_intercept, lower_slope, lower_intercept;
    double mse_numerator, mse_denominator;
    uint32_t length;
    MDB_DEV void init() {
        start_time = end_time = 0;
        first_value = upper_slope = upper_intercept = lower_slope = lower_intercept = (double)__uint_as_float(0x7fc00000u);
        mse_numerator = mse_denominator = 0.0;
        length = 0;
    }

    MDB_DEV bool fit_data_point(const ErrorBound &eb, int64_t timestamp, float value_f32) { // swing.rs:101-198
        double value = (double)value_f32;
        double maximum_deviation = maximum_allowed_deviation(eb, value);
        if (length == 0) {
            start_time = timestamp; end_time = timestamp; first_value = value; length = 1;
            return true;
        }
        bool first_finite = !(first_value != first_value) && !isinf(first_value);
        bool value_finite = !(value != value) && !isinf(value);
        if (!first_finite || !value_finite) {
            if (equal_or_nan(first_value, value)) {
                end_time = timestamp;
                upper_slope = upper_intercept = lower_slope = lower_intercept = value;
                length += 1;
                return true;
            }
            return false;
        }
        if (length == 1) {
            end_time = timestamp;
            compute_slope_and_intercept(start_time, first_value, timestamp, __dadd_rn(value, maximum_deviation),
                                        upper_slope, upper_intercept);
            if (maximum_deviation == 0.0) { // value + 0 and value - 0 are the same operand: the same line, one division
                lower_slope = upper_slope;
                lower_intercept = upper_intercept;
            } else {
                compute_slope_and_intercept(start_time, first_value, timestamp, __dsub_rn(value, maximum_deviation),
                                            lower_slope, lower_intercept);
            }
            length = 2;
            return true;
        }
        double t = (double)timestamp;
        double upper = __dadd_rn(__dmul_rn(upper_slope, t), upper_intercept);
        double lower = __dadd_rn(__dmul_rn(lower_slope, t), lower_intercept);
        if (__dadd_rn(upper, maximum_deviation) < value || __dsub_rn(lower, maximum_deviation) > value) return false;
        end_time = timestamp;
        if (__dsub_rn(upper, maximum_deviation) > value)
            compute_slope_and_intercept(start_time, first_value, timestamp, __dadd_rn(value, maximum_deviation),
                                        upper_slope, upper_intercept);
        if (__dadd_rn(lower, maximum_deviation) < value)
            compute_slope_and_intercept(start_time, first_value, timestamp, __dsub_rn(value, maximum_deviation),
                                        lower_slope, lower_intercept);
        // swing.rs:212-228 (0.0 is added when value == first_value)
        double num = 0.0, den = 0.0;
        if (!equal_or_nan(first_value, value)) {
            double dt = (double)(timestamp - start_time);
            num = __dmul_rn(__dsub_rn(value, first_value), dt);
            den = __dmul_rn(dt, dt);
        }
        mse_numerator = __dadd_rn(mse_numerator, num);
        mse_denominator = __dadd_rn(mse_denominator, den);
        length += 1;
        return true;
    }
    MDB_DEV void model(float &first_out, float &last_out) const { // swing.rs:246-259
        double projected = __ddiv_rn(mse_numerator, mse_denominator);
        double slope = rust_maxd(lower_slope, rust_mind(projected, upper_slope));
        double last = __dadd_rn(__dmul_rn(slope, (double)(end_time - start_time)), first_value);
        first_out = canonical_nan(__double2float_rn(first_value));
        last_out = canonical_nan(__double2float_rn(last));
    }
};

struct FittedModel { // CompressedSegmentBuilder, types.rs:148-166
    uint32_t start_index, end_index;
    float min_value, max_value, model_last_value, bytes_per_value;
    int8_t model_type_id;
    uint8_t values_len; // Swing: 0 -> [] (first < last), 1 -> [0]
    // A Swing model fitted by the warp engine is left `pending`: its boundaries (all the chain needs) are
    // final, but min/max/last still need the two MSE sums of swing.rs:212-228, which are order dependent.
    // They are accumulated afterwards, strictly in order, by one thread per ACCEPTED model
    // (swing_finish), instead of on the latency-critical path of the chain.
    uint8_t pending;
    uint8_t pad;
    double lower_slope, upper_slope; // Swing bounds when the fit ended (valid when pending)
};

// Completes a pending Swing model: the MSE sums over its points in order (swing.rs:180-193, 212-228),
// then Swing::model (swing.rs:246-259) and select_swing (types.rs:122-144).
// The MSE terms of one point (swing.rs:212-228): (0, 0) when the value equals the first value.
MDB_DEV void swing_mse_terms(int64_t t0, double v0, int64_t t, double v, double &x, double &y) {
    x = 0.0;
    y = 0.0;
    if (!equal_or_nan(v0, v)) {
        const double dt = (double)(t - t0);
        x = __dmul_rn(__dsub_rn(v, v0), dt);
        y = __dmul_rn(dt, dt);
    }
}

MDB_DEV void swing_finish_from_sums(FittedModel &m, double num, double den, const int64_t *ts, const float *values);

MDB_DEV void swing_finish(FittedModel &m, const int64_t *ts, const float *values) {
    const int64_t t0 = ts[m.start_index];
    const double v0 = (double)values[m.start_index];
    double num = 0.0, den = 0.0;
    for (uint32_t i = m.start_index + 2; i <= m.end_index; i++) { // the first two points add no term
        double x, y;
        swing_mse_terms(t0, v0, ts[i], (double)values[i], x, y);
        num = __dadd_rn(num, x);
        den = __dadd_rn(den, y);
    }
    swing_finish_from_sums(m, num, den, ts, values);
}

MDB_DEV void swing_finish_from_sums(FittedModel &m, double num, double den, const int64_t *ts, const float *values) {
    const int64_t t0 = ts[m.start_index];
    const double v0 = (double)values[m.start_index];
    const double projected = __ddiv_rn(num, den);
    const double slope = rust_maxd(m.lower_slope, rust_mind(projected, m.upper_slope));
    const double last_d = __dadd_rn(__dmul_rn(slope, (double)(ts[m.end_index] - t0)), v0);
    const float first = canonical_nan(__double2float_rn(v0));
    const float last = canonical_nan(__double2float_rn(last_d));
    m.min_value = rust_minf(first, last);
    m.max_value = rust_maxf(first, last);
    m.values_len = (first < last) ? 0 : 1;
    m.model_last_value = last;
    m.pending = 0;
}

// Tracks, as the chain first touches each point of the unit, whether any sampling-interval change has
// been seen so far.  While none has, every row is regular (a sub-range of a regular range is regular)
// and timestamps_encoded_len needs no scan of its own; this avoids a second pass over the timestamps.
struct RegularityTracker {
    uint32_t max_seen;   // highest index visited so far
    int64_t ts_max_seen; // its timestamp
    int64_t delta0;      // the unit's first sampling interval, ts[1] - ts[0]
    bool irregular;
    // A chain that starts at index `first` of a unit of n points.
    MDB_DEV void init(const int64_t *ts, uint32_t first, uint32_t n) {
        max_seen = first;
        ts_max_seen = ts[first];
        delta0 = n >= 2 ? ts[1] - ts[0] : 0;
        irregular = first > 0 && (ts_max_seen - ts[first - 1]) != delta0;
    }
    MDB_DEV void visit(uint32_t i, int64_t t) { // visits are contiguous: i <= max_seen + 1
        if (i <= max_seen) return;
        if (t - ts_max_seen != delta0) irregular = true;
        max_seen = i;
        ts_max_seen = t;
    }
};

// fit_next_model (compression.rs:280-301) + ModelBuilder (types.rs:61-144).  A fit that is still
// growing when it reaches `budget_end` (< n) is abandoned (aborted = true): only speculative chains are
// given a budget, see spec_chain.
MDB_DEV FittedModel fit_next_model(const ErrorBound &eb, const int64_t *ts, const float *values, uint32_t start,
                                   uint32_t n, RegularityTracker &trk, uint32_t budget_end, bool &aborted) {
    PMCMean pmc; pmc.init();
    Swing swing; swing.init();
    bool pmc_ok = true, swing_ok = true;
    uint32_t i = start;
    aborted = false;
    while ((pmc_ok || swing_ok) && i < n) {
        if (i >= budget_end) { aborted = true; break; }
        float value = values[i];
        int64_t t = ts[i];
        trk.visit(i, t);
        if (pmc_ok) pmc_ok = pmc.fit_value(eb, value);
        if (swing_ok) swing_ok = swing.fit_data_point(eb, t, value);
        i++;
    }
    FittedModel m;
    m.start_index = start;
    m.pending = 0;
    m.pad = 0;
    m.lower_slope = m.upper_slope = 0.0;
    if (aborted) { // the caller discards an abandoned fit
        m.end_index = start;
        m.min_value = m.max_value = m.model_last_value = 0.0f;
        m.bytes_per_value = 1e30f;
        m.model_type_id = PMC_MEAN;
        m.values_len = 0;
        return m;
    }
    float pmc_bpv = __fdiv_rn(29.0f, (float)pmc.length);    // pmc_mean.rs:83-87
    float swing_bpv = __fdiv_rn(30.0f, (float)swing.length); // swing.rs:236-239
    if (swing_bpv < pmc_bpv) { // min_by keeps the first minimum: PMC-Mean wins ties (types.rs:90-94)
        float first, last;
        swing.model(first, last);
        m.model_type_id = SWING;
        m.end_index = start + swing.length - 1;
        m.min_value = rust_minf(first, last);
        m.max_value = rust_maxf(first, last);
        m.values_len = (first < last) ? 0 : 1;
        m.model_last_value = last;
        m.bytes_per_value = swing_bpv;
    } else {
        float value = pmc.model();
        m.model_type_id = PMC_MEAN;
        m.end_index = start + pmc.length - 1;
        m.min_value = m.max_value = m.model_last_value = value;
        m.values_len = 0;
        m.bytes_per_value = pmc_bpv;
    }
    return m;
}

// ------------------------------------------------------------------------------------------------
// encoders shared by the sizing pass (Sink = BitCounter) and the emit pass (Sink = BitWriter)
// ------------------------------------------------------------------------------------------------

// MacaqueV over values[lo..=hi]; seeded -> compress_values_without_first (macaque_v.rs:92-97),
// else compress_values with the first value raw (macaque_v.rs:76-88).
template <typename Sink>
MDB_DEV void macaque_v_encode(const ErrorBound &eb, const float *values, uint32_t lo, uint32_t hi, bool seeded,
                              float seed, Sink &sink, float &min_out, float &max_out) {
    MacaqueVEncoder enc;
    enc.init();
    uint32_t i = lo;
    if (seeded) enc.last_value = seed;
    else enc.first_raw(values[i++], sink);
    for (; i <= hi; i++) enc.compress_value_xor_last_value(eb, values[i], sink);
    min_out = enc.min_value;
    max_out = enc.max_value;
}

MDB_DEV void put_le(uint8_t *p, float f) {
    uint32_t b = __float_as_uint(f);
    p[0] = (uint8_t)b; p[1] = (uint8_t)(b >> 8); p[2] = (uint8_t)(b >> 16); p[3] = (uint8_t)(b >> 24);
}

// types.rs:283-303; returns the length written to out[0..8)
MDB_DEV uint32_t encode_values_for_pmc_mean(float min_value, float max_value, float res_min, float res_max, uint8_t *out) {
    if (min_value > res_min) {
        if (max_value >= res_max) { out[0] = 1; return 1; }
        put_le(out, min_value);
        return 4;
    }
    return 0;
}

// types.rs:325-370
MDB_DEV uint32_t encode_values_for_swing(float min_value, float max_value, bool min_value_is_first, float res_min,
                                         float res_max, uint8_t *out) {
    if (res_min < min_value && max_value < res_max) {
        if (min_value_is_first) { put_le(out, min_value); put_le(out + 4, max_value); }
        else { put_le(out, max_value); put_le(out + 4, min_value); }
        return 8;
    } else if (res_min < min_value) {
        out[0] = min_value_is_first ? 0 : 1;
        put_le(out + 1, min_value);
        return 5;
    } else if (max_value < res_max) {
        out[0] = min_value_is_first ? 2 : 3;
        put_le(out + 1, max_value);
        return 5;
    } else if (!min_value_is_first) {
        out[0] = 0;
        return 1;
    }
    return 0;
}

// Byte length of compress_residual_timestamps(ts[lo..=hi]) (timestamps.rs:56-73) and whether regular.
MDB_DEV uint32_t timestamps_encoded_len(const int64_t *ts, uint32_t lo, uint32_t hi, bool known_regular, uint8_t &regular) {
    uint64_t n = (uint64_t)hi - lo + 1;
    regular = 1;
    if (n <= 2) return 0;
    if (known_regular || are_uncompressed_timestamps_regular(ts + lo, n)) return regular_timestamps_bytes(n);
    regular = 0;
    BitCounter c;
    compress_irregular_residual_timestamps(ts + lo, n, c);
    return (uint32_t)c.bytes();
}

// ------------------------------------------------------------------------------------------------
// pass 1: the chain
// ------------------------------------------------------------------------------------------------

// CompressedSegmentBuilder::finish (types.rs:197-267) as a record.
MDB_DEV void record_model_segment(const ErrorBound &eb, const FittedModel &m, uint32_t res_end, const int64_t *ts,
                                  const float *values, bool unit_regular, SegRecord &rec) {
    rec.start_index = m.start_index;
    rec.model_end_index = m.end_index;
    rec.res_end_index = res_end;
    rec.model_type_id = m.model_type_id;
    rec.model_last_value = m.model_last_value;
    rec.min_value = m.min_value;
    rec.max_value = m.max_value;
    rec.wide = 0;
    rec.pad = 0;
    for (int k = 0; k < 8; k++) rec.values[k] = 0;
    rec.val_len = m.values_len; // Swing [0] or []
    rec.res_len = 0;
    rec.ts_len = timestamps_encoded_len(ts, m.start_index, res_end, unit_regular, rec.regular);
    if (m.end_index < res_end) {
        BitCounter c;
        float res_min, res_max;
        macaque_v_encode(eb, values, m.end_index + 1, res_end, true, m.model_last_value, c, res_min, res_max);
        if (m.model_type_id == PMC_MEAN)
            rec.val_len = encode_values_for_pmc_mean(m.min_value, m.max_value, res_min, res_max, rec.values);
        else
            rec.val_len = encode_values_for_swing(m.min_value, m.max_value, m.values_len == 0, res_min, res_max, rec.values);
        rec.min_value = rust_minf(m.min_value, res_min);
        rec.max_value = rust_maxf(m.max_value, res_max);
        rec.res_len = (uint32_t)c.bytes() + 1; // + the count byte (types.rs:250)
    }
}

// compress_and_store_residuals_in_a_separate_segment (compression.rs:367-400) as a record.
MDB_DEV void record_macaque_v_segment(const ErrorBound &eb, uint32_t lo, uint32_t hi, const int64_t *ts,
                                      const float *values, bool unit_regular, SegRecord &rec, uint32_t defer_min = 0) {
    rec.start_index = lo;
    rec.model_end_index = hi;
    rec.res_end_index = hi;
    rec.model_type_id = MACAQUE_V;
    rec.model_last_value = 0.0f;
    rec.wide = 0;
    rec.pad = 0;
    for (int k = 0; k < 8; k++) rec.values[k] = 0;
    rec.res_len = 0;
    rec.ts_len = timestamps_encoded_len(ts, lo, hi, unit_regular, rec.regular);
    if (defer_min && hi - lo + 1 >= defer_min) { // sized (and later written) by a whole warp
        rec.wide = 1;
        rec.val_len = 0;
        rec.min_value = rec.max_value = 0.0f;
        return;
    }
    BitCounter c;
    macaque_v_encode(eb, values, lo, hi, false, 0.0f, c, rec.min_value, rec.max_value);
    rec.val_len = (uint32_t)c.bytes();
}

// store_compressed_segments_with_model_and_or_residuals (compression.rs:310-362). Returns rows written.
MDB_DEV uint32_t store_segments(const ErrorBound &eb, bool have_model, const FittedModel &m, uint32_t res_end,
                                const int64_t *ts, const float *values, bool unit_regular, SegRecord *recs, uint32_t defer_min = 0) {
    if (have_model) {
        if (res_end - m.end_index <= RESIDUAL_VALUES_MAX_LENGTH) {
            record_model_segment(eb, m, res_end, ts, values, unit_regular, recs[0]);
            return 1;
        }
        record_model_segment(eb, m, m.end_index, ts, values, unit_regular, recs[0]);
        record_macaque_v_segment(eb, m.end_index + 1, res_end, ts, values, unit_regular, recs[1], defer_min);
        return 2;
    }
    record_macaque_v_segment(eb, 0, res_end, ts, values, unit_regular, recs[0], defer_min);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// The chain of try_compress_univariate_time_series (compression.rs:224-263), made parallel.
//
// The reference's loop is a strictly ordered chain over "fit starts": cur -> fit_next_model(cur) ->
// (accepted ? model.end + 1 : cur + 1).  fit_next_model(cur) is a pure function of cur, so the chain
// is determined by the set of indices it visits.  A unit is cut into chunks of `chunk_len` points and
// every chunk runs its own chain, at first SPECULATIVELY from the chunk's first index.  Two chains
// that ever visit the same index are identical from there on, so:
//   round 0   every chunk runs a chain from its first index to the chunk end (spec_chain);
//   propagate the true entry of chunk 0 is index 0; the entry of every later chunk is the exit of the
//             chunk before it.  A chunk whose chain was not computed from its entry is marked dirty;
//   round r   a dirty chunk re-runs from its new entry only UNTIL it visits an index its old chain also
//             visited (not strictly inside an old model), then splices the old chain's tail on.
// Rounds repeat until no chunk is dirty.  Chunk 0 always runs from the true entry, and a chunk is
// clean only if its chain started exactly where the previous chunk's chain left, so by induction the
// concatenated chains are exactly the sequential chain: the result is bit-identical by construction,
// never "approximately the same segmentation".  On smooth data chains re-synchronise within a few
// segments, so round 1 touches a small fraction of the points and round 2 is normally empty.
// ------------------------------------------------------------------------------------------------

constexpr uint32_t IDX_NONE = 0xFFFFFFFFu;

struct ChunkState {            // 48 bytes
    uint32_t entry;            // index (in the unit) this chunk's current chain started from; IDX_NONE: none yet
    uint32_t exit;             // first fit start >= chunk end reached by the chain; IDX_NONE: chain was cut short
    uint32_t truncated_at;     // fit start at which a budgeted chain was cut short (valid when exit == IDX_NONE)
    uint32_t n_models;         // accepted models of the chain, in list buffer `buf`
    uint32_t new_entry;        // entry the next round must use (valid when dirty)
    uint32_t next_start;       // finalize: start of the next model after this chunk's last one, or n
    uint32_t lead_end;         // finalize: last index of a leading MacaqueV row owned by this chunk, or IDX_NONE
    uint32_t rows;             // finalize: segment rows this chunk emits
    uint8_t dirty, exact, buf, skipped;
    uint8_t irregular, pad[3];
    uint32_t first_start;      // start of the chain's first model (valid when n_models > 0)
    uint32_t phase;            // scheduler: PH_QUEUED / PH_RUNNING / PH_DONE / PH_LOCKED (asynchronous scheduling only)
};
static_assert(sizeof(ChunkState) == 48, "ChunkState layout");

MDB_DEV uint32_t models_per_chunk(uint32_t chunk_len) { return chunk_len / 8 + 2; }

// One chain over one chunk.  lists: two buffers of models_per_chunk(chunk_len) models each.
// `budget`: a chain that does not start from a known-exact entry abandons a fit that runs more than
// `budget` points past the chunk end (otherwise constant data would make every chunk fit to the end
// of the unit); the cut is resumed later from an exact entry.
// Fit: the fit_next_model engine -- ScalarFit (one thread) or WarpFit (32 lanes cooperate on every fit,
// mdb_fit_warp.cuh); in the warp case every lane runs this control code redundantly on uniform values
// and stores to global memory are made by lane 0 (list copies are spread over the n_lanes lanes).
// Can fit_next_model(start) return a model that gets stored (bytes_per_value <= 4, compression.rs:238)?  Below 8
// points both models cost more than that (29 / len and 30 / len), and each model is fed the points in order until
// it rejects one (types.rs:88-118), so it is enough to feed each of them the first 8 points.  The points
// [start, limit) exist; with fewer than 8 of them no model can be stored.
MDB_DEV bool fit_reaches_eight_points(const ErrorBound &eb, const int64_t *ts, const float *values, uint32_t start, uint32_t limit) {
    if (limit < start || limit - start < 8) return false;
    bool all = true;
    if (eb.kind == KIND_LOSSLESS) {
        // Lossless PMC-Mean on finite values accepts a point iff it equals the first one: min, max and the average of k
        // equal f32 values (k * v is exact in f64 for k <= 8) are that value, and two different finite values can not both
        // equal their average.  No division needed; non-finite values take the general path below.
        const float first = values[start];
        bool finite = fabsf(first) <= 3.402823466e+38f;
        for (uint32_t k = 1; k < 8; k++) {
            const float v = values[start + k];
            finite = finite && fabsf(v) <= 3.402823466e+38f;
            all = all && v == first;
        }
        if (finite && all) return true;
        if (!finite) {
            PMCMean pmc;
            pmc.init();
            all = true;
            for (uint32_t k = 0; k < 8 && all; k++) all = pmc.fit_value(eb, values[start + k]);
            if (all) return true;
        }
    } else {
        PMCMean pmc;
        pmc.init();
        for (uint32_t k = 0; k < 8 && all; k++) all = pmc.fit_value(eb, values[start + k]);
        if (all) return true;
    }
    Swing swing;
    swing.init();
    all = true;
    for (uint32_t k = 0; k < 8 && all; k++) all = swing.fit_data_point(eb, ts[start + k], values[start + k]);
    return all;
}

struct ScalarFit {
    const ErrorBound &eb;
    const int64_t *ts;
    const float *values;
    uint32_t n;
    RegularityTracker trk;
    MDB_DEV ScalarFit(const ErrorBound &e, const int64_t *t, const float *v, uint32_t n_) : eb(e), ts(t), values(v), n(n_) {}
    MDB_DEV void begin(uint32_t cur) { trk.init(ts, cur, n); }
    MDB_DEV FittedModel fit(uint32_t cur, uint32_t budget_end, bool &aborted) {
        return fit_next_model(eb, ts, values, cur, n, trk, budget_end, aborted);
    }
    MDB_DEV bool irregular() const { return trk.irregular; }
    // After a rejected fit: the next index worth fitting (the one-thread engine simply tries the next one).
    MDB_DEV uint32_t skip_rejected(uint32_t from, uint32_t, uint32_t) { return from; }
};

template <typename Fit>
MDB_DEV void spec_chain(Fit &fitter, uint32_t lane, uint32_t n_lanes, uint32_t n, uint32_t chunk_end, uint32_t budget, ChunkState &st,
                        FittedModel *lists, uint32_t cap) {
    const FittedModel *old_list = lists + (size_t)st.buf * cap;
    FittedModel *new_list = lists + (size_t)(st.buf ^ 1) * cap;
    const uint32_t old_n = st.n_models, old_entry = st.entry, old_exit = st.exit, old_trunc = st.truncated_at;
    const bool resume = old_entry != IDX_NONE && st.new_entry == old_entry && old_exit == IDX_NONE;
    const bool can_sync = !resume && old_entry != IDX_NONE;
    const uint32_t sync_limit = old_exit == IDX_NONE ? old_trunc : chunk_end; // old chain visited [old_entry, sync_limit)
    const uint32_t budget_end = st.exact ? n : (uint32_t)((uint64_t)chunk_end + budget < n ? chunk_end + budget : n);

    uint32_t n_new = 0, cur, p = 0;
    uint32_t first_start = IDX_NONE;
    if (resume) {
        for (uint32_t k = lane; k < old_n; k += n_lanes) new_list[k] = load_shared_record(old_list + k); // copies are spread over the lanes
        n_new = old_n;
        if (old_n) first_start = sync_load(&old_list[0].start_index);
        cur = old_trunc;
    } else {
        cur = st.new_entry;
    }
    fitter.begin(cur);
    uint32_t exit = IDX_NONE, truncated_at = 0;
    bool done = false;
    while (cur < chunk_end) {
        if (can_sync && cur >= old_entry && cur < sync_limit) {
            while (p < old_n && sync_load(&old_list[p].end_index) < cur) p++;
            bool inside = p < old_n && sync_load(&old_list[p].start_index) < cur; // strictly inside old model p
            if (!inside) { // the old chain also started a fit at cur: identical from here on
                if (n_new == 0 && p < old_n) first_start = sync_load(&old_list[p].start_index);
                for (uint32_t k = p + lane; k < old_n; k += n_lanes) new_list[n_new + (k - p)] = load_shared_record(old_list + k);
                n_new += old_n - p;
                exit = old_exit;
                truncated_at = old_trunc;
                done = true;
                break;
            }
        }
        bool aborted;
        FittedModel model = fitter.fit(cur, budget_end, aborted);
        if (aborted) {
            truncated_at = cur;
            done = true;
            break;
        }
        if (model.bytes_per_value <= 4.0f) { // compression.rs:238
            if (n_new == 0) first_start = model.start_index;
            if (lane == 0) new_list[n_new] = model;
            n_new++;
            cur = model.end_index + 1;
        } else {
            // compression.rs:261: this point becomes a residual; refit from the next one -- or, with the warp engine,
            // from the next index at which a fit can yield a stored model at all (every index in between is rejected
            // just the same, one residual point each; see WarpFitT::skip_rejected)
            cur = fitter.skip_rejected(cur + 1, chunk_end, budget_end);
        }
    }
    if (!done) exit = cur;
    st.entry = st.new_entry;
    st.exit = exit;
    st.truncated_at = truncated_at;
    st.n_models = n_new;
    st.first_start = first_start;
    st.buf ^= 1;
    st.dirty = 0;
    if (fitter.irregular()) st.irregular = 1;
}

// Walks the chunks of one unit in order and marks the chunks whose chain must be (re)run.
// Returns the number of chunks marked dirty (0: the unit's chains are final).
// allow_optimistic: also re-run chunks AFTER the first inconsistent one, betting that the re-run of the
// chunk before them splices back into its old chain and keeps its exit.  That bet is made once (after
// round 0); on data whose chains do not re-synchronise (very smooth signal, loose bound, segments of
// thousands of points) later rounds re-run only the first inconsistent chunk, whose entry is exact, so
// the total work stays within ~3x the sequential chain instead of growing quadratically.
// resume_c / resume_e: the walk's position (chunk, entry) up to which the unit's chains are already known
// to be final; kept between rounds so that a round does not re-walk the finished prefix of the unit.
// on_dirty(c) is called for every chunk marked dirty (the kernels use it to build the next round's worklist).
template <typename OnDirty>
MDB_DEV uint32_t spec_propagate_unit(uint32_t n, uint32_t chunk_len, uint32_t n_chunks, ChunkState *st, bool allow_optimistic,
                                     uint32_t &resume_c, uint32_t &resume_e, OnDirty &&on_dirty) {
    uint32_t e = resume_e, dirty = 0;
    bool exact = true; // everything before the first inconsistent chunk is the true sequential chain
    uint32_t c = resume_c;
    for (; c < n_chunks; c++) {
        uint32_t chunk_end = (uint64_t)(c + 1) * chunk_len < n ? (c + 1) * chunk_len : n;
        if (e >= chunk_end) continue; // no fit starts in this chunk: a model spans it
        ChunkState &s = st[c];
        if (exact) { resume_c = c; resume_e = e; } // chunk c is the first one not yet known to be final
        if (s.entry == e) {
            if (s.exit == IDX_NONE) { // right entry, but the chain was cut short: resume it
                if (!exact) break;    // (once everything before it is final, so that it runs without a budget)
                s.new_entry = e;
                s.exact = 1;
                s.dirty = 1;
                on_dirty(c);
                dirty++;
                break; // its exit is unknown, nothing after it can be checked yet
            }
            e = s.exit;
        } else {
            if (!exact && !allow_optimistic) break;
            s.new_entry = e;
            s.exact = exact ? 1 : 0;
            s.dirty = 1;
            on_dirty(c);
            dirty++;
            exact = false;
            if (s.exit == IDX_NONE) break;
            e = s.exit; // optimistic: the re-run will most likely splice into the old chain and keep its exit
        }
    }
    if (exact && c >= n_chunks) { resume_c = n_chunks; resume_e = e; } // the whole unit is final
    return dirty;
}

// ------------------------------------------------------------------------------------------------
// Asynchronous scheduling: the same chains without the global rounds.
//
// Rounds make every unit wait for the slowest chain of the round, and a unit whose chains do not
// re-synchronise (its segmentation depends on where the chain started) needs one round per chunk, each
// as long as a whole chunk's chain, while the rest of the GPU idles.  Here every chunk is a work item in a
// device-side queue served by persistent warps, and each unit keeps its own FRONTIER: the first chunk
// whose chain is not yet known to be final, and the exact entry of that chunk.  Whenever a chain of the unit
// completes, sched_advance moves the frontier over every finished chunk whose chain started from the
// right entry, and at the first chunk that did not it either
//   * re-aims the chunk at the exact entry if no worker has started it yet (so a unit whose turn comes late
//     never runs a speculative chain at all: with at least as many units as warps the whole scheme
//     degenerates into one exact, sequential chain per unit, with no wasted work), or
//   * queues a re-run from the exact entry (which splices into the old chain as in the round scheme), or
//   * returns, if that chunk is running right now: its completion calls sched_advance again.
// A unit's sequential dependency is thereby followed as fast as its own chains complete, concurrently with
// everything else.  Exactness is unchanged: a chunk is final only if its chain started at the exit of the
// final chain before it, and chunk 0 starts at index 0.
// ------------------------------------------------------------------------------------------------

constexpr uint32_t PH_QUEUED = 0, PH_RUNNING = 1, PH_DONE = 2, PH_LOCKED = 3;

struct UnitSched {        // 16 bytes, one per unit
    uint32_t lock;        // sched_advance is serialised per unit
    uint32_t next_c;      // frontier: first chunk not yet known to be final
    uint32_t entry;       // exact entry of that chunk
    uint32_t finished;    // the whole unit is final
};

// Called by ONE thread after a chain of the unit has been published (state stored, fence, phase = PH_DONE).
// st: the unit's chunks.  push(c) appends chunk c of this unit to the work queue (after a fence).
// Returns true if this call made the unit final.
template <typename Push>
MDB_DEV bool sched_advance(UnitSched &us, uint32_t n, uint32_t chunk_len, uint32_t n_chunks, ChunkState *st, Push &&push) {
    while (sync_cas(&us.lock, 0u, 1u) != 0u) sync_pause();
    sync_fence();
    bool became_final = false;
    if (!sync_load(&us.finished)) {
        uint32_t c = sync_load(&us.next_c), e = sync_load(&us.entry);
        while (true) {
            if (c >= n_chunks) {
                sync_store(&us.finished, 1u);
                became_final = true;
                break;
            }
            const uint32_t chunk_end = (uint64_t)(c + 1) * chunk_len < n ? (c + 1) * chunk_len : n;
            if (e >= chunk_end) { // no fit starts in this chunk: a model spans it
                c++;
                continue;
            }
            ChunkState &s = st[c];
            const uint32_t phase = sync_load(&s.phase);
            if (phase == PH_RUNNING || phase == PH_LOCKED) break; // its completion continues from here
            if (phase == PH_QUEUED) {
                // not started yet: aim it at the exact entry instead of a speculative one
                if (sync_cas(&s.phase, PH_QUEUED, PH_LOCKED) == PH_QUEUED) {
                    sync_store(&s.new_entry, e);
                    sync_store8(&s.exact, 1);
                    sync_fence();
                    sync_exch(&s.phase, PH_QUEUED);
                }
                break; // (a lost CAS means a worker has just claimed it)
            }
            sync_fence(); // PH_DONE: the chain's results are visible
            if (sync_load(&s.entry) == e && sync_load(&s.exit) != IDX_NONE) {
                e = sync_load(&s.exit);
                c++;
                continue;
            }
            // wrong entry, or the right one but cut short by the budget of a speculative chain: run it from the exact entry
            sync_store(&s.new_entry, e);
            sync_store8(&s.exact, 1);
            sync_store8(&s.dirty, 1);
            sync_fence();
            sync_exch(&s.phase, PH_QUEUED);
            push(c);
            break;
        }
        sync_store(&us.next_c, c);
        sync_store(&us.entry, e);
    }
    sync_fence();
    sync_exch(&us.lock, 0u);
    return became_final;
}

// After the fixpoint: marks skipped chunks, links every chunk to the start of the next model in the
// unit (the end of its last model's residual run) and assigns the leading MacaqueV row.
MDB_DEV void spec_finalize_unit(uint32_t n, uint32_t chunk_len, uint32_t n_chunks, ChunkState *st, uint8_t &unit_irregular) {
    uint32_t e = 0;
    int64_t last_nonempty = -1;
    uint8_t irregular = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        ChunkState &s = st[c];
        uint32_t chunk_end = (uint64_t)(c + 1) * chunk_len < n ? (c + 1) * chunk_len : n;
        irregular |= s.irregular;
        s.lead_end = IDX_NONE;
        s.next_start = n;
        s.rows = 0;
        if (e >= chunk_end) { s.skipped = 1; continue; }
        s.skipped = 0;
        e = s.exit;
        if (s.n_models == 0) continue;
        const uint32_t first_start = s.first_start;
        if (last_nonempty >= 0) st[last_nonempty].next_start = first_start;
        else if (first_start > 0) s.lead_end = first_start - 1; // compression.rs:350-361: leading residuals
        last_nonempty = c;
    }
    if (last_nonempty < 0 && n > 0) { // no model anywhere: the whole unit is one MacaqueV row
        st[0].skipped = 0;
        st[0].n_models = 0;
        st[0].lead_end = n - 1;
    }
    unit_irregular = irregular;
}

// Number of segment rows a chunk emits (store_compressed_segments_with_model_and_or_residuals,
// compression.rs:310-362): one per model, one more when a model is followed by > 255 residuals, plus
// the leading MacaqueV row if the chunk owns it.
MDB_DEV uint32_t spec_count_rows(const ChunkState &s, const FittedModel *list) {
    if (s.skipped) return 0;
    uint32_t rows = s.lead_end != IDX_NONE ? 1 : 0;
    for (uint32_t k = 0; k < s.n_models; k++) {
        uint32_t next_start = k + 1 < s.n_models ? list[k + 1].start_index : s.next_start;
        uint32_t res_end = next_start - 1;
        rows += (res_end - list[k].end_index <= RESIDUAL_VALUES_MAX_LENGTH) ? 1 : 2;
    }
    return rows;
}

// Writes the chunk's SegRecords in final row order. Returns the number written (== spec_count_rows).
MDB_DEV uint32_t spec_records(const ErrorBound &eb, const int64_t *ts, const float *values, const ChunkState &s,
                              const FittedModel *list, bool unit_regular, SegRecord *recs, uint32_t defer_min = 0) {
    if (s.skipped) return 0;
    uint32_t r = 0;
    if (s.lead_end != IDX_NONE) record_macaque_v_segment(eb, 0, s.lead_end, ts, values, unit_regular, recs[r++], defer_min);
    for (uint32_t k = 0; k < s.n_models; k++) {
        uint32_t next_start = k + 1 < s.n_models ? list[k + 1].start_index : s.next_start;
        r += store_segments(eb, true, list[k], next_start - 1, ts, values, unit_regular, recs + r, defer_min);
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// pass 2: emit one segment row
// ------------------------------------------------------------------------------------------------

MDB_DEV void compress_emit_segment(const ErrorBound &eb, const SegRecord &rec, const int64_t *ts, const float *values,
                                   uint8_t *ts_out, uint8_t *val_out, uint8_t *res_out) {
    // timestamps (timestamps.rs:56-155)
    uint64_t n = (uint64_t)rec.res_end_index - rec.start_index + 1;
    if (rec.ts_len) {
        if (rec.regular) {
            for (uint32_t k = 0; k < rec.ts_len; k++) ts_out[rec.ts_len - 1 - k] = (uint8_t)(n >> (8 * k));
        } else {
            BitWriter w(ts_out);
            compress_irregular_residual_timestamps(ts + rec.start_index, n, w);
            w.finish(true);
        }
    }
    // values
    if (rec.model_type_id == MACAQUE_V && rec.wide) {
        // written by k_emit_macaque_warp
    } else if (rec.model_type_id == MACAQUE_V) {
        BitWriter w(val_out);
        float mn, mx;
        macaque_v_encode(eb, values, rec.start_index, rec.res_end_index, false, 0.0f, w, mn, mx);
        w.finish(false);
    } else {
        for (uint32_t k = 0; k < rec.val_len; k++) val_out[k] = rec.values[k];
    }
    // residuals (types.rs:219-254)
    if (rec.res_len) {
        BitWriter w(res_out);
        float mn, mx;
        macaque_v_encode(eb, values, rec.model_end_index + 1, rec.res_end_index, true, rec.model_last_value, w, mn, mx);
        w.finish(false);
        res_out[rec.res_len - 1] = (uint8_t)(rec.res_end_index - rec.model_end_index);
    }
}

} // namespace mdb
