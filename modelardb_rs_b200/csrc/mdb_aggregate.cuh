// mdb_aggregate.cuh -- per-thread body of the aggregate kernel, K3: COUNT and SUM of one segment row
// computed from the model without materialising its data points (models/mod.rs:98-184,
// swing.rs:264-300, pmc_mean.rs:98-100, macaque_v.rs:220-265).  MIN / MAX are the row's metadata
// columns and are folded by the group reduction.
#pragma once

#include "mdb_device.cuh"

namespace mdb {

// swing.rs:264-300.  Note quirk Q2: the slope is taken to the SEGMENT end_time (which includes the
// residual timestamps), not to the model's last timestamp as grid() does.
MDB_DEV float swing_sum(const Row &r, float first_value, float last_value, uint64_t length, uint64_t res_len) {
    double slope, intercept;
    compute_slope_and_intercept(r.start_time, (double)first_value, r.end_time, (double)last_value, slope, intercept);
    if (are_compressed_timestamps_regular(r.timestamps, r.n_timestamps)) {
        double first = __dadd_rn(__dmul_rn(slope, (double)r.start_time), intercept);
        double last = __dadd_rn(__dmul_rn(slope, (double)r.end_time), intercept);
        double average = __ddiv_rn(__dadd_rn(first, last), 2.0);
        return __double2float_rn(__dmul_rn(average, __ull2double_rn(length - res_len)));
    }
    // irregular: sum the line over the model's timestamps in order, in f64
    uint64_t model_len = length - res_len;
    double sum = 0.0;
    IrregularTimestampDecoder dec;
    dec.init(r.start_time, r.timestamps, r.n_timestamps);
    for (uint64_t k = 0; k < model_len; k++) {
        int64_t t;
        if (k == 0) t = r.start_time;
        else if (k + 1 == length) t = r.end_time;
        else { dec.next(); t = dec.timestamp; }
        sum = __dadd_rn(sum, __dadd_rn(__dmul_rn(slope, (double)t), intercept));
    }
    return __double2float_rn(sum);
}

// macaque_v.rs:220-265: sequential f32 accumulation in stream order.
MDB_DEV float macaque_v_sum(const uint8_t *bytes, uint64_t n_bytes, uint64_t length, bool has_seed, float seed) {
    if (length == 0) return 0.0f;
    MacaqueVDecoder dec;
    dec.init(bytes, n_bytes, has_seed, seed);
    float sum;
    uint64_t remaining;
    if (has_seed) { sum = 0.0f; remaining = length; }
    else { sum = __uint_as_float(dec.last_value); remaining = length - 1; }
    for (uint64_t k = 0; k < remaining; k++) sum = __fadd_rn(sum, dec.next());
    return sum;
}

// COUNT (len) and SUM (sum) of row s. Returns false for a row the reference would panic on.
// defer_min / deferred: a MacaqueV row of at least defer_min (> 0) model values only gets its COUNT here and is
// reported through *deferred; its SUM is left to the kernel that decodes long streams with a whole warp.
// count_only: stop once it is known whether the row is deferred (k_agg_find_wide).
MDB_DEV bool aggregate_segment(const SegmentsView &v, uint64_t s, uint64_t &count, float &sum, uint32_t defer_min = 0,
                               bool *deferred = nullptr, bool count_only = false) {
    Row r = load_row(v, s);
    count = 0;
    sum = 0.0f;
    if (!row_is_well_formed(r)) return false;
    uint64_t res_len = r.n_residuals ? r.residuals[r.n_residuals - 1] : 0;
    uint64_t length = segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
    if (res_len >= length) return false;
    uint64_t model_length = length - res_len;
    float model_last_value, model_sum;
    if (r.model_type_id == PMC_MEAN) {
        float value = decode_values_for_pmc_mean(r.min_value, r.max_value, r.values, (uint32_t)r.n_values);
        model_last_value = value;
        model_sum = __fmul_rn(__ull2float_rn(model_length), value); // pmc_mean.rs:98-100
    } else if (r.model_type_id == SWING) {
        float first, last;
        decode_values_for_swing(r.min_value, r.max_value, r.values, (uint32_t)r.n_values, first, last);
        model_last_value = last;
        model_sum = swing_sum(r, first, last, length, res_len);
    } else {
        if (defer_min && model_length >= defer_min) {
            count = length;
            *deferred = true;
            return true;
        }
        if (count_only) return true;
        model_last_value = __uint_as_float(0x7fc00000u);
        model_sum = macaque_v_sum(r.values, r.n_values, model_length, false, 0.0f);
    }
    count = length;
    if (r.n_residuals == 0) { sum = canonical_nan(model_sum); return true; }
    // models/mod.rs:173-183: residuals are seeded with the DECODED model last value here (quirk Q1)
    float residuals_sum = macaque_v_sum(r.residuals, r.n_residuals - 1, res_len, true, model_last_value);
    sum = canonical_nan(__fadd_rn(model_sum, residuals_sum));
    return true;
}

// Accumulator state of one group (model_simple_aggregates.rs:336-618).
struct GroupAgg {
    int64_t count;
    float min, max;
    double sum;
};

MDB_DEV GroupAgg group_agg_identity() {
    GroupAgg g;
    g.count = 0;
    g.min = 3.402823466e+38f;   // f32::MAX, model_simple_aggregates.rs:97, :413
    g.max = -3.402823466e+38f;  // f32::MIN, model_simple_aggregates.rs:117, :456
    g.sum = 0.0;
    return g;
}

// Fold `b` (later rows) into `a` (earlier rows).
MDB_DEV GroupAgg group_agg_combine(const GroupAgg &a, const GroupAgg &b) {
    GroupAgg g;
    g.count = a.count + b.count;
    g.min = rust_minf(a.min, b.min);
    g.max = rust_maxf(a.max, b.max);
    g.sum = __dadd_rn(a.sum, b.sum);
    return g;
}

} // namespace mdb
