// mdb_grid.cuh -- per-thread bodies of the grid (decompression) kernels, K2.
//
// grid() of the reference (models/mod.rs:190-251) is split three ways so that the bulk of the output
// is written by a streaming, perfectly coalesced kernel:
//   1. grid_prepare_segment: one thread per segment row parses the row once into a 40-byte SegDesc
//      (point count, sampling interval, line coefficients) -- everything that needs a division.
//   2. grid_point: one thread per OUTPUT POINT (tile kernel); for regular-timestamp segments the
//      timestamp is start + j * interval and PMC-Mean / Swing values are closed forms of j, so points
//      are independent of each other and of segment boundaries.
//   3. grid_sequential_segment: what is inherently a serial state machine per row -- irregular
//      (delta-of-delta) timestamps, MacaqueV values and MacaqueV residuals.
#pragma once

#include "mdb_device.cuh"

namespace mdb {

constexpr uint32_t F_TYPE_MASK = 3;      // model type id
constexpr uint32_t F_REGULAR = 4;        // timestamps are start + j * interval
constexpr uint32_t F_TILE_VALUES = 8;    // model values are written by the tile kernel
constexpr uint32_t F_HAS_RESIDUALS = 16;
constexpr uint32_t F_SEQUENTIAL = 32;    // row needs grid_sequential_segment
constexpr uint32_t F_MALFORMED = 64;

struct SegDesc {      // 40 bytes
    int64_t start;
    int64_t interval;
    double a;         // Swing: slope. PMC-Mean: the value (as double bits of a float; exact)
    double b;         // Swing: intercept
    uint32_t model_len;
    uint32_t flags;
};

// Parses row s. Returns the number of data points of the row (len(), models/mod.rs:98-124), 0 if the
// row is one the reference would panic on (flags has F_MALFORMED).
MDB_DEV uint32_t grid_prepare_segment(const SegmentsView &v, uint64_t s, SegDesc &d) {
    Row r = load_row(v, s);
    d.start = r.start_time;
    d.interval = 0;
    d.a = 0.0;
    d.b = 0.0;
    d.model_len = 0;
    d.flags = (uint32_t)(r.model_type_id & 3);
    if (!row_is_well_formed(r)) { d.flags = F_MALFORMED; return 0; }

    uint64_t res_len = r.n_residuals ? r.residuals[r.n_residuals - 1] : 0; // models/mod.rs:277-284
    bool regular = are_compressed_timestamps_regular(r.timestamps, r.n_timestamps);
    uint64_t len;
    if (r.n_timestamps == 0) {
        // timestamps.rs:169-175: one point if start == end, else two
        len = r.start_time == r.end_time ? 1 : 2;
        d.interval = r.end_time - r.start_time;
    } else if (regular) {
        // timestamps.rs:207-223: length is stored, the interval is re-derived by integer division
        len = be_bytes_to_u64(r.timestamps, r.n_timestamps);
        if (len < 2) { d.flags = F_MALFORMED; return 0; }
        uint64_t span = (uint64_t)(r.end_time - r.start_time);
        uint64_t interval = span / (len - 1);
        // (start..=end).step_by(interval) must yield exactly `len` items, else len() and grid() of the
        // reference disagree (and step_by(0) panics)
        if (interval == 0 || span / interval + 1 != len) { d.flags = F_MALFORMED; return 0; }
        d.interval = (int64_t)interval;
    } else {
        len = segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
    }
    if (res_len >= len || len > 0xFFFFFFFFull) { d.flags = F_MALFORMED; return 0; } // model needs >= 1 point
    uint64_t model_len = len - res_len;
    d.model_len = (uint32_t)model_len;

    uint32_t flags = (uint32_t)(r.model_type_id & 3);
    if (regular) flags |= F_REGULAR;
    if (res_len) flags |= F_HAS_RESIDUALS;
    if (r.model_type_id == PMC_MEAN) {
        float value = decode_values_for_pmc_mean(r.min_value, r.max_value, r.values, (uint32_t)r.n_values);
        d.a = (double)value;
        if (regular) flags |= F_TILE_VALUES;
    } else if (r.model_type_id == SWING) {
        if (regular) {
            float first, last;
            decode_values_for_swing(r.min_value, r.max_value, r.values, (uint32_t)r.n_values, first, last);
            // models/mod.rs:223-234: the line ends at the MODEL's last timestamp, not the segment's
            int64_t model_end_time = r.start_time + (int64_t)(model_len - 1) * d.interval;
            compute_slope_and_intercept(r.start_time, (double)first, model_end_time, (double)last, d.a, d.b);
            flags |= F_TILE_VALUES;
        }
    }
    if (!regular || res_len || r.model_type_id == MACAQUE_V) flags |= F_SEQUENTIAL;
    d.flags = flags;
    return (uint32_t)len;
}

// Point j of a regular segment (tile kernel). Writes the timestamp always and the value when the model
// part is a closed form; MacaqueV values and residual values are left to grid_sequential_segment.
MDB_DEV void grid_point(const SegDesc &d, uint32_t j, int64_t *ts_out, float *val_out, uint64_t p) {
    if (!(d.flags & F_REGULAR)) return;
    int64_t t = d.start + (int64_t)j * d.interval;
    ts_out[p] = t;
    if ((d.flags & F_TILE_VALUES) && j < d.model_len) {
        if ((d.flags & F_TYPE_MASK) == PMC_MEAN) val_out[p] = (float)d.a;       // pmc_mean.rs:104-108
        else val_out[p] = swing_value(d.a, d.b, t);                              // swing.rs:304-319
    }
}

// The serial parts of row s; `base` is the row's first output position.
MDB_DEV void grid_sequential_segment(const SegmentsView &v, uint64_t s, const SegDesc &d, uint64_t base,
                                     uint32_t len, int64_t *ts_out, float *val_out) {
    Row r = load_row(v, s);
    const uint32_t model_len = d.model_len;
    const int type = (int)(d.flags & F_TYPE_MASK);
    float last_model_value = 0.0f;

    if (!(d.flags & F_REGULAR)) {
        // timestamps.rs:228-275: start_time, decoded residual timestamps, end_time
        IrregularTimestampDecoder dec;
        dec.init(r.start_time, r.timestamps, r.n_timestamps);
        ts_out[base] = r.start_time;
        uint32_t j = 1;
        while (j + 1 < len && dec.next()) ts_out[base + j++] = dec.timestamp;
        ts_out[base + len - 1] = r.end_time;
        if (type == PMC_MEAN) {
            float value = (float)d.a;
            for (uint32_t k = 0; k < model_len; k++) val_out[base + k] = value;
            last_model_value = value;
        } else if (type == SWING) {
            float first, last;
            decode_values_for_swing(r.min_value, r.max_value, r.values, (uint32_t)r.n_values, first, last);
            int64_t model_end_time = ts_out[base + model_len - 1]; // written above by this thread
            double slope, intercept;
            compute_slope_and_intercept(r.start_time, (double)first, model_end_time, (double)last, slope, intercept);
            for (uint32_t k = 0; k < model_len; k++) {
                last_model_value = swing_value(slope, intercept, ts_out[base + k]);
                val_out[base + k] = last_model_value;
            }
        }
    } else if (type == PMC_MEAN) {
        last_model_value = (float)d.a;
    } else if (type == SWING) {
        last_model_value = swing_value(d.a, d.b, d.start + (int64_t)(model_len - 1) * d.interval);
    }

    if (type == MACAQUE_V) { // macaque_v.rs:272-323
        MacaqueVDecoder dec;
        dec.init(r.values, r.n_values, false, 0.0f);
        last_model_value = __uint_as_float(dec.last_value);
        val_out[base] = last_model_value;
        for (uint32_t k = 1; k < model_len; k++) {
            last_model_value = dec.next();
            val_out[base + k] = last_model_value;
        }
    }

    if (d.flags & F_HAS_RESIDUALS) {
        // models/mod.rs:241-249: residuals are seeded with the last GRIDDED model value (quirk Q1)
        MacaqueVDecoder dec;
        dec.init(r.residuals, r.n_residuals - 1, true, last_model_value);
        for (uint32_t k = model_len; k < len; k++) val_out[base + k] = dec.next();
    }
}

} // namespace mdb
