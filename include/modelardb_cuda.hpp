// modelardb_cuda.hpp -- C++ host side above the C-ABI (modelardb_cuda.h), mirroring the reference's own interfaces
// for the hot path: same names, argument meaning and error behaviour as the Rust functions they stand in for.
//
//   modelardb_compression (crates/modelardb_compression/src/lib.rs:26-34)
//     ErrorBound                                    crates/modelardb_types/src/types.rs:299-335
//     try_compress_univariate_time_series           compression.rs:191-275
//     try_split_and_compress_univariate_time_series compression.rs:147-179   (batch form: many units, one call)
//     len / sum / grid                              models/mod.rs:98-124 / 129-184 / 190-251
//   query operators
//     GridStream                                    crates/modelardb_storage/src/query/grid_exec.rs:197-430
//     Model{Count,Min,Max,Sum,Avg}Accumulator       crates/modelardb_storage/src/optimizer/model_simple_aggregates.rs:336-618
//     grouped_model_aggregates                      the aggregate rule (:203-334) extended to GROUP BY <tag columns>
//   server compressor
//     compress_finished_buffers                     crates/modelardb_server/src/storage/uncompressed_data_manager.rs:530-581
//
// The reference is Rust and its toolchain is not in this image, so the host side is C++ (the Python package under
// modelardb_rs_b200/ is the same mirror for the tests and the benchmark).  Header only; link libmodelardb_cuda.so.
// Every function forwards to the CUDA library: there is no CPU implementation here either.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "modelardb_cuda.h"

namespace modelardb_cuda {

// ModelarDbCompressionError::InvalidArgument and friends (crates/modelardb_compression/src/error.rs): one exception type,
// carrying the library's last error message.
class Error : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};

inline void check(int rc) {
    if (rc != MDBCU_SUCCESS) throw Error(mdbcu_last_error());
}

// crates/modelardb_types/src/types.rs:299-335
class ErrorBound {
public:
    static ErrorBound lossless() { return ErrorBound(MDBCU_LOSSLESS, 0.0f); }
    static ErrorBound try_new_absolute(float value) {
        if (!(value > 0.0f) || !std::isfinite(value)) throw Error("An absolute error bound must be a positive finite value.");
        return ErrorBound(MDBCU_ABSOLUTE, value);
    }
    static ErrorBound try_new_relative(float percentage) {
        if (!(percentage > 0.0f && percentage <= 100.0f)) throw Error("A relative error bound must be a positive value that is at most 100.0%.");
        return ErrorBound(MDBCU_RELATIVE, percentage);
    }
    uint8_t kind() const { return kind_; }
    float value() const { return value_; }

private:
    ErrorBound(uint8_t kind, float value) : kind_(kind), value_(value) {}
    uint8_t kind_;
    float value_;
};

// One GPU, one stream.  Calls on a context are blocking; use one context per host thread.
class Context {
public:
    explicit Context(int device = 0) { check(mdbcu_context_create(device, &ctx_)); }
    ~Context() { mdbcu_context_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    mdbcu_context *get() const { return ctx_; }

private:
    mdbcu_context *ctx_ = nullptr;
};

// The columns of QUERY_COMPRESSED_SCHEMA (crates/modelardb_types/src/schemas.rs:40-52) for a batch of segments; the
// three binary columns as offsets + bytes.  `error` is constant NaN and field_column / tags are per-batch constants
// (types.rs:492-516), so they are not materialised.
struct CompressedSegmentBatch {
    std::vector<int8_t> model_type_ids;
    std::vector<int64_t> start_times, end_times;
    std::vector<float> min_values, max_values;
    std::vector<uint64_t> timestamps_off{0}, values_off{0}, residuals_off{0};
    std::vector<uint8_t> timestamps, values, residuals;

    size_t num_rows() const { return model_type_ids.size(); }
    mdbcu_segments_view view() const {
        mdbcu_segments_view v;
        v.n_segments = num_rows();
        v.model_type_id = model_type_ids.data();
        v.start_time = start_times.data();
        v.end_time = end_times.data();
        v.min_value = min_values.data();
        v.max_value = max_values.data();
        v.timestamps_off = timestamps_off.data();
        v.timestamps_data = timestamps.data();
        v.values_off = values_off.data();
        v.values_data = values.data();
        v.residuals_off = residuals_off.data();
        v.residuals_data = residuals.data();
        return v;
    }
    // Rows [lo, hi) as an independent batch.
    CompressedSegmentBatch slice(size_t lo, size_t hi) const {
        CompressedSegmentBatch out;
        out.model_type_ids.assign(model_type_ids.begin() + lo, model_type_ids.begin() + hi);
        out.start_times.assign(start_times.begin() + lo, start_times.begin() + hi);
        out.end_times.assign(end_times.begin() + lo, end_times.begin() + hi);
        out.min_values.assign(min_values.begin() + lo, min_values.begin() + hi);
        out.max_values.assign(max_values.begin() + lo, max_values.begin() + hi);
        auto cut = [&](const std::vector<uint64_t> &off, const std::vector<uint8_t> &data, std::vector<uint64_t> &o, std::vector<uint8_t> &d) {
            o.clear();
            for (size_t r = lo; r <= hi; r++) o.push_back(off[r] - off[lo]);
            d.assign(data.begin() + off[lo], data.begin() + off[hi]);
        };
        cut(timestamps_off, timestamps, out.timestamps_off, out.timestamps);
        cut(values_off, values, out.values_off, out.values);
        cut(residuals_off, residuals, out.residuals_off, out.residuals);
        return out;
    }
    // The rows with keep[row] != 0, in their original order.
    CompressedSegmentBatch take(const std::vector<uint8_t> &keep) const {
        CompressedSegmentBatch out;
        for (size_t row = 0; row < num_rows(); row++) {
            if (!keep[row]) continue;
            out.model_type_ids.push_back(model_type_ids[row]);
            out.start_times.push_back(start_times[row]);
            out.end_times.push_back(end_times[row]);
            out.min_values.push_back(min_values[row]);
            out.max_values.push_back(max_values[row]);
            auto append = [&](const std::vector<uint64_t> &off, const std::vector<uint8_t> &data, std::vector<uint64_t> &o, std::vector<uint8_t> &d) {
                d.insert(d.end(), data.begin() + off[row], data.begin() + off[row + 1]);
                o.push_back(d.size());
            };
            append(timestamps_off, timestamps, out.timestamps_off, out.timestamps);
            append(values_off, values, out.values_off, out.values);
            append(residuals_off, residuals, out.residuals_off, out.residuals);
        }
        return out;
    }
};

namespace detail {
inline CompressedSegmentBatch copy_out(mdbcu_segments *segments, uint64_t n_units, std::vector<uint64_t> *unit_seg_off) {
    mdbcu_segments_view v;
    const uint64_t *uso = nullptr;
    int rc = mdbcu_segments_get(segments, MDBCU_HOST, &v, &uso);
    CompressedSegmentBatch out;
    if (rc == MDBCU_SUCCESS) {
        const uint64_t s = v.n_segments;
        out.model_type_ids.assign(v.model_type_id, v.model_type_id + s);
        out.start_times.assign(v.start_time, v.start_time + s);
        out.end_times.assign(v.end_time, v.end_time + s);
        out.min_values.assign(v.min_value, v.min_value + s);
        out.max_values.assign(v.max_value, v.max_value + s);
        out.timestamps_off.assign(v.timestamps_off, v.timestamps_off + s + 1);
        out.values_off.assign(v.values_off, v.values_off + s + 1);
        out.residuals_off.assign(v.residuals_off, v.residuals_off + s + 1);
        out.timestamps.assign(v.timestamps_data, v.timestamps_data + out.timestamps_off[s]);
        out.values.assign(v.values_data, v.values_data + out.values_off[s]);
        out.residuals.assign(v.residuals_data, v.residuals_data + out.residuals_off[s]);
        if (unit_seg_off) unit_seg_off->assign(uso, uso + n_units + 1);
    }
    mdbcu_segments_free(segments);
    check(rc);
    return out;
}
} // namespace detail

// compression.rs:191-275.  Different lengths -> InvalidArgument (:202-206); empty input -> empty batch (:208-211).
inline CompressedSegmentBatch try_compress_univariate_time_series(Context &ctx, const std::vector<int64_t> &uncompressed_timestamps,
                                                                  const std::vector<float> &uncompressed_values, ErrorBound error_bound) {
    if (uncompressed_timestamps.size() != uncompressed_values.size())
        throw Error("Uncompressed timestamps and uncompressed values have different lengths.");
    const uint64_t unit_off[2] = {0, uncompressed_timestamps.size()};
    const uint8_t kind = error_bound.kind();
    const float value = error_bound.value();
    mdbcu_segments *segments = nullptr;
    check(mdbcu_compress(ctx.get(), MDBCU_HOST, uncompressed_timestamps.data(), uncompressed_values.data(), unit_off, 1, &kind, &value, &segments));
    return detail::copy_out(segments, 1, nullptr);
}

// The batch form behind try_split_and_compress_univariate_time_series (compression.rs:147-179) and the server's
// compressor (uncompressed_data_manager.rs:563-581): unit u is [unit_off[u], unit_off[u+1]) of the two arrays, with
// its own bound; rows come back ordered by unit, unit_seg_off (optional) gives the rows of each unit.
inline CompressedSegmentBatch try_compress_time_series_batch(Context &ctx, const std::vector<int64_t> &timestamps, const std::vector<float> &values,
                                                             const std::vector<uint64_t> &unit_off, const std::vector<ErrorBound> &error_bounds,
                                                             std::vector<uint64_t> *unit_seg_off = nullptr) {
    if (timestamps.size() != values.size()) throw Error("Uncompressed timestamps and uncompressed values have different lengths.");
    if (unit_off.empty() || unit_off.size() - 1 != error_bounds.size()) throw Error("One error bound per unit is required.");
    if (unit_off.back() > timestamps.size()) throw Error("unit_off exceeds the number of data points.");
    std::vector<uint8_t> kinds;
    std::vector<float> bounds;
    for (const ErrorBound &eb : error_bounds) {
        kinds.push_back(eb.kind());
        bounds.push_back(eb.value());
    }
    mdbcu_segments *segments = nullptr;
    check(mdbcu_compress(ctx.get(), MDBCU_HOST, timestamps.data(), values.data(), unit_off.data(), error_bounds.size(), kinds.data(), bounds.data(),
                         &segments));
    return detail::copy_out(segments, error_bounds.size(), unit_seg_off);
}

// models/mod.rs:98-124 for every row of a batch, as an exclusive prefix sum (point_off[i+1] - point_off[i] = len of row i).
inline std::vector<uint64_t> len(Context &ctx, const CompressedSegmentBatch &batch) {
    std::vector<uint64_t> point_off(batch.num_rows() + 1, 0);
    uint64_t total = 0;
    const mdbcu_segments_view v = batch.view();
    check(mdbcu_grid_count(ctx.get(), MDBCU_HOST, &v, point_off.data(), &total));
    return point_off;
}

// models/mod.rs:129-184 for every row of a batch.
inline std::vector<float> sum(Context &ctx, const CompressedSegmentBatch &batch) {
    std::vector<float> sums(batch.num_rows());
    const mdbcu_segments_view v = batch.view();
    check(mdbcu_segment_sums(ctx.get(), MDBCU_HOST, &v, sums.data()));
    return sums;
}

// models/mod.rs:190-251 for every row of a batch: APPENDS to the two builders like the reference does.
inline void grid(Context &ctx, const CompressedSegmentBatch &batch, std::vector<int64_t> &timestamp_builder, std::vector<float> &value_builder,
                 std::vector<uint64_t> *point_off_out = nullptr) {
    std::vector<uint64_t> point_off = len(ctx, batch);
    const uint64_t total = point_off.back(), before = timestamp_builder.size();
    timestamp_builder.resize(before + total);
    value_builder.resize(before + total);
    uint64_t n = 0;
    const mdbcu_segments_view v = batch.view();
    check(mdbcu_grid(ctx.get(), MDBCU_HOST, &v, timestamp_builder.data() + before, value_builder.data() + before, total, &n));
    if (point_off_out) *point_off_out = std::move(point_off);
}

// grid() with the predicate `t_lo <= timestamp AND timestamp <= t_hi` evaluated inside the call (mdbcu_grid_range): the
// reference reconstructs every point of the selected segments and prunes afterwards (grid_exec.rs:366-387).  APPENDS the
// surviving points; point_off_out receives their exclusive prefix sum per row (0 points for rows outside the range).
inline void grid_range(Context &ctx, const CompressedSegmentBatch &batch, int64_t t_lo, int64_t t_hi, std::vector<int64_t> &timestamp_builder,
                       std::vector<float> &value_builder, std::vector<uint64_t> *point_off_out = nullptr) {
    const mdbcu_segments_view v = batch.view();
    uint64_t n = 0;
    check(mdbcu_grid_range(ctx.get(), MDBCU_HOST, &v, t_lo, t_hi, nullptr, nullptr, nullptr, 0, &n)); // a count
    const uint64_t before = timestamp_builder.size();
    timestamp_builder.resize(before + n);
    value_builder.resize(before + n);
    std::vector<uint64_t> point_off(batch.num_rows() + 1, 0);
    check(mdbcu_grid_range(ctx.get(), MDBCU_HOST, &v, t_lo, t_hi, point_off.data(), n ? timestamp_builder.data() + before : nullptr,
                           n ? value_builder.data() + before : nullptr, n, &n));
    if (point_off_out) *point_off_out = std::move(point_off);
}

// grid_exec.rs:197-430: leftovers of the current batch + the points of the next segment batch, handed out in slices
// of batch_size rows; `tags` of a segment batch (one value per row and tag column) are repeated for every created row.
class GridStream {
public:
    struct Batch {
        std::vector<int64_t> timestamps;
        std::vector<float> values;
        std::vector<std::vector<std::string>> tags; // one vector per tag column
    };
    using Input = std::pair<CompressedSegmentBatch, std::vector<std::vector<std::string>>>; // (segments, tag columns)
    // keep[i] != 0 for the reconstructed points that survive (the reference's maybe_predicate, grid_exec.rs:368-386)
    using Predicate = std::function<std::vector<uint8_t>(const std::vector<int64_t> &, const std::vector<float> &)>;

    struct Options {
        size_t batch_size = 8192;
        size_t n_tag_columns = 0;
        Predicate predicate;
        // Segments that end before the first or start after the second are not reconstructed at all (the push-down the
        // reference applies to its Parquet scan, time_series_table.rs:290-373); the predicate must imply the range.
        std::optional<int64_t> time_range_start, time_range_end;
        // With a time range: ALSO drop the points outside it inside the reconstruction call (grid_range) instead of
        // reconstructing every point of the surviving segments and pruning afterwards (grid_exec.rs:366-387).  Since the
        // predicate implies the range the stream's rows are unchanged; only rows_created() shrinks.
        bool device_time_clip = false;
        // batch_size becomes min(limit, batch_size) as in the reference (grid_exec.rs:239-246); the stream also ends after
        // `limit` rows and, without a predicate, reconstructs only the leading segments needed to reach it.
        std::optional<size_t> limit;
    };

    GridStream(Context &ctx, std::vector<Input> input, size_t batch_size, size_t n_tag_columns = 0)
        : GridStream(ctx, std::move(input), make_options(batch_size, n_tag_columns)) {}

    GridStream(Context &ctx, std::vector<Input> input, Options options)
        : ctx_(ctx), input_(std::move(input)), options_(std::move(options)), current_{{}, {}, std::vector<std::vector<std::string>>(options_.n_tag_columns)} {
        if (options_.batch_size == 0) throw Error("batch_size must be positive");
        if (options_.limit) {
            if (*options_.limit == 0) throw Error("limit must be positive");
            options_.batch_size = std::min(options_.batch_size, *options_.limit);
        }
    }

    // Poll::Ready(Some(batch)) -> true (the batch may be empty, as in the reference); Poll::Ready(None) -> false.
    bool poll_next(Batch &out) {
        if (options_.limit && handed_out_ >= *options_.limit) return false;
        if (remaining() < options_.batch_size && next_input_ < input_.size()) grid_and_append_to_leftovers_in_current_batch(input_[next_input_++]);
        if (next_input_ >= input_.size() && remaining() == 0) return false;
        size_t length = std::min(options_.batch_size, remaining());
        if (options_.limit) length = std::min(length, *options_.limit - handed_out_);
        out.timestamps.assign(current_.timestamps.begin() + offset_, current_.timestamps.begin() + offset_ + length);
        out.values.assign(current_.values.begin() + offset_, current_.values.begin() + offset_ + length);
        out.tags.assign(current_.tags.size(), {});
        for (size_t c = 0; c < current_.tags.size(); c++) out.tags[c].assign(current_.tags[c].begin() + offset_, current_.tags[c].begin() + offset_ + length);
        offset_ += length;
        handed_out_ += length;
        return true;
    }

    size_t batch_size() const { return options_.batch_size; }
    size_t segments_skipped() const { return segments_skipped_; } // by the time range or the limit
    uint64_t rows_created() const { return rows_created_; }       // GridStreamMetrics (grid_exec.rs:433-520)

private:
    static Options make_options(size_t batch_size, size_t n_tag_columns) {
        Options o;
        o.batch_size = batch_size;
        o.n_tag_columns = n_tag_columns;
        return o;
    }
    size_t remaining() const { return current_.timestamps.size() - offset_; }

    void grid_and_append_to_leftovers_in_current_batch(const Input &in) { // grid_exec.rs:261-391
        if (in.second.size() != current_.tags.size()) throw Error("every segment batch must carry the same tag columns");
        const CompressedSegmentBatch *segments = &in.first;
        const std::vector<std::vector<std::string>> *tags = &in.second;
        CompressedSegmentBatch kept;
        std::vector<std::vector<std::string>> kept_tags;
        auto keep_rows = [&](const std::vector<uint8_t> &keep) {
            CompressedSegmentBatch next_kept = segments->take(keep);
            std::vector<std::vector<std::string>> next_tags(tags->size());
            for (size_t c = 0; c < tags->size(); c++)
                for (size_t row = 0; row < keep.size(); row++)
                    if (keep[row]) next_tags[c].push_back((*tags)[c][row]);
            segments_skipped_ += keep.size() - next_kept.num_rows();
            kept = std::move(next_kept);
            kept_tags = std::move(next_tags);
            segments = &kept;
            tags = &kept_tags;
        };
        if (options_.time_range_start || options_.time_range_end) {
            std::vector<uint8_t> keep(segments->num_rows(), 1);
            bool all = true;
            for (size_t row = 0; row < keep.size(); row++) {
                if (options_.time_range_start && segments->end_times[row] < *options_.time_range_start) keep[row] = 0;
                if (options_.time_range_end && segments->start_times[row] > *options_.time_range_end) keep[row] = 0;
                all = all && keep[row];
            }
            if (!all) keep_rows(keep);
        }
        std::vector<uint64_t> point_off(1, 0);
        if (options_.limit && !options_.predicate && segments->num_rows()) {
            // rows still owed beyond the leftovers: the first segments that cover them are enough
            point_off = len(ctx_, *segments);
            const size_t have = handed_out_ + remaining();
            const uint64_t owed = *options_.limit > have ? *options_.limit - have : 0;
            size_t needed = 0;
            while (needed < segments->num_rows() && point_off[needed] < owed) needed++;
            if (needed < segments->num_rows()) {
                std::vector<uint8_t> keep(segments->num_rows(), 0);
                std::fill(keep.begin(), keep.begin() + needed, 1);
                keep_rows(keep);
            }
        }
        Batch next;
        next.timestamps.assign(current_.timestamps.begin() + offset_, current_.timestamps.end());
        next.values.assign(current_.values.begin() + offset_, current_.values.end());
        next.tags.resize(current_.tags.size());
        for (size_t c = 0; c < current_.tags.size(); c++) next.tags[c].assign(current_.tags[c].begin() + offset_, current_.tags[c].end());
        point_off.assign(1, 0);
        if (segments->num_rows()) {
            if (options_.device_time_clip && (options_.time_range_start || options_.time_range_end))
                grid_range(ctx_, *segments, options_.time_range_start.value_or(std::numeric_limits<int64_t>::min()),
                           options_.time_range_end.value_or(std::numeric_limits<int64_t>::max()), next.timestamps, next.values, &point_off);
            else
                grid(ctx_, *segments, next.timestamps, next.values, &point_off);
        }
        rows_created_ += point_off.back();
        for (size_t c = 0; c < next.tags.size(); c++)
            for (size_t row = 0; row < segments->num_rows(); row++)
                next.tags[c].insert(next.tags[c].end(), point_off[row + 1] - point_off[row], (*tags)[c][row]);
        if (options_.predicate) {
            // (the leftovers were filtered when they were created; the predicate is a per-row test, so filtering them
            // again together with the new points changes nothing)
            const std::vector<uint8_t> keep = options_.predicate(next.timestamps, next.values);
            if (keep.size() != next.timestamps.size()) throw Error("the predicate must return one flag per data point");
            size_t w = 0;
            for (size_t i = 0; i < keep.size(); i++) {
                if (!keep[i]) continue;
                if (w != i) {
                    next.timestamps[w] = next.timestamps[i];
                    next.values[w] = next.values[i];
                    for (auto &column : next.tags) column[w] = std::move(column[i]);
                }
                w++;
            }
            next.timestamps.resize(w);
            next.values.resize(w);
            for (auto &column : next.tags) column.resize(w);
        }
        current_ = std::move(next);
        offset_ = 0;
    }

    Context &ctx_;
    std::vector<Input> input_;
    Options options_;
    size_t next_input_ = 0, offset_ = 0, handed_out_ = 0, segments_skipped_ = 0;
    uint64_t rows_created_ = 0;
    Batch current_;
};

// model_simple_aggregates.rs:336-618: update_batch folds a segment batch into the state, state() returns it and resets.
// (merge_batch / evaluate are unreachable!() on the model accumulators and are not provided.)
namespace detail {
struct BatchAggregate {
    int64_t count;
    float min, max;
    double sum;
};
inline BatchAggregate aggregate(Context &ctx, const CompressedSegmentBatch &batch) {
    BatchAggregate a{0, std::numeric_limits<float>::max(), std::numeric_limits<float>::lowest(), 0.0};
    if (batch.num_rows() == 0) return a;
    const mdbcu_segments_view v = batch.view();
    check(mdbcu_aggregate(ctx.get(), MDBCU_HOST, &v, nullptr, 1, &a.count, &a.min, &a.max, &a.sum));
    return a;
}
// Value::min / Value::max of Rust: NaN-ignoring, the receiver wins ties
inline float rust_min(float a, float b) { return a != a ? b : (b < a ? b : a); }
inline float rust_max(float a, float b) { return a != a ? b : (b > a ? b : a); }
} // namespace detail

class ModelCountAccumulator {
public:
    explicit ModelCountAccumulator(Context &ctx) : ctx_(ctx) {}
    void update_batch(const CompressedSegmentBatch &batch) { count_ += detail::aggregate(ctx_, batch).count; }
    int64_t state() { return std::exchange(count_, 0); }

private:
    Context &ctx_;
    int64_t count_ = 0;
};

class ModelMinAccumulator {
public:
    explicit ModelMinAccumulator(Context &ctx) : ctx_(ctx) {}
    void update_batch(const CompressedSegmentBatch &batch) { min_ = detail::rust_min(min_, detail::aggregate(ctx_, batch).min); }
    float state() { return std::exchange(min_, std::numeric_limits<float>::max()); }

private:
    Context &ctx_;
    float min_ = std::numeric_limits<float>::max();
};

class ModelMaxAccumulator {
public:
    explicit ModelMaxAccumulator(Context &ctx) : ctx_(ctx) {}
    void update_batch(const CompressedSegmentBatch &batch) { max_ = detail::rust_max(max_, detail::aggregate(ctx_, batch).max); }
    float state() { return std::exchange(max_, std::numeric_limits<float>::lowest()); }

private:
    Context &ctx_;
    float max_ = std::numeric_limits<float>::lowest();
};

class ModelSumAccumulator {
public:
    explicit ModelSumAccumulator(Context &ctx) : ctx_(ctx) {}
    void update_batch(const CompressedSegmentBatch &batch) { sum_ += detail::aggregate(ctx_, batch).sum; }
    double state() { return std::exchange(sum_, 0.0); }

private:
    Context &ctx_;
    double sum_ = 0.0;
};

class ModelAvgAccumulator {
public:
    explicit ModelAvgAccumulator(Context &ctx) : ctx_(ctx) {}
    void update_batch(const CompressedSegmentBatch &batch) {
        const detail::BatchAggregate a = detail::aggregate(ctx_, batch);
        sum_ += a.sum;
        count_ += (uint64_t)a.count;
    }
    std::pair<uint64_t, double> state() { return {std::exchange(count_, 0), std::exchange(sum_, 0.0)}; } // (count, sum)

private:
    Context &ctx_;
    double sum_ = 0.0;
    uint64_t count_ = 0;
};

// COUNT / MIN / MAX / SUM per distinct combination of tag values straight from segments (the aggregate rule of
// model_simple_aggregates.rs:203-334 extended to GROUP BY <tag columns>).  Rows with equal tags are
// contiguous in what compress and the storage layer produce, so each run of equal tags is one group of ONE
// mdbcu_aggregate call; runs that repeat an earlier key are merged in row order with the accumulators' folds.
// Keys come back in order of first appearance; AVG = sum / count.
struct GroupedAggregates {
    std::vector<std::vector<std::string>> keys;
    std::vector<int64_t> count;
    std::vector<float> min, max;
    std::vector<double> sum;
};

inline GroupedAggregates grouped_model_aggregates(Context &ctx, const CompressedSegmentBatch &batch,
                                                  const std::vector<std::vector<std::string>> &tag_columns) {
    const size_t n = batch.num_rows();
    for (const auto &column : tag_columns)
        if (column.size() != n) throw Error("a tag column needs one value per segment");
    GroupedAggregates out;
    if (n == 0) return out;
    std::vector<uint64_t> group_off;
    for (size_t row = 0; row < n; row++) {
        bool change = row == 0;
        for (const auto &column : tag_columns) change = change || column[row] != column[row - 1];
        if (change) group_off.push_back(row);
    }
    const size_t runs = group_off.size();
    group_off.push_back(n);
    std::vector<int64_t> count(runs);
    std::vector<float> min(runs), max(runs);
    std::vector<double> sum(runs);
    const mdbcu_segments_view v = batch.view();
    check(mdbcu_aggregate(ctx.get(), MDBCU_HOST, &v, group_off.data(), runs, count.data(), min.data(), max.data(), sum.data()));
    std::map<std::vector<std::string>, size_t> slot_of;
    for (size_t r = 0; r < runs; r++) {
        std::vector<std::string> key;
        for (const auto &column : tag_columns) key.push_back(column[group_off[r]]);
        auto found = slot_of.find(key);
        if (found == slot_of.end()) {
            slot_of.emplace(key, out.keys.size());
            out.keys.push_back(std::move(key));
            out.count.push_back(count[r]);
            out.min.push_back(min[r]);
            out.max.push_back(max[r]);
            out.sum.push_back(sum[r]);
        } else {
            const size_t k = found->second;
            out.count[k] += count[r];
            out.min[k] = detail::rust_min(out.min[k], min[r]);
            out.max[k] = detail::rust_max(out.max[k], max[r]);
            out.sum[k] += sum[r];
        }
    }
    return out;
}

// The server's compressor (crates/modelardb_server/src/storage/uncompressed_data_manager.rs:530-581) takes one finished
// buffer at a time -- the data points of ONE series for all fields of its table -- and compresses each field with its
// own error bound.  One buffer per call would starve a GPU: compress_finished_buffers takes every buffer that is waiting
// and compresses all their (buffer, field) pairs with ONE mdbcu_compress; what comes back is what the
// reference sends on, one CompressedSegmentBatch per buffer and field, in the order of the buffers.
struct UncompressedDataBuffer {
    std::vector<int64_t> timestamps;
    std::vector<std::vector<float>> field_columns; // one per field
    std::vector<int> field_column_indices;         // index of each field in the table's schema
    std::vector<ErrorBound> error_bounds;          // one per field
    std::vector<std::string> tag_values;
};

struct CompressedBuffer {
    std::vector<std::string> tag_values;
    std::vector<std::pair<int, CompressedSegmentBatch>> compressed_segments; // (field_column_index, segments), in field order
};

inline std::vector<CompressedBuffer> compress_finished_buffers(Context &ctx, const std::vector<UncompressedDataBuffer> &buffers) {
    std::vector<int64_t> timestamps;
    std::vector<float> values;
    std::vector<uint64_t> unit_off(1, 0);
    std::vector<ErrorBound> bounds;
    for (const UncompressedDataBuffer &b : buffers) {
        if (b.field_columns.size() != b.field_column_indices.size() || b.field_columns.size() != b.error_bounds.size())
            throw Error("one index and one error bound per field column");
        for (size_t f = 0; f < b.field_columns.size(); f++) {
            if (b.field_columns[f].size() != b.timestamps.size())
                throw Error("Uncompressed timestamps and uncompressed values have different lengths.");
            // the C-ABI pairs one timestamp with every value, so a buffer's timestamps are repeated once per field
            timestamps.insert(timestamps.end(), b.timestamps.begin(), b.timestamps.end());
            values.insert(values.end(), b.field_columns[f].begin(), b.field_columns[f].end());
            unit_off.push_back(timestamps.size());
            bounds.push_back(b.error_bounds[f]);
        }
    }
    std::vector<CompressedBuffer> out;
    std::vector<uint64_t> unit_seg_off(unit_off.size(), 0);
    CompressedSegmentBatch all;
    if (!bounds.empty()) all = try_compress_time_series_batch(ctx, timestamps, values, unit_off, bounds, &unit_seg_off);
    size_t u = 0;
    for (const UncompressedDataBuffer &b : buffers) {
        CompressedBuffer compressed{b.tag_values, {}};
        for (int index : b.field_column_indices) {
            compressed.compressed_segments.emplace_back(index, all.slice(unit_seg_off[u], unit_seg_off[u + 1]));
            u++;
        }
        out.push_back(std::move(compressed));
    }
    return out;
}

} // namespace modelardb_cuda
