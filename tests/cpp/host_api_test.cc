// host_api_test.cc -- the C++ host layer (include/modelardb_cuda.hpp) against the CPU oracle (oracle/mdb_oracle.h) on
// the same inputs.  Test infrastructure: built and run by tests/test_gpu_cpp_host_api.py on a machine with a GPU.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/modelardb_cuda.hpp"
#include "../../oracle/mdb_oracle.h"

namespace mc = modelardb_cuda;

static int g_checks = 0, g_failures = 0;
#define CHECK(cond, what)                                                             \
    do {                                                                              \
        g_checks++;                                                                   \
        if (!(cond)) {                                                                \
            g_failures++;                                                             \
            std::printf("FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, what);  \
        }                                                                             \
    } while (0)

template <typename T> static bool same_bytes(const std::vector<T> &a, const T *b, size_t n) {
    return a.size() == n && (n == 0 || std::memcmp(a.data(), b, n * sizeof(T)) == 0);
}

static bool equals_oracle(const mc::CompressedSegmentBatch &got, const mdbo_segments_view &want) {
    const size_t s = want.n_segments;
    return got.num_rows() == s && same_bytes(got.model_type_ids, want.model_type_id, s) && same_bytes(got.start_times, want.start_time, s) &&
           same_bytes(got.end_times, want.end_time, s) && same_bytes(got.min_values, want.min_value, s) &&
           same_bytes(got.max_values, want.max_value, s) && same_bytes(got.timestamps_off, want.timestamps_off, s + 1) &&
           same_bytes(got.values_off, want.values_off, s + 1) && same_bytes(got.residuals_off, want.residuals_off, s + 1) &&
           same_bytes(got.timestamps, want.timestamps_data, want.timestamps_off[s]) && same_bytes(got.values, want.values_data, want.values_off[s]) &&
           same_bytes(got.residuals, want.residuals_data, want.residuals_off[s]);
}

int main() {
    std::mt19937 rng(7);
    std::normal_distribution<double> noise(0.0, 0.1);
    std::uniform_real_distribution<float> uniform(-1e3f, 1e3f);
    const size_t n = 20000;
    std::vector<int64_t> ts(3 * n);
    std::vector<float> values(3 * n);
    for (size_t i = 0; i < n; i++) {
        ts[i] = ts[n + i] = ts[2 * n + i] = 1600000000000000LL + 1000LL * (int64_t)i;
        values[i] = (float)(100.0 + 10.0 * std::sin(2.0 * M_PI * (double)i / 1000.0) + noise(rng)); // models
        values[n + i] = uniform(rng);                                                                // one MacaqueV row
        values[2 * n + i] = 42.5f;                                                                   // one PMC-Mean model
    }
    const std::vector<uint64_t> unit_off = {0, n, 2 * n, 3 * n};
    const std::vector<mc::ErrorBound> bounds = {mc::ErrorBound::try_new_relative(1.0f), mc::ErrorBound::lossless(),
                                                mc::ErrorBound::try_new_absolute(0.5f)};
    const uint8_t kinds[3] = {MDBCU_RELATIVE, MDBCU_LOSSLESS, MDBCU_ABSOLUTE};
    const float eb_values[3] = {1.0f, 0.0f, 0.5f};

    mc::Context ctx(0);

    // ---- the batch form against the oracle
    std::vector<uint64_t> unit_seg_off, want_uso(4);
    const mc::CompressedSegmentBatch batch = mc::try_compress_time_series_batch(ctx, ts, values, unit_off, bounds, &unit_seg_off);
    mdbo_segments *oracle_segments = mdbo_compress(ts.data(), values.data(), unit_off.data(), 3, kinds, eb_values, 4, want_uso.data());
    mdbo_segments_view want;
    mdbo_segments_view_get(oracle_segments, &want);
    CHECK(equals_oracle(batch, want), "batch compress: every column bit-identical");
    CHECK(unit_seg_off == want_uso, "rows of each unit");

    // ---- one series at a time (compression.rs:191-275)
    for (int u = 0; u < 3; u++) {
        const std::vector<int64_t> uts(ts.begin() + u * n, ts.begin() + (u + 1) * n);
        const std::vector<float> uval(values.begin() + u * n, values.begin() + (u + 1) * n);
        const mc::CompressedSegmentBatch one = mc::try_compress_univariate_time_series(ctx, uts, uval, bounds[u]);
        const mc::CompressedSegmentBatch expected = batch.slice(unit_seg_off[u], unit_seg_off[u + 1]);
        CHECK(one.num_rows() == expected.num_rows() && one.values == expected.values && one.timestamps == expected.timestamps &&
                  one.residuals == expected.residuals && one.start_times == expected.start_times && one.model_type_ids == expected.model_type_ids,
              "univariate compress equals its rows of the batch");
    }

    // ---- len, grid, sum (models/mod.rs:98-251)
    const std::vector<uint64_t> point_off = mc::len(ctx, batch);
    std::vector<uint64_t> want_point_off(want.n_segments + 1);
    const uint64_t total = mdbo_grid_count(&want, want_point_off.data(), 4);
    CHECK(point_off == want_point_off && total == 3 * n, "len of every row");
    std::vector<int64_t> grid_ts;
    std::vector<float> grid_val;
    mc::grid(ctx, batch, grid_ts, grid_val);
    std::vector<int64_t> want_ts(total);
    std::vector<float> want_val(total);
    mdbo_grid(&want, want_ts.data(), want_val.data(), total, 4);
    CHECK(grid_ts == want_ts && grid_ts == ts, "grid timestamps");
    CHECK(same_bytes(grid_val, want_val.data(), total), "grid values bit-identical");
    { // the time predicate inside the call (mdbcu_grid_range) against grid-then-filter (grid_exec.rs:366-387)
        const int64_t t_lo = ts[n / 3], t_hi = ts[n + n / 2]; // from inside the first unit to the middle of the second (timestamps repeat per unit)
        std::vector<int64_t> r_ts, f_ts;
        std::vector<float> r_val, f_val;
        std::vector<uint64_t> r_off;
        mc::grid_range(ctx, batch, t_lo, t_hi, r_ts, r_val, &r_off);
        for (uint64_t i = 0; i < total; i++)
            if (want_ts[i] >= t_lo && want_ts[i] <= t_hi) {
                f_ts.push_back(want_ts[i]);
                f_val.push_back(want_val[i]);
            }
        CHECK(r_ts == f_ts && !f_ts.empty() && f_ts.size() < total, "grid_range timestamps");
        CHECK(same_bytes(r_val, f_val.data(), f_val.size()), "grid_range values bit-identical");
        CHECK(r_off.size() == batch.num_rows() + 1 && r_off.back() == f_ts.size(), "grid_range point offsets");
    }
    const std::vector<float> sums = mc::sum(ctx, batch);
    std::vector<float> want_sums(want.n_segments);
    mdbo_segment_sums(&want, want_sums.data(), 4);
    CHECK(same_bytes(sums, want_sums.data(), want.n_segments), "sum of every row");

    // ---- GridStream (grid_exec.rs:197-430): two segment batches with one tag column, batch_size 4096
    {
        std::vector<mc::GridStream::Input> input;
        size_t cuts[3] = {0, batch.num_rows() / 2, batch.num_rows()};
        for (int b = 0; b < 2; b++) {
            mc::CompressedSegmentBatch part = batch.slice(cuts[b], cuts[b + 1]);
            std::vector<std::string> tag(part.num_rows(), b == 0 ? "first" : "second");
            input.push_back({std::move(part), {std::move(tag)}});
        }
        mc::GridStream stream(ctx, std::move(input), 4096, 1);
        std::vector<int64_t> all_ts;
        std::vector<float> all_val;
        size_t first_rows = 0, batches = 0;
        mc::GridStream::Batch out;
        bool sizes_ok = true;
        while (stream.poll_next(out)) {
            batches++;
            sizes_ok = sizes_ok && out.timestamps.size() <= 4096 && out.tags[0].size() == out.timestamps.size();
            for (const std::string &t : out.tags[0]) first_rows += t == "first";
            all_ts.insert(all_ts.end(), out.timestamps.begin(), out.timestamps.end());
            all_val.insert(all_val.end(), out.values.begin(), out.values.end());
        }
        CHECK(sizes_ok && batches >= total / 4096, "batches of at most batch_size rows");
        CHECK(all_ts == want_ts && same_bytes(all_val, want_val.data(), total), "the stream is the concatenated grid");
        CHECK(first_rows == point_off[cuts[1]], "tags repeated once per created row");
    }

#ifdef HOST_API_EXTENDED // (see tests/test_gpu_cpp_host_api.py for which run defines it)
    // ---- GridStream options: predicate after reconstruction, time range and limit pushed down to whole segments
    {
        const std::vector<std::string> series_of_row = [&] {
            std::vector<std::string> t;
            for (int u = 0; u < 3; u++) t.insert(t.end(), unit_seg_off[u + 1] - unit_seg_off[u], "series-" + std::to_string(u));
            return t;
        }();
        auto run = [&](mc::GridStream::Options options, std::vector<int64_t> &all_ts, std::vector<float> &all_val, std::vector<std::string> &all_tag,
                       size_t &skipped, uint64_t &created) {
            options.n_tag_columns = 1;
            std::vector<mc::GridStream::Input> input;
            const size_t cut = batch.num_rows() / 3;
            input.push_back({batch.slice(0, cut), {std::vector<std::string>(series_of_row.begin(), series_of_row.begin() + cut)}});
            input.push_back({batch.slice(cut, batch.num_rows()), {std::vector<std::string>(series_of_row.begin() + cut, series_of_row.end())}});
            mc::GridStream stream(ctx, std::move(input), options);
            mc::GridStream::Batch out;
            bool sizes_ok = true;
            while (stream.poll_next(out)) {
                sizes_ok = sizes_ok && out.timestamps.size() <= stream.batch_size() && out.tags[0].size() == out.timestamps.size();
                all_ts.insert(all_ts.end(), out.timestamps.begin(), out.timestamps.end());
                all_val.insert(all_val.end(), out.values.begin(), out.values.end());
                all_tag.insert(all_tag.end(), out.tags[0].begin(), out.tags[0].end());
            }
            skipped = stream.segments_skipped();
            created = stream.rows_created();
            return sizes_ok;
        };
        std::vector<std::string> want_tag;
        for (size_t row = 0; row < batch.num_rows(); row++) want_tag.insert(want_tag.end(), point_off[row + 1] - point_off[row], series_of_row[row]);
        // (all three series share their timestamps: a time window selects the same stretch of each)
        const int64_t lo = ts[n / 3], hi = ts[n / 2];
        mc::GridStream::Options pruned_options, pushed_options;
        pruned_options.batch_size = pushed_options.batch_size = 5000;
        pruned_options.predicate = pushed_options.predicate = [&](const std::vector<int64_t> &t, const std::vector<float> &) {
            std::vector<uint8_t> keep(t.size());
            for (size_t i = 0; i < t.size(); i++) keep[i] = t[i] >= lo && t[i] <= hi;
            return keep;
        };
        pushed_options.time_range_start = lo;
        pushed_options.time_range_end = hi;
        std::vector<int64_t> a_ts, b_ts, expect_ts;
        std::vector<float> a_val, b_val, expect_val;
        std::vector<std::string> a_tag, b_tag, expect_tag;
        size_t a_skipped, b_skipped;
        uint64_t a_created, b_created;
        const bool a_ok = run(pruned_options, a_ts, a_val, a_tag, a_skipped, a_created);
        const bool b_ok = run(pushed_options, b_ts, b_val, b_tag, b_skipped, b_created);
        for (size_t i = 0; i < want_ts.size(); i++)
            if (want_ts[i] >= lo && want_ts[i] <= hi) {
                expect_ts.push_back(want_ts[i]);
                expect_val.push_back(want_val[i]);
                expect_tag.push_back(want_tag[i]);
            }
        CHECK(a_ok && b_ok, "batch sizes with a predicate");
        CHECK(a_ts == expect_ts && same_bytes(a_val, expect_val.data(), expect_val.size()) && a_tag == expect_tag, "predicate after reconstruction");
        CHECK(b_ts == expect_ts && same_bytes(b_val, expect_val.data(), expect_val.size()) && b_tag == expect_tag, "time range pushed down: same rows");
        CHECK(a_skipped == 0 && a_created == total && b_skipped > 0 && b_created < total, "segments outside the range are never reconstructed");

        // the same range evaluated per point inside the reconstruction call: only the rows of the answer are ever created
        for (int with_predicate = 0; with_predicate < 2; with_predicate++) {
            mc::GridStream::Options clipped_options = pushed_options;
            clipped_options.device_time_clip = true;
            if (!with_predicate) clipped_options.predicate = nullptr;
            std::vector<int64_t> c_ts;
            std::vector<float> c_val;
            std::vector<std::string> c_tag;
            size_t c_skipped;
            uint64_t c_created;
            const bool c_ok = run(clipped_options, c_ts, c_val, c_tag, c_skipped, c_created);
            CHECK(c_ok && c_ts == expect_ts && same_bytes(c_val, expect_val.data(), expect_val.size()) && c_tag == expect_tag, "time range clipped in the grid call: same rows");
            CHECK(c_skipped == b_skipped && c_created == expect_ts.size(), "clipped: no row outside the range is created");
        }
        {   // a half-open range clips one side only
            mc::GridStream::Options half;
            half.batch_size = 5000;
            half.time_range_start = lo;
            half.device_time_clip = true;
            std::vector<int64_t> h_ts;
            std::vector<float> h_val;
            std::vector<std::string> h_tag;
            size_t h_skipped;
            uint64_t h_created;
            const bool h_ok = run(half, h_ts, h_val, h_tag, h_skipped, h_created);
            size_t expect_rows = 0;
            bool same = true;
            for (size_t i = 0; i < want_ts.size(); i++)
                if (want_ts[i] >= lo) {
                    same = same && expect_rows < h_ts.size() && h_ts[expect_rows] == want_ts[i] && h_tag[expect_rows] == want_tag[i] &&
                           std::memcmp(&h_val[expect_rows], &want_val[i], 4) == 0;
                    expect_rows++;
                }
            CHECK(h_ok && same && h_ts.size() == expect_rows && h_created == expect_rows, "clipped with a start only");
        }

        for (size_t limit : {size_t(1), size_t(4097), size_t(30000)}) {
            mc::GridStream::Options options;
            options.batch_size = 4096;
            options.limit = limit;
            std::vector<int64_t> l_ts;
            std::vector<float> l_val;
            std::vector<std::string> l_tag;
            size_t l_skipped;
            uint64_t l_created;
            const bool l_ok = run(options, l_ts, l_val, l_tag, l_skipped, l_created);
            CHECK(l_ok && l_ts.size() == limit && std::equal(l_ts.begin(), l_ts.end(), want_ts.begin()) && same_bytes(l_val, want_val.data(), limit) &&
                      std::equal(l_tag.begin(), l_tag.end(), want_tag.begin()),
                  "LIMIT: the first rows of the stream");
            CHECK(l_skipped > 0 && l_created < total, "LIMIT: later segments are not reconstructed");
        }
        bool threw = false;
        try {
            mc::GridStream::Options options;
            options.limit = 0;
            mc::GridStream stream(ctx, {}, options);
        } catch (const mc::Error &) { threw = true; }
        CHECK(threw, "a limit of zero is rejected");
    }

    // ---- GROUP BY tag straight from segments
    {
        std::vector<std::string> tag;
        for (int u = 0; u < 3; u++) tag.insert(tag.end(), unit_seg_off[u + 1] - unit_seg_off[u], u == 2 ? "a" : (u == 0 ? "a" : "b")); // units 0 and 2 share a key
        const mc::GroupedAggregates g = mc::grouped_model_aggregates(ctx, batch, {tag});
        std::vector<int64_t> unit_count(3);
        std::vector<float> unit_min(3), unit_max(3);
        std::vector<double> unit_sum(3);
        mdbo_aggregate(&want, want_uso.data(), 3, unit_count.data(), unit_min.data(), unit_max.data(), unit_sum.data(), 1);
        CHECK(g.keys.size() == 2 && g.keys[0] == std::vector<std::string>{"a"} && g.keys[1] == std::vector<std::string>{"b"}, "keys in order of first appearance");
        CHECK(g.count[0] == unit_count[0] + unit_count[2] && g.count[1] == unit_count[1], "grouped COUNT");
        CHECK(g.min[0] == std::min(unit_min[0], unit_min[2]) && g.max[0] == std::max(unit_max[0], unit_max[2]) && g.min[1] == unit_min[1] && g.max[1] == unit_max[1],
              "grouped MIN / MAX");
        CHECK(std::fabs(g.sum[0] - (unit_sum[0] + unit_sum[2])) <= 1e-12 * std::fabs(unit_sum[0] + unit_sum[2]) && std::fabs(g.sum[1] - unit_sum[1]) <= 1e-12 * std::fabs(unit_sum[1]),
              "grouped SUM");
        bool threw = false;
        try { mc::grouped_model_aggregates(ctx, batch, {std::vector<std::string>(3, "x")}); } catch (const mc::Error &) { threw = true; }
        CHECK(threw, "a tag column needs one value per segment");
    }

    // ---- the server compressor's finished buffers, many per call (uncompressed_data_manager.rs:530-581)
    {
        std::vector<mc::UncompressedDataBuffer> buffers;
        for (int b = 0; b < 4; b++) {
            const size_t m = 500 + 1500 * b;
            mc::UncompressedDataBuffer buffer;
            buffer.timestamps.assign(ts.begin(), ts.begin() + m);
            buffer.field_columns = {std::vector<float>(values.begin() + b * 100, values.begin() + b * 100 + m),
                                    std::vector<float>(values.begin() + n, values.begin() + n + m)};
            buffer.field_column_indices = {1, 3};
            buffer.error_bounds = {mc::ErrorBound::try_new_relative(1.0f), mc::ErrorBound::lossless()};
            buffer.tag_values = {"buffer-" + std::to_string(b)};
            buffers.push_back(std::move(buffer));
        }
        const std::vector<mc::CompressedBuffer> compressed = mc::compress_finished_buffers(ctx, buffers);
        bool all_equal = compressed.size() == buffers.size();
        for (size_t b = 0; all_equal && b < buffers.size(); b++) {
            all_equal = compressed[b].tag_values == buffers[b].tag_values && compressed[b].compressed_segments.size() == 2 &&
                        compressed[b].compressed_segments[0].first == 1 && compressed[b].compressed_segments[1].first == 3;
            for (size_t f = 0; all_equal && f < 2; f++) { // each (buffer, field) pair is what one call of the reference produces
                const uint64_t one_off[2] = {0, buffers[b].timestamps.size()};
                const uint8_t kind = buffers[b].error_bounds[f].kind();
                const float value = buffers[b].error_bounds[f].value();
                mdbo_segments *alone = mdbo_compress(buffers[b].timestamps.data(), buffers[b].field_columns[f].data(), one_off, 1, &kind, &value, 1, nullptr);
                mdbo_segments_view alone_view;
                mdbo_segments_view_get(alone, &alone_view);
                all_equal = equals_oracle(compressed[b].compressed_segments[f].second, alone_view);
                mdbo_segments_free(alone);
            }
        }
        CHECK(all_equal, "many finished buffers in one call: one segment batch per buffer and field, bit-identical");
        CHECK(mc::compress_finished_buffers(ctx, {}).empty(), "no buffers, no call");
        bool threw = false;
        buffers[1].field_columns[0].pop_back();
        try { mc::compress_finished_buffers(ctx, buffers); } catch (const mc::Error &) { threw = true; }
        CHECK(threw, "different lengths are the reference's error");
    }

#endif // HOST_API_EXTENDED

    // ---- accumulators (model_simple_aggregates.rs:336-618)
    {
        int64_t want_count;
        float want_min, want_max;
        double want_sum;
        mdbo_aggregate(&want, nullptr, 1, &want_count, &want_min, &want_max, &want_sum, 1);
        mc::ModelCountAccumulator count(ctx);
        mc::ModelMinAccumulator min(ctx);
        mc::ModelMaxAccumulator max(ctx);
        mc::ModelSumAccumulator sum(ctx);
        mc::ModelAvgAccumulator avg(ctx);
        for (size_t lo = 0; lo < batch.num_rows(); lo += 100) { // fed in several batches, like DataFusion does
            const mc::CompressedSegmentBatch part = batch.slice(lo, std::min(batch.num_rows(), lo + 100));
            count.update_batch(part); min.update_batch(part); max.update_batch(part); sum.update_batch(part); avg.update_batch(part);
        }
        const double got_sum = sum.state();
        const auto avg_state = avg.state();
        CHECK(count.state() == want_count && (int64_t)avg_state.first == want_count, "COUNT");
        CHECK(min.state() == want_min && max.state() == want_max, "MIN / MAX");
        CHECK(std::fabs(got_sum - want_sum) <= 1e-12 * std::fabs(want_sum) && std::fabs(avg_state.second - want_sum) <= 1e-12 * std::fabs(want_sum), "SUM");
        CHECK(count.state() == 0 && sum.state() == 0.0, "state() resets");
    }

    // ---- error behaviour
    {
        bool threw = false;
        try { mc::try_compress_univariate_time_series(ctx, std::vector<int64_t>(3), std::vector<float>(2), mc::ErrorBound::lossless()); } catch (const mc::Error &) { threw = true; }
        CHECK(threw, "different lengths are an InvalidArgument (compression.rs:202-206)");
        threw = false;
        try { mc::ErrorBound::try_new_relative(101.0f); } catch (const mc::Error &) { threw = true; }
        CHECK(threw, "relative bounds above 100 % are rejected (types.rs:325-334)");
        threw = false;
        try { mc::ErrorBound::try_new_absolute(-1.0f); } catch (const mc::Error &) { threw = true; }
        CHECK(threw, "absolute bounds must be positive (types.rs:312-321)");
        const mc::CompressedSegmentBatch empty = mc::try_compress_univariate_time_series(ctx, {}, {}, mc::ErrorBound::lossless());
        CHECK(empty.num_rows() == 0, "empty input gives an empty batch (compression.rs:208-211)");
    }

    mdbo_segments_free(oracle_segments);
    std::printf("host_api_test: %d checks, %d failures\n", g_checks, g_failures);
    return g_failures ? 1 : 0;
}
