// cabi_on_oracle.cc -- TEST INFRASTRUCTURE: the entry points of include/modelardb_cuda.h that the C++ host layer
// (include/modelardb_cuda.hpp) calls, implemented on the CPU oracle.  Linked INSTEAD of libmodelardb_cuda.so by the CPU
// test of the host layer (tests/test_gpu_cpp_host_api.py::test_cpp_host_layer_logic_on_the_oracle), so that the
// C++ operators' own logic -- leftovers, slicing, tags, limits, grouping, batching -- is exercised on a machine without a
// GPU, exactly like tests/test_operators_host_logic_cpu.py does for the Python mirror.  Never part of the product.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/modelardb_cuda.h"
#include "../../oracle/mdb_oracle.h"

static thread_local std::string g_error;
static int fail(const char *message) {
    g_error = message;
    return MDBCU_FAILURE;
}

struct mdbcu_context {
    int device;
};
struct mdbcu_segments {
    mdbo_segments *owned;
    std::vector<uint64_t> unit_seg_off;
};

static mdbo_segments_view as_oracle(const mdbcu_segments_view *v) {
    mdbo_segments_view o;
    static_assert(sizeof(mdbo_segments_view) == sizeof(mdbcu_segments_view), "the two views have the same layout");
    std::memcpy(&o, v, sizeof(o));
    return o;
}

extern "C" {

const char *mdbcu_last_error(void) { return g_error.c_str(); }

int mdbcu_context_create(int device, mdbcu_context **out) {
    *out = new mdbcu_context{device};
    return MDBCU_SUCCESS;
}
void mdbcu_context_destroy(mdbcu_context *ctx) { delete ctx; }

int mdbcu_compress(mdbcu_context *, mdbcu_space space, const int64_t *timestamps, const float *values, const uint64_t *unit_off,
                   uint64_t n_units, const uint8_t *eb_kind, const float *eb_value, mdbcu_segments **out) {
    if (space != MDBCU_HOST) return fail("the oracle shim only takes host memory");
    for (uint64_t u = 0; u < n_units; u++) {
        if (eb_kind[u] > MDBCU_RELATIVE) return fail("compress: unknown error bound kind");
        if (eb_kind[u] == MDBCU_ABSOLUTE && !(eb_value[u] > 0.0f)) return fail("compress: invalid absolute error bound");
        if (eb_kind[u] == MDBCU_RELATIVE && !(eb_value[u] > 0.0f && eb_value[u] <= 100.0f)) return fail("compress: invalid relative error bound");
        if (unit_off[u + 1] < unit_off[u]) return fail("compress: unit_off must not decrease");
    }
    mdbcu_segments *s = new mdbcu_segments;
    s->unit_seg_off.assign(n_units + 1, 0);
    s->owned = mdbo_compress(timestamps, values, unit_off, n_units, eb_kind, eb_value, 1, s->unit_seg_off.data());
    *out = s;
    return MDBCU_SUCCESS;
}

uint64_t mdbcu_segments_len(const mdbcu_segments *segments) {
    mdbo_segments_view v;
    mdbo_segments_view_get(segments->owned, &v);
    return v.n_segments;
}

int mdbcu_segments_get(mdbcu_segments *segments, mdbcu_space space, mdbcu_segments_view *view, const uint64_t **unit_seg_off) {
    if (space != MDBCU_HOST) return fail("the oracle shim only takes host memory");
    mdbo_segments_view v;
    mdbo_segments_view_get(segments->owned, &v);
    std::memcpy(view, &v, sizeof(v));
    if (unit_seg_off) *unit_seg_off = segments->unit_seg_off.data();
    return MDBCU_SUCCESS;
}

void mdbcu_segments_free(mdbcu_segments *segments) {
    if (!segments) return;
    mdbo_segments_free(segments->owned);
    delete segments;
}

int mdbcu_grid_count(mdbcu_context *, mdbcu_space, const mdbcu_segments_view *segments, uint64_t *point_off, uint64_t *total) {
    const mdbo_segments_view v = as_oracle(segments);
    const uint64_t n = mdbo_grid_count(&v, point_off, 1);
    if (n == ~0ull) return fail("grid_count: malformed segment row");
    if (total) *total = n;
    return MDBCU_SUCCESS;
}

int mdbcu_grid(mdbcu_context *, mdbcu_space, const mdbcu_segments_view *segments, int64_t *timestamps_out, float *values_out,
               uint64_t capacity, uint64_t *n_points) {
    const mdbo_segments_view v = as_oracle(segments);
    const uint64_t total = mdbo_grid_count(&v, nullptr, 1);
    if (total == ~0ull) return fail("grid: malformed segment row");
    if (n_points) *n_points = total;
    if (total > capacity) return fail("grid: capacity too small");
    mdbo_grid(&v, timestamps_out, values_out, capacity, 1);
    return MDBCU_SUCCESS;
}

// (the oracle reconstructs everything and prunes afterwards, as the reference does: grid_exec.rs:366-387)
int mdbcu_grid_range(mdbcu_context *, mdbcu_space, const mdbcu_segments_view *segments, int64_t t_lo, int64_t t_hi, uint64_t *point_off,
                     int64_t *timestamps_out, float *values_out, uint64_t capacity, uint64_t *n_points) {
    const mdbo_segments_view v = as_oracle(segments);
    std::vector<uint64_t> off(segments->n_segments + 1, 0);
    const uint64_t total = mdbo_grid_count(&v, off.data(), 1);
    if (total == ~0ull) return fail("grid_range: malformed segment row");
    std::vector<int64_t> ts(total);
    std::vector<float> val(total);
    mdbo_grid(&v, ts.data(), val.data(), total, 1);
    uint64_t n = 0;
    for (uint64_t i = 0; i < total; i++) n += ts[i] >= t_lo && ts[i] <= t_hi;
    if (n_points) *n_points = n;
    if (point_off) {
        point_off[0] = 0;
        for (uint64_t r = 0; r < segments->n_segments; r++) {
            uint64_t c = 0;
            for (uint64_t i = off[r]; i < off[r + 1]; i++) c += ts[i] >= t_lo && ts[i] <= t_hi;
            point_off[r + 1] = point_off[r] + c;
        }
    }
    if (!timestamps_out && !values_out) return MDBCU_SUCCESS;
    if (n > capacity) return fail("grid_range: capacity too small");
    uint64_t k = 0;
    for (uint64_t i = 0; i < total; i++)
        if (ts[i] >= t_lo && ts[i] <= t_hi) {
            timestamps_out[k] = ts[i];
            values_out[k++] = val[i];
        }
    return MDBCU_SUCCESS;
}

int mdbcu_segment_sums(mdbcu_context *, mdbcu_space, const mdbcu_segments_view *segments, float *sums_out) {
    const mdbo_segments_view v = as_oracle(segments);
    mdbo_segment_sums(&v, sums_out, 1);
    return MDBCU_SUCCESS;
}

int mdbcu_aggregate(mdbcu_context *, mdbcu_space, const mdbcu_segments_view *segments, const uint64_t *group_off, uint64_t n_groups,
                    int64_t *count, float *min, float *max, double *sum) {
    const mdbo_segments_view v = as_oracle(segments);
    mdbo_aggregate(&v, group_off, group_off ? n_groups : 1, count, min, max, sum, 1);
    return MDBCU_SUCCESS;
}

} // extern "C"
