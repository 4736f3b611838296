// emu.cc -- DEBUGGING HARNESS, not a product path and not a fallback.
//
// The build container has no GPU, and a gpurun round trip takes minutes.  This file compiles the
// per-thread kernel bodies of modelardb_rs_b200/csrc/mdb_{grid,aggregate,compress}.cuh for the host
// (MDB_DEV -> inline, intrinsics -> mdb_host_shim.h) and steps them in plain loops, so that logic
// errors in those bodies are caught by `pytest -m "not gpu"` before GPU time is spent.  What it cannot
// cover -- launch geometry, scans, shared-memory tiling, warp primitives, the C-ABI -- is covered by
// the `-m gpu` tests.  Nothing under modelardb_rs_b200/ links or loads this.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../modelardb_rs_b200/csrc/mdb_aggregate.cuh"
#include "../../modelardb_rs_b200/csrc/mdb_compress.cuh"
#include "../../modelardb_rs_b200/csrc/mdb_grid.cuh"
#include "../../modelardb_rs_b200/csrc/mdb_swing_sums.cuh"

// The warp-cooperative fit (mdb_fit_warp.cuh) on the host: 32 fibers per warp, see warp_emu.h.
#define MDB_WARP_EMU
#include "warp_emu.h"
#include "../../modelardb_rs_b200/csrc/mdb_fit_warp.cuh"
#include "../../modelardb_rs_b200/csrc/mdb_fit_lanes.cuh"
#include "../../modelardb_rs_b200/csrc/mdb_fit_screen.cuh"
#include "../../modelardb_rs_b200/csrc/mdb_macaque_warp.cuh"

using namespace mdb;

struct EmuSegments {
    std::vector<int8_t> model_type_id;
    std::vector<int64_t> start_time, end_time;
    std::vector<float> min_value, max_value;
    std::vector<uint64_t> ts_off, val_off, res_off, unit_seg_off;
    std::vector<uint8_t> ts_data, val_data, res_data;
};

// One chain of one chunk with the chosen engine: 1 = the one-thread fit, 2 = the warp-cooperative fit (every lane runs
// spec_chain on its own copy of the chunk state, as the lanes of k_spec_chain_warp / k_spec_async do; lane 0's copy is kept).
using WarpFitScreen = WarpFitScreenT<MDB_FIT_POINTS_PER_LANE>;
// k_lanes_units + k_lanes_regular for one unit: what the screened engine (mdb_fit_screen.cuh) is told about it
static LaneUnit emu_lane_unit(const ErrorBound &eb, const int64_t *uts, uint32_t n) {
    LaneUnit lu = lane_unit_init(uts, n, eb);
    if (lu.ok)
        for (uint32_t i = 0; i < n; i++)
            if (uts[i] != uts[0] + (int64_t)i * (uts[1] - uts[0])) { lu.irregular = 1; break; }
    return lu;
}
static void emu_run_chain(int engine, const ErrorBound &eb, const int64_t *uts, const float *uval, uint32_t n, uint32_t chunk_end,
                          uint32_t budget, ChunkState &st, FittedModel *lists, uint32_t cap) {
    if (engine == 5) {
        std::vector<double> smem(WarpFitScreen::SMEM_DOUBLES);
        const LaneUnit lu = emu_lane_unit(eb, uts, n);
        const ChunkState before = st;
        ChunkState after = st;
        warp_emu::run([&](int lane) {
            ChunkState mine = before;
            WarpFitScreen fitter(eb, uts, uval, n, smem.data(), &lu);
            spec_chain(fitter, (uint32_t)lane, 32u, n, chunk_end, budget, mine, lists, cap);
            if (lane == 0) after = mine;
        });
        st = after;
        return;
    }
    if (engine != 2) {
        ScalarFit fitter(eb, uts, uval, n);
        spec_chain(fitter, 0u, 1u, n, chunk_end, budget, st, lists, cap);
        return;
    }
    std::vector<double> smem(WarpFit::SMEM_DOUBLES);
    const ChunkState before = st;
    ChunkState after = st;
    warp_emu::run([&](int lane) {
        ChunkState mine = before;
        WarpFit fitter(eb, uts, uval, n, smem.data());
        spec_chain(fitter, (uint32_t)lane, 32u, n, chunk_end, budget, mine, lists, cap);
        if (lane == 0) after = mine;
    });
    st = after;
}
static int g_emu_engine = 1;
static int g_emu_lanes = 0; // 1: one lane per chunk first (mdb_fit_lanes.cuh: k_lanes_units, k_lanes_regular, k_spec_lanes), then the stitching
static uint64_t g_emu_lane_chunks = 0, g_emu_lane_bailed = 0;
static uint32_t g_emu_lane_warmup = 0, g_emu_lane_rounds = 4;
static uint64_t g_emu_lane_reruns = 0;

// k_lanes_units + k_lanes_regular + k_spec_lanes for one unit: every chunk's chain from the chunk's first index, one point
// per step, values read straight from the array (the kernel reads the same values through its ring).
template <int KIND>
static void emu_unit_lanes_kind(const LaneUnit &lu, const float *uval, uint32_t n, uint32_t L, uint32_t C, uint32_t cap, std::vector<ChunkState> &st,
                                std::vector<FittedModel> &lists) {
    // one chunk's chain, as a lane of k_spec_lanes runs it (rerun: a round, from the entry spec_propagate_unit gave the chunk)
    auto run = [&](uint32_t c, bool rerun) {
        const uint32_t lo = c * L, chunk_end = (uint32_t)std::min<uint64_t>((uint64_t)lo + L, n), limit = (uint32_t)std::min<uint64_t>((uint64_t)chunk_end + L, n);
        const ChunkState s0 = st[c];
        if (rerun && (!s0.dirty || s0.entry == IDX_NONE || (s0.new_entry == s0.entry && s0.exit == IDX_NONE))) return;
        LaneChain chain;
        FittedModel *list = lists.data() + ((size_t)c * 2 + (s0.buf ^ 1)) * cap;
        bool done = false;
        if (rerun) {
            chain.begin(s0.new_entry, lo, chunk_end, limit, n);
            chain.set_old_chain(lists.data() + ((size_t)c * 2 + s0.buf) * cap, s0.n_models, s0.entry, s0.exit, s0.truncated_at);
            done = chain.splice_at(s0.new_entry, list);
        } else {
            chain.begin(lo > g_emu_lane_warmup ? lo - g_emu_lane_warmup : 0u, lo, chunk_end, limit, n);
        }
        while (!done) done = chain.template step<KIND>(lu, uval[chain.fit.idx], list);
        g_emu_lane_chunks++;
        if (chain.bailed) {
            g_emu_lane_bailed++;
            return;
        }
        lane_chain_publish(chain, st[c]);
        st[c].phase = PH_DONE;
    };
    for (uint32_t c = 0; c < C; c++) run(c, false);
    uint32_t resume_c = 0, resume_e = 0;
    size_t last = C;
    for (uint32_t round = 0; round < g_emu_lane_rounds; round++) { // k_spec_propagate + k_spec_lanes over its worklist
        std::vector<uint32_t> work;
        spec_propagate_unit(n, L, C, st.data(), true, resume_c, resume_e, [&](uint32_t c) { work.push_back(c); });
        if (work.empty() || (round >= 1 && work.size() * 10 > last * 9)) break;
        for (uint32_t c : work) run(c, true);
        g_emu_lane_reruns += work.size();
        last = work.size();
    }
}
static void emu_unit_lanes(const ErrorBound &eb, const int64_t *uts, const float *uval, uint32_t n, uint32_t L, uint32_t C, uint32_t cap,
                           std::vector<ChunkState> &st, std::vector<FittedModel> &lists) {
    LaneUnit lu = lane_unit_init(uts, n, eb);
    if (!lu.ok) return;
    for (uint32_t i = 0; i < n; i++)
        if (uts[i] != uts[0] + (int64_t)i * (uts[1] - uts[0])) return; // k_lanes_regular
    if (eb.kind == KIND_LOSSLESS) emu_unit_lanes_kind<KIND_LOSSLESS>(lu, uval, n, L, C, cap, st, lists);
    else if (eb.kind == KIND_ABSOLUTE) emu_unit_lanes_kind<KIND_ABSOLUTE>(lu, uval, n, L, C, cap, st, lists);
    else emu_unit_lanes_kind<KIND_RELATIVE>(lu, uval, n, L, C, cap, st, lists);
}

// The asynchronous scheduler (k_spec_async + sched_advance) for one unit, single-threaded: up to `in_flight`
// "workers" hold a claimed chunk at a time, and a seeded generator decides whether the next event is a worker
// claiming the next queue item or one of the claimed chains completing (publish + sched_advance).  Returns
// false if the schedule stalls or the queue overflows.
static bool emu_unit_async(const ErrorBound &eb, const int64_t *uts, const float *uval, uint32_t n, uint32_t L, uint32_t C, uint32_t cap,
                           std::vector<ChunkState> &st, std::vector<FittedModel> &lists, uint32_t seed, uint32_t in_flight, uint32_t *runs_out) {
    uint64_t rng = 0x9E3779B97F4A7C15ull * (seed + 1);
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    std::vector<uint32_t> queue;
    for (uint32_t c = 0; c < C; c++)
        if (st[c].phase == PH_QUEUED) queue.push_back(c); // k_sched_count / k_sched_fill: chunks no lane has run
    size_t head = 0;
    UnitSched us;
    us.lock = 0; us.next_c = 0; us.entry = 0; us.finished = 0;
    if (g_emu_lanes) sched_advance(us, n, L, C, st.data(), [&](uint32_t cc) { queue.push_back(cc); }); // k_sched_kick
    struct Claimed { uint32_t c; ChunkState s; };
    std::vector<Claimed> claimed;
    uint32_t runs = 0;
    while (!us.finished) {
        const bool can_claim = head < queue.size() && claimed.size() < in_flight;
        if (!can_claim && claimed.empty()) return false; // stalled: nothing queued, nothing running, unit not final
        if (can_claim && (claimed.empty() || (next() & 1))) {
            uint32_t c = queue[head++];
            if (sync_cas(&st[c].phase, PH_QUEUED, PH_RUNNING) != PH_QUEUED) return false; // one queue entry per PH_QUEUED state
            if (c < us.next_c) { st[c].phase = PH_DONE; continue; } // the frontier is past it
            Claimed w;
            w.c = c;
            w.s = st[c];
            uint32_t ce = std::min<uint64_t>((uint64_t)(c + 1) * L, n);
            emu_run_chain(g_emu_engine, eb, uts, uval, n, ce, L, w.s, lists.data() + (size_t)c * 2 * cap, cap); // (new list is unread until published)
            runs++;
            claimed.push_back(w);
        } else {
            size_t k = next() % claimed.size();
            Claimed w = claimed[k];
            claimed.erase(claimed.begin() + k);
            w.s.phase = PH_RUNNING;
            st[w.c] = w.s;
            st[w.c].phase = PH_DONE;
            sched_advance(us, n, L, C, st.data(), [&](uint32_t cc) { queue.push_back(cc); });
            if (queue.size() > 3 * (size_t)C + 8) return false;
        }
    }
    if (runs_out) *runs_out = runs;
    return true;
}

extern "C" {

// chunk_len == 0: one chunk per unit (the plain sequential chain).
// sched_seed == 0: rounds (k_spec_chain + k_spec_propagate); otherwise the asynchronous scheduler with that seed and
// `in_flight` concurrent workers; rounds_out then receives the largest number of chain runs of any unit.
EmuSegments *emu_compress_sched(const int64_t *ts, const float *values, const uint64_t *unit_off, uint64_t n_units,
                                const uint8_t *eb_kind, const float *eb_value, uint32_t chunk_len, uint32_t *rounds_out,
                                uint32_t sched_seed, uint32_t in_flight);

EmuSegments *emu_compress(const int64_t *ts, const float *values, const uint64_t *unit_off, uint64_t n_units,
                          const uint8_t *eb_kind, const float *eb_value, uint32_t chunk_len, uint32_t *rounds_out) {
    return emu_compress_sched(ts, values, unit_off, n_units, eb_kind, eb_value, chunk_len, rounds_out, 0, 0);
}

EmuSegments *emu_compress_sched(const int64_t *ts, const float *values, const uint64_t *unit_off, uint64_t n_units,
                                const uint8_t *eb_kind, const float *eb_value, uint32_t chunk_len, uint32_t *rounds_out,
                                uint32_t sched_seed, uint32_t in_flight) {
    EmuSegments *out = new EmuSegments();
    out->unit_seg_off.assign(n_units + 1, 0);
    std::vector<SegRecord> recs;
    std::vector<uint32_t> row_unit;
    uint32_t max_rounds = 0;
    for (uint64_t u = 0; u < n_units; u++) {
        uint32_t n = (uint32_t)(unit_off[u + 1] - unit_off[u]);
        out->unit_seg_off[u + 1] = out->unit_seg_off[u];
        if (n == 0) continue;
        const int64_t *uts = ts + unit_off[u];
        const float *uval = values + unit_off[u];
        ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
        uint32_t L = chunk_len ? chunk_len : n;
        uint32_t C = (n + L - 1) / L;
        uint32_t cap = models_per_chunk(L);
        std::vector<ChunkState> st(C);
        std::vector<FittedModel> lists((size_t)C * 2 * cap);
        for (uint32_t c = 0; c < C; c++) { // k_spec_init
            std::memset(&st[c], 0, sizeof(ChunkState));
            st[c].entry = IDX_NONE; st[c].exit = IDX_NONE; st[c].new_entry = c * L; st[c].dirty = 1; st[c].exact = c == 0;
        }
        uint32_t rounds = 0, resume_c = 0, resume_e = 0;
        if (sched_seed) {
            for (uint32_t c = 0; c < C; c++) st[c].phase = PH_QUEUED;
            if (g_emu_lanes) emu_unit_lanes(eb, uts, uval, n, L, C, cap, st, lists);
            if (!emu_unit_async(eb, uts, uval, n, L, C, cap, st, lists, sched_seed + (uint32_t)u, in_flight ? in_flight : 1, &rounds)) return nullptr;
        }
        while (!sched_seed) { // rounds: k_spec_chain over dirty chunks, then k_spec_propagate per unit
            for (uint32_t c = 0; c < C; c++)
                if (st[c].dirty) {
                    uint32_t cs = c * L, ce = std::min<uint64_t>((uint64_t)(c + 1) * L, n);
                    (void)cs;
                    emu_run_chain(g_emu_engine, eb, uts, uval, n, ce, L, st[c], lists.data() + (size_t)c * 2 * cap, cap);
                }
            rounds++;
            if (spec_propagate_unit(n, L, C, st.data(), rounds == 1, resume_c, resume_e, [](uint32_t) {}) == 0) break;
            if (rounds > 4 * C + 8) return nullptr; // must converge: one chunk becomes final per round at worst
        }
        max_rounds = std::max(max_rounds, rounds);
        uint8_t irregular;
        spec_finalize_unit(n, L, C, st.data(), irregular);
        for (uint32_t c = 0; c < C; c++) {
            FittedModel *list = lists.data() + ((size_t)c * 2 + st[c].buf) * cap;
            if (!st[c].skipped)
                for (uint32_t k = 0; k < st[c].n_models; k++) // k_swing_finish: the warp engine leaves Swing models pending
                    if (list[k].pending) swing_finish(list[k], uts, uval);
            uint32_t rows = spec_count_rows(st[c], list);
            size_t base = recs.size();
            recs.resize(base + rows);
            uint32_t wrote = spec_records(eb, uts, uval, st[c], list, !irregular, recs.data() + base);
            if (wrote != rows) return nullptr;
            row_unit.insert(row_unit.end(), rows, (uint32_t)u);
            out->unit_seg_off[u + 1] += rows;
        }
    }
    if (rounds_out) *rounds_out = max_rounds;
    uint64_t S = recs.size();
    out->model_type_id.resize(S); out->start_time.resize(S); out->end_time.resize(S);
    out->min_value.resize(S); out->max_value.resize(S);
    out->ts_off.assign(S + 1, 0); out->val_off.assign(S + 1, 0); out->res_off.assign(S + 1, 0);
    for (uint64_t r = 0; r < S; r++) { // gather + scans
        const SegRecord &rec = recs[r];
        uint64_t a = unit_off[row_unit[r]];
        out->model_type_id[r] = rec.model_type_id;
        out->start_time[r] = ts[a + rec.start_index];
        out->end_time[r] = ts[a + rec.res_end_index];
        out->min_value[r] = rec.min_value;
        out->max_value[r] = rec.max_value;
        out->ts_off[r + 1] = out->ts_off[r] + rec.ts_len;
        out->val_off[r + 1] = out->val_off[r] + rec.val_len;
        out->res_off[r + 1] = out->res_off[r] + rec.res_len;
    }
    out->ts_data.assign(out->ts_off[S] + 1, 0xAA);
    out->val_data.assign(out->val_off[S] + 1, 0xAA);
    out->res_data.assign(out->res_off[S] + 1, 0xAA);
    for (uint64_t r = 0; r < S; r++) { // pass 2: one "thread" per row
        uint32_t u = row_unit[r];
        ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
        compress_emit_segment(eb, recs[r], ts + unit_off[u], values + unit_off[u], out->ts_data.data() + out->ts_off[r],
                              out->val_data.data() + out->val_off[r], out->res_data.data() + out->res_off[r]);
    }
    return out;
}

// For every start index: does fit_reaches_eight_points agree with "fit_next_model returns a stored model"?
// Returns the number of disagreements (the warp engine's skip_rejected relies on there being none).
uint64_t emu_check_eight_points(const int64_t *ts, const float *values, uint32_t n, uint8_t eb_kind, float eb_value) {
    ErrorBound eb = make_error_bound(eb_kind, eb_value);
    uint64_t bad = 0;
    for (uint32_t s = 0; s < n; s++) {
        RegularityTracker trk;
        trk.init(ts, s, n);
        bool aborted;
        FittedModel m = fit_next_model(eb, ts, values, s, n, trk, n, aborted);
        bool stored = m.bytes_per_value <= 4.0f;
        if (stored != fit_reaches_eight_points(eb, ts, values, s, n)) bad++;
    }
    return bad;
}

// Engine of the chains in emu_compress*: 1 the one-thread fit, 2 the warp-cooperative fit on 32 fibers.
void emu_set_engine(int engine) { g_emu_engine = engine; }
// 1: with the asynchronous scheduler, every chunk's chain is first run by a "lane" (mdb_fit_lanes.cuh) and the scheduler only stitches.
void emu_set_lanes(int on) { g_emu_lanes = on; }
void emu_set_lane_warmup(uint32_t points) { g_emu_lane_warmup = points; }
void emu_set_lane_rounds(uint32_t rounds) { g_emu_lane_rounds = rounds; }
uint64_t emu_lane_reruns() { return g_emu_lane_reruns; }
void emu_lane_counters(uint64_t *chunks, uint64_t *bailed) { *chunks = g_emu_lane_chunks; *bailed = g_emu_lane_bailed; }
uint64_t emu_division_mismatches() { return g_emu_division_mismatches; }
void emu_screen_counters(uint64_t *out8, int clear) { for (int i = 0; i < 8; i++) { out8[i] = g_screen_counters[i]; if (clear) g_screen_counters[i] = 0; } }

// The body of k_swing_finish for the model [start, end] of one unit, and the plain loop of swing_finish (mdb_compress.cuh) beside it.
void emu_swing_sums(const int64_t *ts, const float *values, int regular, uint32_t start, uint32_t end, double *out4) {
    const double delta_d = (double)(ts[1] - ts[0]);
    swing_sums_one_lane(ts, values, regular != 0, delta_d, start, end, out4[0], out4[1]);
    const int64_t t0 = ts[start];
    const double v0 = (double)values[start];
    double num = 0.0, den = 0.0;
    for (uint32_t i = start + 2; i <= end; i++) {
        double x, y;
        swing_mse_terms(t0, v0, ts[i], (double)values[i], x, y);
        num = __dadd_rn(num, x);
        den = __dadd_rn(den, y);
    }
    out4[2] = num;
    out4[3] = den;
}

// mdbcu_debug_fit_models on the host: fit_next_model at each start with either engine (records of 40 bytes).
struct EmuDebugFit {
    uint32_t start_index, end_index;
    float min_value, max_value, model_last_value, bytes_per_value;
    int32_t model_type_id, values_len, aborted, irregular;
};
void emu_fit_models(const int64_t *ts, const float *values, uint32_t n, uint8_t eb_kind, float eb_value, int engine, const uint32_t *starts,
                    const uint32_t *budget_ends, uint32_t n_starts, EmuDebugFit *out) {
    ErrorBound eb = make_error_bound(eb_kind, eb_value);
    std::vector<double> smem(WarpFit::SMEM_DOUBLES);
    for (uint32_t k = 0; k < n_starts; k++) {
        FittedModel m;
        bool aborted = false, irregular = false;
        if (engine == 5) {
            const LaneUnit lu = emu_lane_unit(eb, ts, n);
            warp_emu::run([&](int lane) {
                WarpFitScreen f(eb, ts, values, n, smem.data(), &lu);
                f.begin(starts[k]);
                bool ab = false;
                FittedModel mm = f.fit(starts[k], budget_ends[k], ab);
                if (!ab && mm.pending) swing_finish(mm, ts, values);
                if (lane == 0) { m = mm; aborted = ab; irregular = f.irregular(); }
            });
        } else if (engine == 2) {
            warp_emu::run([&](int lane) {
                WarpFit f(eb, ts, values, n, smem.data());
                f.begin(starts[k]);
                bool ab = false;
                FittedModel mm = f.fit(starts[k], budget_ends[k], ab);
                if (!ab && mm.pending) swing_finish(mm, ts, values);
                if (lane == 0) { m = mm; aborted = ab; irregular = f.irregular(); }
            });
        } else {
            ScalarFit f(eb, ts, values, n);
            f.begin(starts[k]);
            m = f.fit(starts[k], budget_ends[k], aborted);
            irregular = f.irregular();
        }
        EmuDebugFit d;
        d.start_index = m.start_index; d.end_index = m.end_index;
        d.min_value = m.min_value; d.max_value = m.max_value; d.model_last_value = m.model_last_value;
        d.bytes_per_value = m.bytes_per_value;
        d.model_type_id = m.model_type_id; d.values_len = m.values_len; d.aborted = aborted; d.irregular = irregular;
        out[k] = d;
    }
}

static uint64_t g_emu_wide_runs = 0;
extern "C" uint64_t emu_wide_runs() { return g_emu_wide_runs; }
// warp_macaque_v_decode (mdb_macaque_warp.cuh) on one stream: out receives `count` values, *last_out the decoder's last value.
void emu_warp_macaque_decode(const uint8_t *bytes, uint64_t n_bytes, uint32_t count, int has_seed, float seed, float *out, float *last_out) {
    std::vector<uint32_t> stage(STAGE_WORDS + 1);
    warp_emu::run([&](int lane) {
        const float last = warp_macaque_v_decode(bytes, n_bytes, count, has_seed != 0, seed, stage.data(), lane,
                                                 [&](uint32_t k0, float value, bool valid) {
                                                     if (valid) out[k0 + lane] = value;
                                                 },
                                                 [&](uint32_t k0, const uint32_t(&v)[WIDE_RUN_PER_LANE]) { // a verified run of 256 `0` codes
                                                     g_emu_wide_runs += lane == 0;
                                                     for (int j = 0; j < WIDE_RUN_PER_LANE; j++) std::memcpy(&out[k0 + WIDE_RUN_PER_LANE * lane + j], &v[j], 4);
                                                 });
        if (lane == 0) *last_out = last;
    });
}

// warp_macaque_v_encode (mdb_macaque_warp.cuh) on one series: first through the counter (how k_records_macaque_warp sizes a
// row), then through the writer (k_emit_macaque_warp).  Returns the bytes written to out (capacity >= 6 * count + 8).
uint64_t emu_warp_macaque_encode(uint8_t eb_kind, float eb_value, const float *values, uint32_t count, uint8_t *out, float *min_out,
                                 float *max_out, uint64_t *counted_bytes) {
    const ErrorBound eb = make_error_bound(eb_kind, eb_value);
    std::vector<uint32_t> stage(STAGE_WORDS + 1);
    uint64_t written = 0;
    warp_emu::run([&](int lane) {
        WarpCodeCounter counter;
        float mn, mx;
        warp_macaque_v_encode(eb, values, 0, count - 1, counter, lane, mn, mx);
        WarpCodeWriter writer;
        writer.init(out, stage.data(), lane);
        float mn2, mx2;
        warp_macaque_v_encode(eb, values, 0, count - 1, writer, lane, mn2, mx2);
        writer.finish();
        if (lane == 0) {
            *counted_bytes = counter.bytes();
            *min_out = mn2;
            *max_out = mx2;
            written = (uint64_t)(writer.out - out);
            if (!(__float_as_uint(mn) == __float_as_uint(mn2) && __float_as_uint(mx) == __float_as_uint(mx2))) written = ~0ull; // both passes agree
        }
    });
    return written;
}

uint64_t emu_segments_len(const EmuSegments *s) { return s->model_type_id.size(); }
void emu_segments_view(const EmuSegments *s, SegmentsView *v, const uint64_t **unit_seg_off) {
    v->n_segments = s->model_type_id.size();
    v->model_type_id = s->model_type_id.data();
    v->start_time = s->start_time.data();
    v->end_time = s->end_time.data();
    v->min_value = s->min_value.data();
    v->max_value = s->max_value.data();
    v->timestamps_off = s->ts_off.data();
    v->timestamps_data = s->ts_data.data();
    v->values_off = s->val_off.data();
    v->values_data = s->val_data.data();
    v->residuals_off = s->res_off.data();
    v->residuals_data = s->res_data.data();
    if (unit_seg_off) *unit_seg_off = s->unit_seg_off.data();
}
void emu_segments_free(EmuSegments *s) { delete s; }

// Returns total points or (uint64_t)-1 on a malformed row. point_off: S+1.
uint64_t emu_grid_count(const SegmentsView *v, uint64_t *point_off) {
    uint64_t total = 0;
    for (uint64_t s = 0; s < v->n_segments; s++) {
        SegDesc d;
        uint32_t len = grid_prepare_segment(*v, s, d);
        if (d.flags & F_MALFORMED) return (uint64_t)-1;
        if (point_off) point_off[s] = total;
        total += len;
    }
    if (point_off) point_off[v->n_segments] = total;
    return total;
}

uint64_t emu_grid(const SegmentsView *v, int64_t *ts_out, float *val_out, uint64_t capacity) {
    uint64_t S = v->n_segments;
    std::vector<SegDesc> desc(S);
    std::vector<uint64_t> po(S + 1, 0);
    for (uint64_t s = 0; s < S; s++) { // prepare + scan
        uint32_t len = grid_prepare_segment(*v, s, desc[s]);
        if (desc[s].flags & F_MALFORMED) return (uint64_t)-1;
        po[s + 1] = po[s] + len;
    }
    if (po[S] > capacity) return (uint64_t)-1;
    std::memset(ts_out, 0xAA, po[S] * 8);
    std::memset(val_out, 0xAA, po[S] * 4);
    for (uint64_t s = 0; s < S; s++) // tile kernel: one "thread" per output point
        for (uint64_t p = po[s]; p < po[s + 1]; p++) grid_point(desc[s], (uint32_t)(p - po[s]), ts_out, val_out, p);
    for (uint64_t s = 0; s < S; s++) // sequential kernel
        if (desc[s].flags & F_SEQUENTIAL)
            grid_sequential_segment(*v, s, desc[s], po[s], (uint32_t)(po[s + 1] - po[s]), ts_out, val_out);
    return po[S];
}

int emu_segment_sums(const SegmentsView *v, float *sums, uint64_t *counts) {
    for (uint64_t s = 0; s < v->n_segments; s++) {
        uint64_t c;
        float sum;
        if (!aggregate_segment(*v, s, c, sum)) return 1;
        sums[s] = sum;
        if (counts) counts[s] = c;
    }
    return 0;
}

int emu_aggregate(const SegmentsView *v, const uint64_t *group_off, uint64_t n_groups, int64_t *count, float *mn,
                  float *mx, double *sum) {
    uint64_t whole[2] = {0, v->n_segments};
    if (!group_off) { group_off = whole; n_groups = 1; }
    for (uint64_t g = 0; g < n_groups; g++) {
        GroupAgg acc = group_agg_identity();
        for (uint64_t s = group_off[g]; s < group_off[g + 1]; s++) {
            uint64_t c;
            float sm;
            if (!aggregate_segment(*v, s, c, sm)) return 1;
            GroupAgg row;
            row.count = (int64_t)c; row.min = v->min_value[s]; row.max = v->max_value[s]; row.sum = (double)sm;
            acc = group_agg_combine(acc, row);
        }
        count[g] = acc.count; mn[g] = acc.min; mx[g] = acc.max; sum[g] = acc.sum;
    }
    return 0;
}

} // extern "C"
