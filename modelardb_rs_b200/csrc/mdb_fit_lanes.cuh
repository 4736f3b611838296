// mdb_fit_lanes.cuh -- fit_next_model (compression.rs:280-301) and the chain of try_compress_univariate_time_series
// (compression.rs:224-263) with ONE LANE PER CHAIN: every thread walks its own chunk of a unit point by point, exactly
// as the reference's loop does, and the 32 lanes of a warp advance 32 different chains per instruction.
//
// Why this exists beside the warp-cooperative fit (mdb_fit_warp.cuh).  The cooperative fit makes ONE chain fast (128
// points per step) but pays for it: prefix scans, speculated bounds that are verified and sometimes redone, in-order
// sums through shared memory -- 11 warp instructions per data point on the benchmark.  The sequential algorithm itself
// needs ~100 THREAD instructions per point (3 to 4 warp instructions per point when all lanes are busy), no scan, no
// verification and no retry, and a unit cut into chunks offers hundreds of thousands of chains.  So the bulk of the work
// is done here, one lane per chunk, every chain speculative from its chunk's first index (chunk 0 of a unit is exact);
// the exact stitching of the chunks -- re-running a chunk from the exit of the final chain before it until the new
// chain meets the old one -- is a few percent of the points, is latency bound (a unit's chunks are stitched in order),
// and stays with the cooperative fit and the asynchronous scheduler (sched_advance, k_spec_async).
//
// What a lane assumes, and what happens when the assumption fails (never a different result, only a different path):
//   * the unit's timestamps are regular (t_i = t_0 + i * delta, delta > 0) and below 2^53 in magnitude, so (double)t_i and
//     (double)(t_i - t_k) are computed from the index without loading a timestamp (both exact, see lane_feed).  A pre-pass
//     over the timestamps (k_lanes_units / k_lanes_regular, coalesced, bandwidth bound) decides this per unit; other
//     units are left entirely to the cooperative engine, which reads the timestamps.
//   * every value it meets is finite.  A lane that reads a NaN or an infinity abandons its chunk untouched, and the
//     cooperative engine (whose one-thread path is literally the reference's special-value code) runs that chunk.
//   * a fit ends within one chunk length past the end of its chunk; longer fits (long constant or linear stretches)
//     are cut, and resumed by the cooperative engine, which is the right tool for them.
//
// Arithmetic: the same operations in the same order as the one-thread fit (mdb_compress.cuh), with the divisions done
// by ddiv_fast_in_range (operands in range by construction for finite f32 values, see mdb_fit_warp.cuh) and PMC-Mean's
// relative test in its division-free exact form (RelTest).  Swing's two error sums (swing.rs:212-228) are accumulated
// point by point like everything else, so a lane's models are complete (nothing is left `pending` for k_swing_finish).
#pragma once

#if defined(__CUDACC__) || defined(MDB_WARP_EMU)

#include "mdb_fit_warp.cuh"

namespace mdb {

// The relative test of models/mod.rs:53-80 without the division (derivation: WarpFitT::within_relative).
struct RelTest {
    double mid;
    bool mid_passes, exact_ok;
};
MDB_DEV RelTest make_rel_test(const ErrorBound &eb) {
    RelTest r;
    float y = __fdiv_rn(eb.value, 100.0f);
    for (int k = 0; k < 8 && y > 0.0f && __fmul_rn(y, 100.0f) > eb.value; k++) y = __uint_as_float(__float_as_uint(y) - 1u);
    for (int k = 0; k < 8 && __fmul_rn(__uint_as_float(__float_as_uint(y) + 1u), 100.0f) <= eb.value; k++)
        y = __uint_as_float(__float_as_uint(y) + 1u);
    const float y_next = __uint_as_float(__float_as_uint(y) + 1u);
    r.exact_ok = eb.kind == KIND_RELATIVE && y >= 0.0f && y < 1e30f && __fmul_rn(y, 100.0f) <= eb.value && __fmul_rn(y_next, 100.0f) > eb.value;
    r.mid = __dmul_rn(__dadd_rn((double)y, (double)y_next), 0.5);
    r.mid_passes = (__float_as_uint(y) & 1u) == 0u;
    return r;
}

// Per unit, written by k_lanes_units (+ k_lanes_regular): what a lane needs to walk a chunk of the unit.
struct LaneUnit {          // 48 bytes
    double t0d;            // (double)ts[0]
    double delta_d;        // (double)(ts[1] - ts[0])
    double eb_dev;         // ErrorBound::dev
    double rel_mid;        // RelTest::mid
    float eb_value;
    uint8_t kind;          // KIND_*
    uint8_t rel_mid_passes;
    uint8_t ok;            // the lanes may run this unit's chunks
    uint8_t pad;
    uint32_t irregular;    // set by k_lanes_regular when an interval differs from the first one (clears ok)
    uint32_t pad2;
};
static_assert(sizeof(LaneUnit) == 48, "LaneUnit layout");

constexpr double LANES_T_LIMIT = 9007199254740992.0; // 2^53: below it every integer timestamp is an exact double

// The unit's constants, and whether it qualifies at all: at least two points, a positive first interval that the last
// timestamp is consistent with, |timestamps| < 2^53, an exact relative test.  (Whether EVERY interval equals the first
// one is checked separately, over all timestamps: k_lanes_regular.)
MDB_DEV LaneUnit lane_unit_init(const int64_t *uts, uint64_t n, const ErrorBound &eb) {
    LaneUnit lu;
    const RelTest rt = make_rel_test(eb);
    lu.eb_dev = eb.dev;
    lu.eb_value = eb.value;
    lu.kind = (uint8_t)eb.kind;
    lu.rel_mid = rt.mid;
    lu.rel_mid_passes = rt.mid_passes ? 1 : 0;
    lu.pad = 0;
    lu.irregular = 0;
    lu.pad2 = 0;
    lu.t0d = 0.0;
    lu.delta_d = 0.0;
    bool ok = n >= 2 && (eb.kind != KIND_RELATIVE || rt.exact_ok);
    if (ok) {
        const int64_t t0 = uts[0], t1 = uts[1], tl = uts[n - 1];
        const uint64_t delta = (uint64_t)t1 - (uint64_t)t0; // (wraps instead of overflowing; a wrapped value fails the range test)
        ok = t1 > t0 && delta < (1ull << 31) && fabs((double)t0) < LANES_T_LIMIT && fabs((double)tl) < LANES_T_LIMIT &&
             (uint64_t)tl == (uint64_t)t0 + (n - 1) * delta;
        lu.t0d = (double)t0;
        lu.delta_d = (double)(int64_t)delta;
    }
    lu.ok = ok ? 1 : 0;
    return lu;
}

// The state of one fit (ModelBuilder, types.rs:61-144) as a lane keeps it.  While a model accepts every point, its length
// is the number of points fed so far (idx - start): plen / slen only hold a value once the model has rejected a point.
constexpr uint32_t LEN_ALIVE = 0xFFFFFFFFu;
struct LaneFit {
    uint32_t start;  // first point of the fit
    uint32_t idx;    // next point to feed
    double i_d;      // (double)idx
    double k_d;      // (double)(idx - start)
    // PMC-Mean (pmc_mean.rs:31-53); min / max start at +inf / -inf, which a finite first value replaces exactly as
    // f32::min / f32::max replace the reference's initial NaN
    float mn, mx;
    double sum;
    uint32_t plen;
    // Swing (swing.rs:34-80): the intercepts are functions of the slopes (icpt_of); num / den are the two error sums
    double v0, t0d, us, ls, num, den;
    uint32_t slen;

    MDB_DEV void begin(uint32_t at) {
        start = at;
        idx = at;
        i_d = (double)at;
        k_d = 0.0;
        mn = __uint_as_float(0x7f800000u);
        mx = __uint_as_float(0xff800000u);
        sum = 0.0;
        plen = LEN_ALIVE;
        v0 = t0d = us = ls = num = den = 0.0;
        slen = LEN_ALIVE;
    }
    MDB_DEV bool growing() const { return (plen == LEN_ALIVE) | (slen == LEN_ALIVE); }
    MDB_DEV uint32_t pmc_len() const { return plen == LEN_ALIVE ? idx - start : plen; }
    MDB_DEV uint32_t swing_len() const { return slen == LEN_ALIVE ? idx - start : slen; }
};

// models/mod.rs:53-80 for finite values, branch-free.
template <int KIND> MDB_DEV bool lane_within(float eb_value, double rel_mid, bool rel_mid_passes, float real_value, float approx) {
    const bool eq = real_value == approx;
    if (KIND == KIND_ABSOLUTE) return eq | (fabsf(__fsub_rn(real_value, approx)) <= eb_value);
    if (KIND == KIND_RELATIVE) {
        const float diff = __fsub_rn(real_value, approx);
        const double lhs = fabs((double)diff), rhs = __dmul_rn(rel_mid, fabs((double)real_value));
        return eq | (lhs < rhs) | (rel_mid_passes & (lhs == rhs));
    }
    return eq;
}

// try_to_update_models (types.rs:88-118) for the finite value v at index f.idx of a regular unit.
//   (double)t        = t0d + idx * delta_d: the sum is an integer below 2^53, so the fused multiply-add returns it exactly;
//   (double)(t - t0) = RN(k * delta): k and delta are exact doubles, so the rounded product is the conversion of the
//                      integer difference (swing_sums_one_lane in mdb_compress_api.inl uses the same identity).
template <int KIND> MDB_DEV void lane_feed(const LaneUnit &u, LaneFit &f, float v) {
    const double vd = (double)v;
    const double td = __fma_rn(f.i_d, u.delta_d, u.t0d);
    const uint32_t k = f.idx - f.start;       // points fed before this one
    const double k1_d = __dadd_rn(f.k_d, 1.0);
    // ---- PMC-Mean (pmc_mean.rs:58-75); alive: it has accepted all k points, so its length is k
    {
        const bool alive = f.plen == LEN_ALIVE;
        const float nmn = v < f.mn ? v : f.mn, nmx = v > f.mx ? v : f.mx;
        const double nsum = __dadd_rn(f.sum, vd);
        const float avg = __double2float_rn(ddiv_fast_in_range(nsum, k1_d));
        const bool ok = lane_within<KIND>(u.eb_value, u.rel_mid, u.rel_mid_passes != 0, nmn, avg) &
                        lane_within<KIND>(u.eb_value, u.rel_mid, u.rel_mid_passes != 0, nmx, avg);
        if (alive & ok) {
            f.mn = nmn;
            f.mx = nmx;
            f.sum = nsum;
        }
        if (alive & !ok) f.plen = k;
    }
    // ---- Swing (swing.rs:101-198)
    {
        const bool alive = f.slen == LEN_ALIVE;
        const bool first = k == 0, second = k == 1, has = k >= 2;
        const double v0 = first ? vd : f.v0, t0d = first ? td : f.t0d;
        const double dev = max_dev_k<KIND>(ErrorBound{KIND, u.eb_value, u.eb_dev}, vd);
        // the candidate lines through (t0, v0) and (t, v +- dev): swing.rs:323-340.  When v +- dev equals v0 the
        // difference is +0 and so is the quotient, which is the slope the reference's equal-values branch returns.
        const double dt = __dmul_rn(f.k_d, u.delta_d);
        double su, sl;
        ddiv_fast2_in_range(__dsub_rn(__dadd_rn(vd, dev), v0), __dsub_rn(__dsub_rn(vd, dev), v0), first ? 1.0 : dt, su, sl);
        // the bounds in force (zero lines before the second point; never used then)
        const double up = __dadd_rn(__dmul_rn(f.us, td), icpt_of(f.us, v0, t0d));
        const double lw = __dadd_rn(__dmul_rn(f.ls, td), icpt_of(f.ls, v0, t0d));
        const bool rejected = has & ((__dadd_rn(up, dev) < vd) | (__dsub_rn(lw, dev) > vd));
        const bool take_u = second | (has & (__dsub_rn(up, dev) > vd)), take_l = second | (has & (__dadd_rn(lw, dev) < vd));
        // the terms of the two error sums (swing.rs:180-193, 212-228): from the third point on, (0, 0) when v equals v0
        const bool term = has & !(vd == v0);
        const double x = __dmul_rn(__dsub_rn(vd, v0), dt), y = __dmul_rn(dt, dt);
        if (alive & !rejected) {
            f.v0 = v0;
            f.t0d = t0d;
            if (take_u) f.us = su;
            if (take_l) f.ls = sl;
            if (has) {
                f.num = __dadd_rn(f.num, term ? x : 0.0);
                f.den = __dadd_rn(f.den, term ? y : 0.0);
            }
        }
        if (alive & rejected) f.slen = k;
    }
    f.idx += 1;
    f.i_d = __dadd_rn(f.i_d, 1.0);
    f.k_d = k1_d;
}

// ModelBuilder::finish (types.rs:88-144) of a fit that ended on its own (both models failed, or the data ended).
// Returns false when the model would not be stored (bytes_per_value > 4, compression.rs:238): fewer than 8 points.
MDB_DEV bool lane_finish(const LaneUnit &u, const LaneFit &f, FittedModel &m) {
    const uint32_t plen = f.pmc_len(), slen = f.swing_len();
    if (plen < 8 && slen < 8) return false; // 29 / len and 30 / len both exceed 4 bytes per value
    const float pmc_bpv = __fdiv_rn(29.0f, (float)plen);    // pmc_mean.rs:83-87
    const float swing_bpv = __fdiv_rn(30.0f, (float)slen);  // swing.rs:236-239
    m.start_index = f.start;
    m.pad = 0;
    m.pending = 0;
    m.lower_slope = m.upper_slope = 0.0;
    if (swing_bpv < pmc_bpv) { // min_by keeps the first minimum: PMC-Mean wins ties (types.rs:90-94)
        // Swing::model (swing.rs:246-259) and select_swing (types.rs:122-144), as swing_finish_from_sums
        const double projected = __ddiv_rn(f.num, f.den);
        const double slope = rust_maxd(f.ls, rust_mind(projected, f.us));
        const double dt_end = __dmul_rn((double)(slen - 1), u.delta_d); // (double)(end_time - start_time)
        const float first = canonical_nan(__double2float_rn(f.v0));
        const float last = canonical_nan(__double2float_rn(__dadd_rn(__dmul_rn(slope, dt_end), f.v0)));
        m.model_type_id = SWING;
        m.end_index = f.start + slen - 1;
        m.min_value = rust_minf(first, last);
        m.max_value = rust_maxf(first, last);
        m.values_len = (first < last) ? 0 : 1;
        m.model_last_value = last;
        m.bytes_per_value = swing_bpv;
    } else {
        const float value = canonical_nan(__double2float_rn(__ddiv_rn(f.sum, (double)plen))); // pmc_mean.rs:91-93
        m.model_type_id = PMC_MEAN;
        m.end_index = f.start + plen - 1;
        m.min_value = m.max_value = m.model_last_value = value;
        m.values_len = 0;
        m.bytes_per_value = pmc_bpv;
    }
    return m.bytes_per_value <= 4.0f;
}

// One lane's chain over one chunk: the control flow of spec_chain (mdb_compress.cuh) for a chunk that has no earlier
// chain to splice into.
//
// Warm-up.  A chain that starts at an arbitrary index is not the unit's sequential chain: the two only coincide from the
// first fit start they have in common, and on the benchmark's data that takes about eight fits (measured: mean 2500
// points, 90 % within 5800).  A speculative lane therefore starts `warm-up` points BEFORE its chunk and only keeps what
// happens from its first fit start at or after the chunk's first index (`store_from`): that start is the chain's ENTRY.
// The stitching compares it with the exit of the final chain before the chunk -- equal means the lane's chain is the
// sequential chain from there on, as for any other speculative chain; different means the chunk is re-run from the exact
// entry until the new chain meets the lane's.
struct LaneChain {
    uint32_t store_from, chunk_end, limit, n; // fit starts in [store_from, chunk_end) are kept; a fit may read up to `limit` (<= n)
    uint32_t entry;               // first fit start >= store_from (IDX_NONE: not reached yet)
    uint32_t n_models, first_start;
    uint32_t exit, truncated_at;
    bool bailed;                  // met a non-finite value, or was cut before reaching its chunk: the chunk is left to the cooperative engine
    // Re-run of a chunk that already has a chain (a "round", see k_spec_lanes): the new chain runs from its entry only UNTIL it
    // starts a fit where the old chain also started one -- from there on the two are the same chain (fit_next_model is a
    // function of the start index), so the old chain's tail is copied.  The same rule as spec_chain's (mdb_compress.cuh).
    const FittedModel *old_list;  // nullptr: no earlier chain
    uint32_t old_n, old_entry, old_exit, old_trunc, sync_limit, old_p;
    LaneFit fit;

    MDB_DEV void begin(uint32_t start, uint32_t store_from_, uint32_t chunk_end_, uint32_t limit_, uint32_t n_) {
        store_from = store_from_;
        chunk_end = chunk_end_;
        limit = limit_;
        n = n_;
        entry = start >= store_from_ ? start : IDX_NONE;
        n_models = 0;
        first_start = IDX_NONE;
        exit = IDX_NONE;
        truncated_at = 0;
        bailed = false;
        old_list = nullptr;
        old_n = old_entry = old_exit = old_trunc = sync_limit = old_p = 0;
        fit.begin(start);
    }

    // The old chain visited [old_entry, sync_limit) and stored old_n models there; it left the chunk at old_exit (IDX_NONE: cut at old_trunc).
    MDB_DEV void set_old_chain(const FittedModel *list, uint32_t n_old, uint32_t entry_old, uint32_t exit_old, uint32_t trunc_old) {
        old_list = list;
        old_n = n_old;
        old_entry = entry_old;
        old_exit = exit_old;
        old_trunc = trunc_old;
        sync_limit = exit_old == IDX_NONE ? trunc_old : chunk_end;
        old_p = 0;
    }

    // A fit is about to start at `cur`: did the old chain start one there too?  Then its tail is this chain's: returns true, chain complete.
    MDB_DEV bool splice_at(uint32_t cur, FittedModel *list) {
        if (old_list == nullptr || cur < old_entry || cur >= sync_limit) return false;
        while (old_p < old_n && old_list[old_p].end_index < cur) old_p++;
        if (old_p < old_n && old_list[old_p].start_index < cur) return false; // strictly inside an old model: the old chain did not stop here
        if (n_models == 0 && old_p < old_n) first_start = old_list[old_p].start_index;
        for (uint32_t k = old_p; k < old_n; k++) list[n_models + (k - old_p)] = old_list[k];
        n_models += old_n - old_p;
        exit = old_exit;
        truncated_at = old_trunc;
        return true;
    }

    // Feeds one point.  Returns true when the chain is complete (entry / exit / truncated_at / bailed are final).
    template <int KIND> MDB_DEV bool step(const LaneUnit &u, float v, FittedModel *list) {
        if (!(fabsf(v) <= 3.402823466e+38f)) {
            bailed = true;
            return true;
        }
        lane_feed<KIND>(u, fit, v);
        if (fit.growing() && fit.idx < limit) return false;
        if (fit.growing() && limit < n) { // still growing where a speculative chain's budget ends: cut (spec_chain: aborted)
            truncated_at = fit.start;
            bailed = entry == IDX_NONE;
            return true;
        }
        FittedModel m;
        uint32_t next;
        if (lane_finish(u, fit, m)) {
            if (entry != IDX_NONE) {
                if (n_models == 0) first_start = m.start_index;
                list[n_models++] = m;
            }
            next = m.end_index + 1;
        } else {
            next = fit.start + 1; // compression.rs:261: the point becomes a residual
        }
        if (entry == IDX_NONE && next >= store_from) entry = next;
        if (next >= chunk_end) {
            exit = next;
            return true;
        }
        if (splice_at(next, list)) return true;
        fit.begin(next);
        return false;
    }
};

// The chunk state a completed lane chain leaves behind (what spec_chain stores for a fresh chunk).
MDB_DEV void lane_chain_publish(const LaneChain &c, ChunkState &st) {
    st.entry = c.entry;
    st.exit = c.exit;
    st.truncated_at = c.truncated_at;
    st.n_models = c.n_models;
    st.first_start = c.first_start;
    st.buf ^= 1;
    st.dirty = 0;
}

} // namespace mdb

#endif // __CUDACC__ || MDB_WARP_EMU
