// mdb_fit_screen.cuh -- fit_next_model (compression.rs:280-301) by one warp, 32 * P points per step, with Swing's
// decisions SCREENED in f32 and only the doubtful ones evaluated in the reference's own f64 arithmetic.  Results are
// bit-identical to the one-thread fit (mdb_compress.cuh) and to the exact cooperative engine (mdb_fit_warp.cuh), which this
// engine contains and falls back on.
//
// Why.  The exact engine (WarpFitT) spends 58 % of its instructions on Swing: an exact f64 quotient pair for EVERY point,
// f64 prefix scans over them, and a verification walk that re-evaluates the reference's comparisons for every point.  But
// the outcome of a fit only depends on (1) where Swing rejects a point, and (2) the exact slopes of the LAST candidate that
// tightened each bound.  Everything else is a chain of comparisons whose results are obvious for nearly every point.
//
// How.  On a unit with regular timestamps (t_i = t_0 + i * delta; checked once per unit by k_lanes_units / k_lanes_regular)
// every line of a fit passes through its first point (t_0', v_0), so a line is its slope, and with dt = k * delta the four
// comparisons of swing.rs:146-178 at point k are, in real arithmetic, comparisons between slopes per index:
//     reject            <=>  lo_k / k > U   or   hi_k / k < L           (hi_k = v_k + dev_k - v_0, lo_k = v_k - dev_k - v_0)
//     tighten the upper <=>  hi_k / k < U ;   tighten the lower <=>  lo_k / k > L
// and U (L) is the running minimum (maximum) of the candidates hi_r / r (lo_r / r) that tightened.  The reference evaluates
// them in f64 with roundings, so its decision equals the real-arithmetic one whenever the real difference exceeds a bound
// on those roundings.  The screen computes hi_k and lo_k with the reference's own f64 operations (their error is the
// reference's), converts them to f32 slopes (relative error 2^-22), scans the running minimum / maximum over the step in
// f32, and compares with a tolerance tau_k that covers
//     * the f32 errors of both slopes:                  2^-21 (|bound| + |candidate|)
//     * the reference's roundings in slope * t + intercept and in the comparison:
//                                                       2^-50 (|bound| * Tmax / delta + |v_0| + |v_k|) / k
//       (derivation at tolerance()): with timestamps at epoch scale this is the larger term -- the reference's own
//       evaluation is that noisy, which is why its decisions near the boundary can only be reproduced by its own operations.
// A comparison whose |difference| exceeds tau_k is CERTAIN.  A step's points are accepted up to the first certain reject or
// the first doubtful point; for the accepted stretch only the last tightening candidate of each bound is computed exactly
// (one f64 division each, by the reference's formula); a doubtful point is then evaluated exactly as the reference does
// (swing.rs:146-178 on the exact bounds), its decision is applied, and screening resumes after it.  By induction over the
// points the sequence of bounds is the sequential one.  PMC-Mean keeps the exact scan form of the exact engine.
//
// A fit that needs more than a few exact evaluations (constant or exactly linear data: every comparison is a tie), a
// non-finite or huge value, sums that are not exactly representable, a lossless bound, an irregular unit: the exact engine
// runs the fit (fit_exact), and after a few such fits in a row the rest of the chain.
#pragma once

#if defined(__CUDACC__) || defined(MDB_WARP_EMU)

#include "mdb_fit_lanes.cuh"

namespace mdb {

#ifdef MDB_WARP_EMU
static unsigned long long g_screen_counters[8]; // [0] fits, [1] exact fits, [2] passes, [3] exact point evaluations, [4] exact candidates, [5] quiet steps, [6] second rounds
#define MDB_SCREEN_COUNT(i) do { if ((threadIdx.x & 31) == 0) g_screen_counters[i]++; } while (0)
__device__ __forceinline__ float rcp_approx_f32(float x) { return 1.0f / x; }
__device__ __forceinline__ float mdb_fmaf(float a, float b, float c) { return std::fmaf(a, b, c); }
#else
__device__ unsigned long long g_screen_counters[8];
#ifdef MDB_FIT_COUNTERS
#define MDB_SCREEN_COUNT(i) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_screen_counters[i], 1ull); } while (0)
#else
#define MDB_SCREEN_COUNT(i) do { } while (0)
#endif
__device__ __forceinline__ float rcp_approx_f32(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); // one MUFU.RCP, at most 1 ulp off (the arguments are point counts: no subnormals)
    return r;
}
__device__ __forceinline__ float mdb_fmaf(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#endif

#ifndef MDB_SCREEN_MAX_EXACT_POINTS
#define MDB_SCREEN_MAX_EXACT_POINTS 6 // exact point evaluations per fit before the exact engine takes the fit
#endif

template <int P> struct WarpFitScreenT {
    using Exact = WarpFitT<P, false>;
    static constexpr int STEP = 32 * P;
    static constexpr int SMEM_DOUBLES = Exact::SMEM_DOUBLES;

    Exact ex;           // the exact engine: constants, regularity tracking, and the fits this engine does not take
    bool screen;        // this unit's fits may be screened
    double t0u, delta;  // (double)ts[0], (double)(ts[1] - ts[0]) of the unit: (double)ts[i] = t0u + i * delta exactly
    float kap;          // >= 3.1 * 2^-53 * max|t| / delta: the reference's line evaluation noise per unit of slope per index
    uint32_t exact_run; // fits in a row that the exact engine had to take
    bool pmc_close;     // PMC-Mean won the last fit or came close: follow it exactly from the start of the next one (see fit_s)

    __device__ __forceinline__ WarpFitScreenT(const ErrorBound &e, const int64_t *t, const float *v, uint32_t n_, double *smem_, const LaneUnit *lu)
        : ex(e, t, v, n_, smem_), screen(false), t0u(0.0), delta(1.0), kap(0.0f), exact_run(0), pmc_close(false) {
        if (lu != nullptr && n_ >= 2) {
            const LaneUnit u = *lu;
            // (lane_unit_init: positive interval below 2^31, |timestamps| < 2^53, an exact relative test; k_lanes_regular: every interval)
            screen = u.ok && !u.irregular && (e.kind == KIND_RELATIVE || e.kind == KIND_ABSOLUTE);
            t0u = u.t0d;
            delta = u.delta_d;
            const double t_last = __fma_rn((double)(n_ - 1), delta, t0u);
            const double t_max = fmax(fabs(t0u), fabs(t_last));
            // 3.1 u * t_max / delta, rounded up generously (the quotient and the conversion are within 2^-23 of the real value)
            kap = __double2float_rn(__dmul_rn(__ddiv_rn(t_max, delta), SCREEN_C_KAP)) * 1.000001f;
            screen = screen && kap < 1e30f && (e.kind != KIND_RELATIVE || ex.rel_exact_ok) && e.dev >= 0.0 && e.dev < 1e30;
        }
    }

    __device__ __forceinline__ void begin(uint32_t cur) {
        ex.begin(cur);
        exact_run = 0;
    }
    __device__ __forceinline__ bool irregular() const { return ex.irregular(); }
    __device__ __forceinline__ uint32_t skip_rejected(uint32_t from, uint32_t chunk_end, uint32_t budget_end) {
        return ex.skip_rejected(from, chunk_end, budget_end);
    }

    // ---- the exact engine, out of line (its registers must not weigh on the screened step; see fit_scalar_impl) ----
    struct ExactResult {
        FittedModel m;
        uint32_t max_seen;
        bool irregular, aborted;
    };
    static __device__ __noinline__ ExactResult fit_exact_impl(ErrorBound eb, const int64_t *ts, const float *values, uint32_t n, double *smem,
                                                              uint32_t start, uint32_t budget_end, uint32_t max_seen, int64_t delta0, bool irregular) {
        Exact f(eb, ts, values, n, smem);
        f.max_seen = max_seen;
        f.delta0 = delta0;
        f.irregular_ = irregular;
        ExactResult r;
        r.m = f.fit(start, budget_end, r.aborted);
        r.max_seen = f.max_seen;
        r.irregular = f.irregular_;
        return r;
    }
    __device__ __forceinline__ FittedModel fit_exact(uint32_t start, uint32_t budget_end, bool &aborted) {
        MDB_SCREEN_COUNT(1);
        const ExactResult r = fit_exact_impl(ex.eb, ex.ts, ex.values, ex.n, ex.smem, start, budget_end, ex.max_seen, ex.delta0, ex.irregular_);
        ex.max_seen = r.max_seen;
        ex.irregular_ = r.irregular;
        aborted = r.aborted;
        if (exact_run < 0xFFFFu) exact_run++;
        return r.m;
    }

    __device__ __forceinline__ FittedModel fit(uint32_t start, uint32_t budget_end, bool &aborted) {
        // after three exact fits in a row the data is of the kind the screen cannot decide (ties everywhere): stay with the
        // exact engine, and look again every 16 fits
        if (!screen || (exact_run >= 3 && (exact_run & 15u) != 0u)) return fit_exact(start, budget_end, aborted);
        if (ex.eb.kind == KIND_RELATIVE) return fit_s<KIND_RELATIVE>(start, budget_end, aborted);
        return fit_s<KIND_ABSOLUTE>(start, budget_end, aborted);
    }

    // tau_k of the header comment, in slope-per-index units: tau = tw (|upper bound| + |lower bound|) + tz with
    //     tw = C_F32 + kap / k,   tz = C_F32 (|upper candidate| + |lower candidate|) + e_val / k + floor.
    // * C_F32 = 2^-20: every f32 slope is within 2^-21.3 of its real value (conversion of the numerator 2^-24, 1 / k within
    //   2^-22, product 2^-24), and the difference of two of them is rounded once more (2^-24 of itself).
    // * The reference evaluates up = RN(RN(U t) + RN(v0 - RN(U t0))) (swing.rs:146-149).  With M = |U| max|t| and u = 2^-53 the
    //   four roundings add up to |up - line| <= 1.0001 u (3 M + |v0| + |line|); RN(up +- dev) then compares with v like the real
    //   number unless they are within u |v|; and the numerators hi / lo of the candidate differ from the real v +- dev - v0 by at
    //   most u (2 |v| + 2 dev + |v0|).  With |line| <= |D| + dev + |v| (D: the real difference that is being tested) a decision is
    //   the real-arithmetic one if |D| > 1.0001 u (3 M + 2 |v0| + 3 dev + 4 |v|).  M = (slope per index) * max|t| / delta:
    //   kap = 3.1 u max|t| / delta (the slope's f32 image is within 2^-21 of it), and with dev <= |v| for a relative bound
    //   e_val = 4.1 u (|v0| + dev_abs + 2 max|v|) covers the rest.
    // * floor: products that underflow in f32 (values of magnitude 1e-30) have an absolute, not a relative, error.
    static constexpr float SCREEN_C_F32 = 9.5367431640625e-07f; // 2^-20
    static constexpr float SCREEN_C_VAL = 4.552e-16f;           // 4.1 * 2^-53
    static constexpr double SCREEN_C_KAP = 3.4417e-16;          // 3.1 * 2^-53

    template <typename T> static __device__ __forceinline__ T pick(const T (&a)[P], int j) {
        T x = a[0];
#pragma unroll
        for (int i = 1; i < P; i++)
            if (i == j) x = a[i];
        return x;
    }

    // (double)ts[i] and (double)(ts[i] - ts[j]) of a regular unit, from the indices (lane_feed: both are exact)
    __device__ __forceinline__ double time_of(uint32_t i) const { return __fma_rn((double)i, delta, t0u); }

    // The candidate slope the reference stores when point `idx` tightens a bound (swing.rs:151-178 -> 323-340): the line
    // through (t0, v0) and (t, v + dev) (upper) or (t, v - dev) (lower).  Every lane computes it (uniform).
    // k: the point's index within the fit (> 0), v: its value.  The operands of the division are in the range in which
    // ddiv_fast_in_range is the correctly rounded quotient (mdb_fit_warp.cuh: finite f32 values, a time difference >= 1).
    template <int KIND> __device__ __forceinline__ double exact_candidate(uint32_t k, float v, double v0, bool upper) const {
        MDB_SCREEN_COUNT(4);
        const double vd = (double)v;
        const double dev = max_dev_k<KIND>(ex.eb, vd);
        const double target = upper ? __dadd_rn(vd, dev) : __dsub_rn(vd, dev);
        const double dt = __dmul_rn((double)k, delta);
        return v0 == target ? 0.0 : ddiv_fast_in_range(__dsub_rn(target, v0), dt);
    }

    template <int KIND> __device__ __forceinline__ FittedModel fit_s(uint32_t start, uint32_t budget_end, bool &aborted) {
        const int lane = threadIdx.x & 31;
        const int p0 = lane * P;
        const uint32_t n = ex.n;
        const float *values = ex.values;
        const uint32_t limit = budget_end < n ? budget_end : n;
        aborted = false;
        MDB_SCREEN_COUNT(0);

        // PMC-Mean state (pmc_mean.rs:31-53)
        bool pmc_ok = true;
        float p_mn = __uint_as_float(0x7f800000u), p_mx = __uint_as_float(0xff800000u); // (+inf / -inf: replaced by the first value like the reference's NaN)
        double p_sum = 0.0;
        uint32_t p_len = 0;
        unsigned p_umax = 0u, p_umin1 = 0xffffffffu; // largest |value| summed so far / smallest non-zero one minus one, as bit patterns
        // Swing state (swing.rs:34-80): exact slopes, and their f32 images per index for the screen
        bool swing_ok = true;
        double v0 = 0.0;
        uint32_t iu = 1, il = 1; // the points (index within the fit) whose candidates are the bounds in force (valid from the second point on)
        float vu = 0.0f, vl = 0.0f; // ... and their values
        float Ub = __uint_as_float(0x7f800000u), Lb = __uint_as_float(0xff800000u);
        float a0 = 0.0f; // |v0|
        uint32_t s_len = 0;
        int exact_points = 0;
        bool calm = false; // the previous step moved no bound (then a quiet step is likely)
        // PMC-Mean, bounded.  Swing wins 19 fits out of 20 on noisy data, and then PMC-Mean's length only matters as "shorter
        // than 29/30 of Swing's" (types.rs:88-98).  So PMC-Mean is first only BOUNDED from above: a point accepted by it has
        // min and max within the bound of the mean, hence max - min <= bound(|min| + |max|) (relative; 2 * bound absolute),
        // with the reference's roundings (2^-23) covered by the 2^-20 in `yq`.  The first point at which the running range
        // exceeds that is certainly rejected, so p_len <= its index: a running min / max and one comparison per point, no
        // sum, no division.  Only if that bound does not decide the choice of the model is PMC-Mean followed exactly, in a
        // second round over the fit's points (and, while it stays close, from the start of the next fits: pmc_close).
        bool cheap = !pmc_close;
        bool p_dead = false;   // (bounded) a point has certainly been rejected ...
        uint32_t p_bound = 0;  // ... so at most this many were accepted
        float c_mn = __uint_as_float(0x7f800000u), c_mx = __uint_as_float(0xff800000u);
        const float yq = KIND == KIND_RELATIVE ? __fmul_rn(__double2float_rn(ex.rel_mid), 1.000002f) : __fmul_rn(ex.eb.value, 2.000002f);
        bool swing_by_bound = false;

        uint32_t base = start;
        for (;;) { // one round; a second one (PMC-Mean alone, exactly) if the bound does not decide
        base = start;
        float vn[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            const uint32_t idx = start + (uint32_t)(p0 + j);
            vn[j] = idx < limit ? values[idx] : 0.0f;
        }

        while ((!cheap && pmc_ok) || swing_ok) {
            if (base >= limit) { // out of points: the end of the data, or the budget of a speculative chain
                aborted = limit < n;
                break;
            }
            const int cnt = (int)((limit - base) < (uint32_t)STEP ? (limit - base) : (uint32_t)STEP);
            float v[P];
            double vd[P];
            // largest |value| of the step as a bit pattern, and (for PMC-Mean's sums) the smallest non-zero one minus one: slots past
            // `limit` hold 0.0f, which is neutral for both (0 - 1 wraps to the largest unsigned number)
            unsigned umax = 0u, umin1 = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < P; j++) {
                v[j] = vn[j];
                const uint32_t idx = base + (uint32_t)(STEP + p0 + j);
                vn[j] = idx < limit ? values[idx] : 0.0f;
                vd[j] = (double)v[j];
                const unsigned bits = __float_as_uint(v[j]) & 0x7fffffffu;
                umax = max(umax, bits);
                umin1 = min(umin1, bits - 1u);
            }
            umax = __reduce_max_sync(FULL_MASK, umax);
            // NaN, infinity, or a magnitude (1e28) at which the f32 screen could overflow: the exact engine takes the fit
            if (umax >= 0x6e013f39u) return fit_exact(start, budget_end, aborted);
            const float vmax = __uint_as_float(umax);

            // ------------------------------------------------------------------ PMC-Mean (exact: the scan form of WarpFitT::fit_k)
            // Every value here is finite, so f32::min / f32::max (pmc_mean.rs:59-60) are fminf / fmaxf up to the sign of a zero,
            // which nothing that leaves this function depends on: the tests below treat +0 and -0 alike, and the model is the
            // mean.  Slots past `limit` hold 0.0f: they come after every real point in the prefix order and add nothing to a sum.
            if (cheap && !p_dead) { // the bound (see above)
                float lmn[P], lmx[P];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    lmn[j] = j ? fminf(lmn[j - 1], v[j]) : v[0];
                    lmx[j] = j ? fmaxf(lmx[j - 1], v[j]) : v[0];
                }
                float amn = lmn[P - 1], amx = lmx[P - 1];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    amn = fminf(amn, __shfl_up_sync(FULL_MASK, amn, d));
                    amx = fmaxf(amx, __shfl_up_sync(FULL_MASK, amx, d));
                }
                float pmn = __shfl_up_sync(FULL_MASK, amn, 1), pmx = __shfl_up_sync(FULL_MASK, amx, 1);
                if (lane == 0) { pmn = c_mn; pmx = c_mx; }
                else { pmn = fminf(c_mn, pmn); pmx = fmaxf(c_mx, pmx); }
                int dead_j = P;
                float mn_last = 0.0f, mx_last = 0.0f;
#pragma unroll
                for (int j = P - 1; j >= 0; j--) {
                    const float mn = fminf(pmn, lmn[j]), mx = fmaxf(pmx, lmx[j]);
                    if (j == P - 1) { mn_last = mn; mx_last = mx; }
                    const float range = __fsub_rn(mx, mn);
                    const bool dead = KIND == KIND_RELATIVE ? range > __fmul_rn(yq, __fadd_rn(fabsf(mn), fabsf(mx))) : range > yq;
                    if (dead) dead_j = j;
                }
                // (slots past `limit` hold 0.0f and may look rejected: they come after every real point and are not counted)
                const int first_dead = __reduce_min_sync(FULL_MASK, dead_j < P ? p0 + dead_j : IDX_INF);
                if (first_dead < cnt) {
                    p_dead = true;
                    p_bound = (base - start) + (uint32_t)first_dead;
                } else {
                    c_mn = __shfl_sync(FULL_MASK, mn_last, 31);
                    c_mx = __shfl_sync(FULL_MASK, mx_last, 31);
                }
            }
            if (!cheap && pmc_ok) {
                float lmn[P], lmx[P];
                double lS[P];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    lmn[j] = j ? fminf(lmn[j - 1], v[j]) : v[0];
                    lmx[j] = j ? fmaxf(lmx[j - 1], v[j]) : v[0];
                    lS[j] = j ? __dadd_rn(lS[j - 1], vd[j]) : vd[0];
                }
                // every addend is a multiple of 2^q and every partial sum is below 2^(emax + 1 + len_bits): all of them are exact
                // iff that span fits into 53 bits, and then the order of the additions does not matter (mdb_fit_warp.cuh)
                const unsigned n_umax = max(p_umax, umax), n_umin1 = min(p_umin1, __reduce_min_sync(FULL_MASK, umin1));
                const uint32_t total_len = p_len + (uint32_t)cnt;
                const int len_bits = 32 - __clz((int)total_len);
                const int emax = max((int)(n_umax >> 23), 1) - 127, q = max((int)((n_umin1 + 1u) >> 23), 1) - 127 - 23;
                const bool exact = n_umax == 0u || (emax + 1 + len_bits - q) <= 53;
                if (!exact) return fit_exact(start, budget_end, aborted);
                float amn = lmn[P - 1], amx = lmx[P - 1];
                double aS = lS[P - 1];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { // (a lane below d receives its own value back: harmless for min / max)
                    amn = fminf(amn, __shfl_up_sync(FULL_MASK, amn, d));
                    amx = fmaxf(amx, __shfl_up_sync(FULL_MASK, amx, d));
                    const double oS = __shfl_up_sync(FULL_MASK, aS, d);
                    if (lane >= d) aS = __dadd_rn(oS, aS);
                }
                float pmn = __shfl_up_sync(FULL_MASK, amn, 1), pmx = __shfl_up_sync(FULL_MASK, amx, 1);
                double pS = __shfl_up_sync(FULL_MASK, aS, 1);
                if (lane == 0) { pmn = p_mn; pmx = p_mx; pS = p_sum; }
                else { pmn = fminf(p_mn, pmn); pmx = fmaxf(p_mx, pmx); pS = __dadd_rn(p_sum, pS); }
                float mn_last = 0.0f, mx_last = 0.0f;
                double S[P];
                int fail_p = IDX_INF; // (a failure in a slot past cnt is no failure: compared with cnt below)
#pragma unroll
                for (int j = P - 1; j >= 0; j--) {
                    const float mn = fminf(pmn, lmn[j]), mx = fmaxf(pmx, lmx[j]);
                    if (j == P - 1) { mn_last = mn; mx_last = mx; }
                    S[j] = __dadd_rn(pS, lS[j]);
                    const uint32_t len_l = p_len + (uint32_t)(p0 + j) + 1;
                    const float avg = __double2float_rn(ddiv_fast_in_range(S[j], (double)len_l)); // pmc_mean.rs:63
                    bool ok;
                    if (KIND == KIND_RELATIVE) {
                        ok = ex.within_relative(mn, avg) & ex.within_relative(mx, avg);
                    } else {
                        bool no_division_here = false;
                        ok = within_bound_k<KIND, true>(ex.eb, mn, avg, no_division_here) & within_bound_k<KIND, true>(ex.eb, mx, avg, no_division_here);
                    }
                    if (!ok) fail_p = p0 + j;
                }
                fail_p = __reduce_min_sync(FULL_MASK, fail_p);
                if (fail_p >= cnt) { // every point of the step is accepted (after a step cut short by `limit` the state is not used again)
                    p_mn = __shfl_sync(FULL_MASK, mn_last, 31);
                    p_mx = __shfl_sync(FULL_MASK, mx_last, 31);
                    p_sum = __shfl_sync(FULL_MASK, S[P - 1], 31);
                    p_len += (uint32_t)cnt;
                } else { // PMC-Mean ends here: only its length and sum are still needed
                    pmc_ok = false;
                    if (fail_p > 0) {
                        p_sum = __shfl_sync(FULL_MASK, pick(S, (fail_p - 1) % P), (fail_p - 1) / P);
                        p_len += (uint32_t)fail_p;
                    }
                }
                p_umax = n_umax;
                p_umin1 = n_umin1;
            }

            // ------------------------------------------------------------------ Swing (screened)
            if (swing_ok) {
                int lo = 0; // first point of the step not yet processed
                const uint32_t kb = base - start; // points of the fit before this step
                if (s_len == 0) { // swing.rs:106-112: the first point is stored
                    v0 = __shfl_sync(FULL_MASK, vd[0], 0);
                    a0 = fabsf(__shfl_sync(FULL_MASK, v[0], 0));
                    s_len = 1;
                    lo = 1;
                }
                const float dev_abs = KIND == KIND_ABSOLUTE ? __fmul_rn(__double2float_rn(ex.eb.dev), 1.000001f) : 0.0f;
                const float e_val = __fmul_rn(__fadd_rn(__fadd_rn(a0, dev_abs), __fadd_rn(vmax, vmax)), SCREEN_C_VAL);
                const float inf = __uint_as_float(0x7f800000u);
                // Quiet step.  Inside a long model most steps change nothing: every point lies strictly inside the cone, i.e.
                // its candidate interval [lo_k / k, hi_k / k] contains [L, U] with room to spare -- then it neither rejects nor
                // tightens.  In value space that is hi_k - k U > T and k L - lo_k > T, tested in plain f32 (no division, no
                // scan) against one tolerance for the whole step: T = k tau_k of tolerance() at the step's largest k, plus the
                // f32 errors of this test itself (hi_k, lo_k within 2^-22 (|v0| + |v| + dev), k U within 2^-24 of itself).  The two
                // reject tests follow: lo_k < k L - T <= k U - T needs U >= L, which holds for the f32 images (checked) and
                // therefore for the exact bounds up to 2^-21 (|U| + |L|), which T also covers.
                // Tried after a step in which nothing happened; a step that is not quiet takes the normal path below.
                bool quiet = false;
                if (calm && lo == 0 && s_len >= 2 && cnt == STEP && Ub >= Lb) {
                    const float kmaxf = (float)(kb + (uint32_t)STEP);
                    const float w_b = __fadd_rn(fabsf(Ub), fabsf(Lb));
                    const float a_v = __fadd_rn(__fadd_rn(a0, vmax), KIND == KIND_ABSOLUTE ? dev_abs : vmax);
                    const float t_q = __fmul_rn(__fadd_rn(mdb_fmaf(w_b, mdb_fmaf(kmaxf, 1.6e-6f, kap), mdb_fmaf(a_v, 2.4e-6f, e_val)), __fmul_rn(kmaxf, 1e-37f)), 1.000001f);
                    const float v0f = __double2float_rn(v0), c32 = __double2float_rn(ex.eb.dev);
                    bool busy = false;
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        const float w = __fsub_rn(v[j], v0f);
                        const float dv = KIND == KIND_RELATIVE ? __fmul_rn(fabsf(v[j]), c32) : c32;
                        const float kf = (float)(kb + (uint32_t)(p0 + j));
                        const float x = mdb_fmaf(-kf, Ub, __fadd_rn(w, dv)), y = mdb_fmaf(kf, Lb, -__fsub_rn(w, dv));
                        busy |= !(fminf(x, y) > t_q);
                    }
                    quiet = !__any_sync(FULL_MASK, busy);
                    if (quiet) {
                        MDB_SCREEN_COUNT(5);
                        s_len += (uint32_t)cnt;
                    }
                }
                if (!quiet) {
                    calm = true; // until a bound moves or the accepted stretch ends inside the step
                    // candidate slopes per index: the reference's own numerators (swing.rs:151-178 -> 323-340) over k, in f32;
                    // tolerance(): tau = tw * (|upper bound| + |lower bound|) + tz.  Slots that are not points of this fit's step
                    // (the fit's first point, slots past the data) become NEUTRAL candidates (+inf / -inf): they never reject,
                    // never tighten, and all four differences are infinite, so the loops below need no masks.
                    float su[P], sl[P], tw[P], tz[P];
    #pragma unroll
                    for (int j = 0; j < P; j++) {
                        const double dev = max_dev_k<KIND>(ex.eb, vd[j]);
                        const double hi = __dsub_rn(__dadd_rn(vd[j], dev), v0), lw = __dsub_rn(__dsub_rn(vd[j], dev), v0);
                        const uint32_t k = kb + (uint32_t)(p0 + j);
                        const float rk = k ? rcp_approx_f32((float)k) : 0.0f;
                        const float u = __fmul_rn(__double2float_rn(hi), rk), l = __fmul_rn(__double2float_rn(lw), rk);
                        tw[j] = mdb_fmaf(rk, kap, SCREEN_C_F32);
                        tz[j] = mdb_fmaf(__fadd_rn(fabsf(u), fabsf(l)), SCREEN_C_F32, mdb_fmaf(rk, e_val, 1e-37f));
                        const bool real = (p0 + j >= lo) && (p0 + j < cnt);
                        su[j] = real ? u : inf;
                        sl[j] = real ? l : -inf;
                    }

                    while (lo < cnt && swing_ok) {
                        MDB_SCREEN_COUNT(2);
                        // running minimum of the upper / maximum of the lower candidates over the step
                        float am = fminf(su[0], su[1]), ax = fmaxf(sl[0], sl[1]);
    #pragma unroll
                        for (int j = 2; j < P; j++) {
                            am = fminf(am, su[j]);
                            ax = fmaxf(ax, sl[j]);
                        }
    #pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            am = fminf(am, __shfl_up_sync(FULL_MASK, am, d)); // (a lane below d receives its own value back)
                            ax = fmaxf(ax, __shfl_up_sync(FULL_MASK, ax, d));
                        }
                        float ru = __shfl_up_sync(FULL_MASK, am, 1), rl = __shfl_up_sync(FULL_MASK, ax, 1);
                        if (lane == 0) { ru = inf; rl = -inf; }
                        ru = fminf(ru, Ub); // (+inf / -inf until the second point of the fit has set the bounds, swing.rs:126-143)
                        rl = fmaxf(rl, Lb);
                        // this lane's points against the bounds the screen assumes before each of them, up to the lane's first
                        // EVENT: a point that is certainly rejected, or doubtful.  (What a lane computes after its first event,
                        // or at and after another lane's earlier one, is discarded: only points before `stop` count.)
                        int ev_j = P;             // this lane's first event
                        bool ev_certain = false;  // ... is a certain reject
                        int pos_u = -1, pos_l = -1; // this lane's last point before its first event that tightens the upper / the lower bound
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            // (|bounds| capped: the second point of a fit meets infinite bounds, tightens both untested, and must count as
                            // certain; real slopes stay below 5e28 because every |value| is below 1e28, and tw is at most 3.2)
                            const float tau = mdb_fmaf(tw[j], fminf(__fadd_rn(fabsf(ru), fabsf(rl)), 1e30f), tz[j]);
                            const float a = __fsub_rn(sl[j], ru), b = __fsub_rn(rl, su[j]); // > 0: rejected (above the upper / below the lower line)
                            const float c = __fsub_rn(ru, su[j]), d = __fsub_rn(sl[j], rl); // > 0: tightens the upper / the lower bound
                            const float margin = __fsub_rn(fminf(fminf(fabsf(a), fabsf(b)), fminf(fabsf(c), fabsf(d))), tau); // > 0: all four are certain
                            // an event: doubtful (margin <= 0) or rejected (a or b > 0; a zero there is doubtful anyway)
                            const bool event = fmaxf(-margin, fmaxf(a, b)) >= 0.0f;
                            const bool first = event && ev_j == P;
                            if (first) { ev_j = j; ev_certain = margin > 0.0f; }
                            const bool live = ev_j == P;
                            if (live && c > 0.0f) pos_u = j;
                            if (live && d > 0.0f) pos_l = j;
                            ru = fminf(ru, su[j]); // (tightens iff c > 0)
                            rl = fmaxf(rl, sl[j]);
                        }
                        const int first_ev = __reduce_min_sync(FULL_MASK, ev_j < P ? p0 + ev_j : IDX_INF);
                        const int stop = min(first_ev, cnt); // the points [lo, stop) are accepted as the screen assumed
                        // the last tightening candidate of each bound within [lo, stop): its index and value (the exact slope is
                        // computed when it is needed: for a doubtful point, or when the fit ends) and its f32 image
                        {
                            const bool mine = p0 < stop; // (then this lane's points before its first event lie before `stop`)
                            const int gu = __reduce_max_sync(FULL_MASK, mine && pos_u >= 0 ? p0 + pos_u : -1);
                            const int gl = __reduce_max_sync(FULL_MASK, mine && pos_l >= 0 ? p0 + pos_l : -1);
                            if (gu >= 0 || gl >= 0 || stop < cnt) calm = false;
                            if (gu >= 0) {
                                iu = kb + (uint32_t)gu;
                                Ub = __shfl_sync(FULL_MASK, pick(su, gu % P), gu / P);
                                vu = __shfl_sync(FULL_MASK, pick(v, gu % P), gu / P);
                            }
                            if (gl >= 0) {
                                il = kb + (uint32_t)gl;
                                Lb = __shfl_sync(FULL_MASK, pick(sl, gl % P), gl / P);
                                vl = __shfl_sync(FULL_MASK, pick(v, gl % P), gl / P);
                            }
                        }
                        s_len += (uint32_t)(stop - lo);
                        lo = stop;
                        if (stop < cnt) {
                            const int owner = stop / P, jo = stop % P;
                            if (__shfl_sync(FULL_MASK, ev_certain ? 1 : 0, owner)) {
                                swing_ok = false; // certainly rejected
                            } else {
                                // a doubtful point: the reference's own tests (swing.rs:146-178) on the exact bounds
                                MDB_SCREEN_COUNT(3);
                                if (++exact_points > MDB_SCREEN_MAX_EXACT_POINTS || s_len < 2) return fit_exact(start, budget_end, aborted);
                                const uint32_t idx = base + (uint32_t)stop;
                                const float vp = __shfl_sync(FULL_MASK, pick(v, jo), owner);
                                const double us = exact_candidate<KIND>(iu, vu, v0, true), ls = exact_candidate<KIND>(il, vl, v0, false);
                                const double t0d = time_of(start);
                                const double vdp = (double)vp, tdp = time_of(idx);
                                const double dvp = max_dev_k<KIND>(ex.eb, vdp);
                                const double up = __dadd_rn(__dmul_rn(us, tdp), icpt_of(us, v0, t0d));
                                const double lw = __dadd_rn(__dmul_rn(ls, tdp), icpt_of(ls, v0, t0d));
                                if ((__dadd_rn(up, dvp) < vdp) | (__dsub_rn(lw, dvp) > vdp)) {
                                    swing_ok = false;
                                } else {
                                    if (__dsub_rn(up, dvp) > vdp) {
                                        iu = idx - start;
                                        vu = vp;
                                        Ub = __shfl_sync(FULL_MASK, pick(su, jo), owner);
                                    }
                                    if (__dadd_rn(lw, dvp) < vdp) {
                                        il = idx - start;
                                        vl = vp;
                                        Lb = __shfl_sync(FULL_MASK, pick(sl, jo), owner);
                                    }
                                    s_len += 1;
                                    lo = stop + 1;
                                    // the points up to and including this one are done: neutral for the next pass
    #pragma unroll
                                    for (int j = 0; j < P; j++) {
                                        if (p0 + j < lo) { su[j] = inf; sl[j] = -inf; }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            base += (uint32_t)cnt; // cnt < STEP only when `limit` cut the step short
        }
        if (aborted || !cheap) break;
        // Swing has ended (or the data has).  PMC-Mean accepted at most p_bound points: if even that many lose against Swing
        // (the comparison of types.rs:88-98; 29 / len falls with len), Swing is the model whatever PMC-Mean's length is.
        if (p_dead && __fdiv_rn(30.0f, (float)s_len) < __fdiv_rn(29.0f, (float)p_bound)) {
            swing_by_bound = true;
            break;
        }
        MDB_SCREEN_COUNT(6);
        cheap = false;     // second round: PMC-Mean alone and exactly, from the fit's first point
        swing_ok = false;  // (Swing's results stay as they are)
        }

        FittedModel m;
        m.start_index = start;
        m.pending = 0;
        m.pad = 0;
        m.lower_slope = m.upper_slope = 0.0;
        exact_run = 0;
        if (aborted) {
            m.end_index = start;
            m.min_value = m.max_value = m.model_last_value = 0.0f;
            m.bytes_per_value = 1e30f;
            m.model_type_id = PMC_MEAN;
            m.values_len = 0;
            return m;
        }
        const float pmc_bpv = swing_by_bound ? 1e30f : __fdiv_rn(29.0f, (float)p_len); // pmc_mean.rs:83-87
        const float swing_bpv = __fdiv_rn(30.0f, (float)s_len); // swing.rs:236-239
        // follow PMC-Mean exactly from the start of the next fit while it wins or stays within a quarter of winning
        pmc_close = !swing_by_bound && !(swing_bpv < __fdiv_rn(29.0f, __fmul_rn((float)p_len, 1.25f)));
        if (swing_bpv < pmc_bpv) {
            // boundaries and bounds are final; Swing::model (swing.rs:246-259) is completed by swing_finish
            m.model_type_id = SWING;
            m.end_index = start + s_len - 1;
            m.min_value = m.max_value = m.model_last_value = 0.0f;
            m.values_len = 0;
            m.bytes_per_value = swing_bpv;
            m.pending = 1;
            m.lower_slope = s_len >= 2 ? exact_candidate<KIND>(il, vl, v0, false) : (double)__uint_as_float(0x7fc00000u);
            m.upper_slope = s_len >= 2 ? exact_candidate<KIND>(iu, vu, v0, true) : (double)__uint_as_float(0x7fc00000u);
        } else {
            const float value = canonical_nan(__double2float_rn(__ddiv_rn(p_sum, (double)p_len))); // pmc_mean.rs:91-93
            m.model_type_id = PMC_MEAN;
            m.end_index = start + p_len - 1;
            m.min_value = m.max_value = m.model_last_value = value;
            m.values_len = 0;
            m.bytes_per_value = pmc_bpv;
        }
        return m;
    }
};

} // namespace mdb

#endif // __CUDACC__ || MDB_WARP_EMU
