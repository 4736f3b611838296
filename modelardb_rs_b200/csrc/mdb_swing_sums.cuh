// mdb_swing_sums.cuh -- the two error sums of one pending Swing model (swing.rs:212-228), added by ONE thread in point order:
// the body of k_swing_finish (mdb_compress_api.inl), a lane per model.  In a header of its own so that tests/emu can run it on the
// host against the plain loop of swing_finish (mdb_compress.cuh) for every alignment and length.
#pragma once
#include "mdb_compress.cuh"

namespace mdb {

// Four consecutive values, 16-byte aligned.
struct ValueQuad {
    float x, y, z, w;
};
MDB_DEV ValueQuad load_value_quad(const float *p) {
    ValueQuad q;
#ifdef __CUDACC__
    // (volatile: the load stays where it is written, in registers of its own; a plain __ldg was sunk by ptxas to just before its
    // first use, into the registers of the quad consumed last, which is no lead at all)
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p));
#else
    q.x = p[0];
    q.y = p[1];
    q.z = p[2];
    q.w = p[3];
#endif
    return q;
}

// uts / uval: the unit's columns; [start, end]: the model's points (the first two add no term).
// regular: the unit's timestamps are t0 + i * delta with 0 <= delta < 2^31 (then the timestamps are not read at all).
MDB_DEV void swing_sums_one_lane(const int64_t *__restrict__ uts, const float *__restrict__ uval, bool regular, double delta_d, uint32_t start,
                                 uint32_t end, double &num, double &den) {
    const double v0 = (double)uval[start];
    num = 0.0;
    den = 0.0;
    uint32_t i = start + 2;
    if (regular) {
        // t[i] - t[start] = k * interval exactly; k and the interval are exact doubles, so the rounded product is
        // the same double as the conversion of the integer difference
        double kd = 2.0;
        auto term = [&](float vf) {
            const double v = (double)vf;
            const double dt = __dmul_rn(kd, delta_d);
            const bool eq = equal_or_nan(v0, v);
            const double x = __dmul_rn(__dsub_rn(v, v0), dt), y = __dmul_rn(dt, dt);
            num = __dadd_rn(num, eq ? 0.0 : x);
            den = __dadd_rn(den, eq ? 0.0 : y);
            kd = __dadd_rn(kd, 1.0);
        };
        auto terms = [&](const ValueQuad &q) {
            term(q.x);
            term(q.y);
            term(q.z);
            term(q.w);
        };
        for (; i <= end && (reinterpret_cast<uintptr_t>(uval + i) & 15); i++) term(uval[i]);
        // Quads of four values, loaded TWO quads ahead of the additions into three buffers that take turns.  A lane is a stream
        // of its own with nothing else to do while a load is under way: with one load in flight per lane the kernel read at
        // 2.0 of the 6.4 TB/s (ncu, profiles/r02_swing_finish_ncu_full.txt: 72 % of the stall samples on the first use of the
        // quad just loaded).  The loop is unrolled by the three buffers by hand: rotating one set of names makes ptxas copy
        // registers at the top of the body, and a copy waits for the load it copies (measured: no gain at all; this form
        // 1.96 -> 1.66 ms per 10^9 points).  Only quads inside the model are loaded.
        if (i + 3 <= end) {
            const float *quads = uval + i;
            const uint32_t n_quads = (end - i + 1) / 4;
            ValueQuad a = load_value_quad(quads), b = a, c = a;
            if (n_quads > 1) b = load_value_quad(quads + 4);
            uint32_t k = 0; // the quad that is added next
#pragma unroll 1
            for (;;) {
                if (k + 2 < n_quads) c = load_value_quad(quads + 4 * (k + 2));
                terms(a);
                if (++k == n_quads) break;
                if (k + 2 < n_quads) a = load_value_quad(quads + 4 * (k + 2));
                terms(b);
                if (++k == n_quads) break;
                if (k + 2 < n_quads) b = load_value_quad(quads + 4 * (k + 2));
                terms(c);
                if (++k == n_quads) break;
            }
            i += 4 * n_quads;
        }
        for (; i <= end; i++) term(uval[i]);
    } else {
        const int64_t t0 = uts[start];
        for (; i <= end; i++) {
            double x, y;
            swing_mse_terms(t0, v0, uts[i], (double)uval[i], x, y);
            num = __dadd_rn(num, x);
            den = __dadd_rn(den, y);
        }
    }
}

} // namespace mdb
