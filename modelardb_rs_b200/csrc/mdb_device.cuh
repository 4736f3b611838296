// mdb_device.cuh -- device-side arithmetic and bit-stream primitives shared by the compress, grid and
// aggregate kernels.  Everything that decides a bit of the output is written with explicit
// round-to-nearest intrinsics (__dmul_rn, __dadd_rn, ...) so that no FMA contraction can occur no
// matter how the translation unit is compiled: the reference is Rust, which never fuses a*b+c.
//
// Reference paths are relative to crates/modelardb_compression/src/.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define MDB_DEV __device__ __forceinline__
#define MDB_DEV_NOINLINE __device__ __noinline__
#else
// Host build of the per-thread bodies, used ONLY by tests/emu (a debugging harness that steps the
// kernels' thread functions in a loop on this GPU-less build container).  Not part of the product: the
// shim lives in tests/emu/ and is only on the include path of the emulator's own build.
#include "mdb_host_shim.h"
#define MDB_DEV inline
#define MDB_DEV_NOINLINE inline
#endif

namespace mdb {

// ---------------------------------------------------------------------------------------------
// Words shared between warps that run at the same time (the chain scheduler, mdb_compress.cuh):
// device-scope atomics, L1-bypassing loads / stores and a device-scope fence.  The host versions in
// mdb_host_shim.h are the single-threaded equivalents.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
MDB_DEV uint32_t sync_cas(uint32_t *p, uint32_t expect, uint32_t val) { return atomicCAS(p, expect, val); }
MDB_DEV uint32_t sync_exch(uint32_t *p, uint32_t val) { return atomicExch(p, val); }
MDB_DEV uint32_t sync_add(uint32_t *p, uint32_t val) { return atomicAdd(p, val); }
MDB_DEV uint32_t sync_load(const uint32_t *p) { return *(const volatile uint32_t *)p; }
MDB_DEV void sync_store(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
MDB_DEV void sync_store8(uint8_t *p, uint8_t v) { *(volatile uint8_t *)p = v; }
MDB_DEV void sync_fence() { __threadfence(); }
MDB_DEV void sync_pause() { __nanosleep(100); }
// 16-byte aligned records written by another SM earlier in the same kernel: read around L1
template <typename T> MDB_DEV T load_shared_record(const T *p) {
    static_assert(sizeof(T) % 16 == 0, "record size");
    union alignas(16) {
        T value;
        uint4 q[sizeof(T) / 16];
    } u;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) u.q[i] = __ldcg(src + i);
    return u.value;
}
#endif

constexpr int KIND_LOSSLESS = 0, KIND_ABSOLUTE = 1, KIND_RELATIVE = 2;
constexpr int PMC_MEAN = 0, SWING = 1, MACAQUE_V = 2;

// ---------------------------------------------------------------------------------------------
// Rust float semantics
// ---------------------------------------------------------------------------------------------

// f32::min / f32::max as x86-64 LLVM lowers llvm.minnum/maxnum: NaN-ignoring, receiver wins ties
// (+0.0 vs -0.0).  Same expression as the oracle's rust_min/rust_max.
MDB_DEV float rust_minf(float a, float b) { return (a != a) ? b : (b < a ? b : a); }
MDB_DEV float rust_maxf(float a, float b) { return (a != a) ? b : (b > a ? b : a); }
MDB_DEV double rust_mind(double a, double b) { return (a != a) ? b : (b < a ? b : a); }
MDB_DEV double rust_maxd(double a, double b) { return (a != a) ? b : (b > a ? b : a); }

// The sign and payload of a NaN PRODUCED BY ARITHMETIC are unspecified in Rust and differ between machines: x86 SSE propagates
// the quieted payload of a NaN operand (0x7fc00000 for the usual f32::NAN input) and GENERATES the default NaN with the sign bit
// set (inf - inf, 0 * inf: 0xffc00000); the GPU generates 0x7fffffff.  Model parameters and sums that come out NaN are
// canonicalised to 0x7fc00000 here.  That equals the reference on x86-64 when the NaN comes from a default-NaN input, and
// differs from it in sign / payload when the NaN is generated (a MacaqueV row's f32 sum over +inf and -inf) or when an input NaN
// carries another payload: NaN-NESS is identical, NaN sign and payload of computed values are the one documented exception to
// bit-identity (the tests compare them with nan_payload_matters=False).  NaNs that are bit COPIES of input
// values (MacaqueV-coded values) keep their payload exactly.
MDB_DEV float canonical_nan(float x) { return (x != x) ? __uint_as_float(0x7fc00000u) : x; }

// models/mod.rs:92-95
MDB_DEV bool equal_or_nan(double a, double b) { return a == b || (a != a && b != b); }
MDB_DEV bool equal_or_nanf(float a, float b) { return a == b || (a != a && b != b); }

struct ErrorBound {
    int kind;
    float value;    // the f32 bound as given
    double dev;     // Absolute: value as f64 * 0.99; Relative: value as f64 / 100.1 (models/mod.rs:83-90)
};

MDB_DEV ErrorBound make_error_bound(int kind, float value) {
    ErrorBound eb;
    eb.kind = kind;
    eb.value = value;
    eb.dev = kind == KIND_ABSOLUTE ? __dmul_rn((double)value, 0.99)
           : kind == KIND_RELATIVE ? __ddiv_rn((double)value, 100.1) : 0.0;
    return eb;
}

// models/mod.rs:53-80
MDB_DEV bool is_value_within_error_bound(const ErrorBound &eb, float real_value, float approximate_value) {
    if (equal_or_nanf(real_value, approximate_value)) return true; // f32 -> f64 is exact, compare in f32
    if (eb.kind == KIND_ABSOLUTE) return fabsf(__fsub_rn(real_value, approximate_value)) <= eb.value;
    if (eb.kind == KIND_RELATIVE) {
        float difference = __fsub_rn(real_value, approximate_value);
        float result = fabsf(__fdiv_rn(difference, real_value));
        return __fmul_rn(result, 100.0f) <= eb.value;
    }
    return false;
}

// models/mod.rs:83-90
MDB_DEV double maximum_allowed_deviation(const ErrorBound &eb, double value) {
    if (eb.kind == KIND_ABSOLUTE) return eb.dev;
    if (eb.kind == KIND_RELATIVE) return fabs(__dmul_rn(value, eb.dev));
    return 0.0;
}

// swing.rs:323-340
MDB_DEV void compute_slope_and_intercept(int64_t start_time, double first_value, int64_t end_time,
                                                            double last_value, double &slope, double &intercept) {
    if (equal_or_nan(first_value, last_value)) {
        slope = 0.0;
        intercept = first_value;
    } else {
        slope = __ddiv_rn(__dsub_rn(last_value, first_value), (double)(end_time - start_time));
        intercept = __dsub_rn(first_value, __dmul_rn(slope, (double)start_time));
    }
}

// (slope * t + intercept) as f32 -- swing.rs:316
MDB_DEV float swing_value(double slope, double intercept, int64_t t) {
    return __double2float_rn(__dadd_rn(__dmul_rn(slope, (double)t), intercept));
}

MDB_DEV float le_f32(const uint8_t *p) {
    return __uint_as_float((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}

// types.rs:307-321 (n is validated to be 0, 1 or 4 by the callers)
MDB_DEV float decode_values_for_pmc_mean(float min_value, float max_value, const uint8_t *values, uint32_t n) {
    if (n == 0) return min_value;
    if (n == 1) return max_value;
    return le_f32(values);
}

// types.rs:374-407; returns false where the reference panics
MDB_DEV bool decode_values_for_swing(float min_value, float max_value, const uint8_t *values, uint32_t n,
                                                        float &first, float &last) {
    if (n == 0) { first = min_value; last = max_value; return true; }
    if (n == 1) { first = max_value; last = min_value; return true; }
    if (n == 5) {
        float v = le_f32(values + 1);
        switch (values[0]) {
        case 0: first = v; last = max_value; return true;
        case 1: first = max_value; last = v; return true;
        case 2: first = min_value; last = v; return true;
        case 3: first = v; last = min_value; return true;
        default: return false;
        }
    }
    if (n == 8) { first = le_f32(values); last = le_f32(values + 4); return true; }
    return false;
}

// ---------------------------------------------------------------------------------------------
// MSB-first bit streams (models/bits.rs).  The reference reads one bit per loop trip
// (bits.rs:61-82); here a 64-bit window is refilled bytewise and fields are extracted with shifts.
// ---------------------------------------------------------------------------------------------

struct BitReader {
    const uint8_t *p;      // next byte to load
    const uint8_t *end;
    uint64_t buf;          // valid bits are the top `nbits`
    int nbits;
    uint64_t consumed;     // bits handed out so far
    uint64_t total_bits;

    MDB_DEV void init(const uint8_t *bytes, uint64_t n_bytes) {
        p = bytes; end = bytes + n_bytes; buf = 0; nbits = 0; consumed = 0; total_bits = 8 * n_bytes;
    }
    MDB_DEV void refill() {
        while (nbits <= 56 && p < end) {
            buf |= (uint64_t)(*p++) << (56 - nbits);
            nbits += 8;
        }
    }
    // n in [0, 32]; bits past the end read as zero (the reference would panic; callers bound it)
    MDB_DEV uint32_t read(int n) {
        if (n == 0) return 0;
        if (nbits < n) refill();
        uint32_t v = (uint32_t)(buf >> (64 - n));
        buf <<= n;
        nbits -= n;
        if (nbits < 0) nbits = 0;
        consumed += n;
        return v;
    }
    MDB_DEV uint64_t read64(int n) { // n in [0, 64]
        if (n <= 32) return read(n);
        uint64_t hi = read(n - 32);
        return (hi << 32) | read(32);
    }
    MDB_DEV bool is_empty() const { return (consumed >> 3) >= (total_bits >> 3); } // bits.rs:45-47
    MDB_DEV uint64_t remaining_bits() const { return consumed >= total_bits ? 0 : total_bits - consumed; }
};

// Counts bits only: used to size the byte columns before they are allocated.
struct BitCounter {
    uint64_t bits = 0;
    MDB_DEV void append(uint64_t, int n) { bits += n; }
    MDB_DEV uint64_t bytes() const { return (bits + 7) >> 3; }
};

// Writes MSB-first into global memory; finish(pad_ones) mirrors finish() / finish_with_one_bits()
// (bits.rs:157-173).
struct BitWriter {
    uint8_t *out;
    uint64_t acc = 0; // pending bits in the low `n` bits
    int n = 0;
    MDB_DEV explicit BitWriter(uint8_t *o) : out(o) {}
    MDB_DEV void append(uint64_t value, int nb) { // nb in [0, 64]; low nb bits of value
        if (nb > 32) {
            append(value >> 32, nb - 32);
            nb = 32;
        }
        if (nb == 0) return;
        acc = (acc << nb) | (value & ((1ull << nb) - 1)); // n < 8 before, so n + 32 <= 39 bits pending
        n += nb;
        while (n >= 8) {
            *out++ = (uint8_t)(acc >> (n - 8));
            n -= 8;
        }
    }
    MDB_DEV void finish(bool pad_ones) {
        if (n > 0) {
            int pad = 8 - n;
            uint8_t b = (uint8_t)(acc << pad);
            if (pad_ones) b |= (uint8_t)((1u << pad) - 1);
            *out++ = b;
            n = 0;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// MacaqueTS (models/timestamps.rs)
// ---------------------------------------------------------------------------------------------

MDB_DEV bool are_compressed_timestamps_regular(const uint8_t *b, uint64_t n) {
    return n == 0 || (b[0] & 128) == 0; // timestamps.rs:199-202
}

MDB_DEV uint64_t be_bytes_to_u64(const uint8_t *b, uint64_t n) {
    uint64_t v = 0;
    for (uint64_t i = 0; i < n && i < 8; i++) v = (v << 8) | b[i];
    return v;
}

// Number of bytes of the regular encoding of `length` (timestamps.rs:99-108).
MDB_DEV uint32_t regular_timestamps_bytes(uint64_t length) {
    int bits = (64 - __clzll((long long)length)) + 1;
    return (uint32_t)((bits + 7) >> 3);
}

// timestamps.rs:283-292
MDB_DEV uint64_t read_decode_and_compute_delta(BitReader &bits, int bits_to_read, uint64_t last_delta) {
    uint64_t encoded = bits.read64(bits_to_read);
    uint64_t dod = encoded;
    if (encoded > (1ull << (bits_to_read - 1)))
        dod = encoded | (bits_to_read >= 64 ? 0ull : (~0ull << bits_to_read));
    return last_delta + dod;
}

// Decoder of the irregular stream (timestamps.rs:228-275) as an iterator over the residual
// timestamps (everything between start_time and end_time).
struct IrregularTimestampDecoder {
    BitReader bits;
    uint64_t last_delta;
    int64_t timestamp;
    MDB_DEV void init(int64_t start_time, const uint8_t *b, uint64_t n) {
        bits.init(b, n);
        bits.read(1);
        last_delta = 0;
        timestamp = start_time;
    }
    // Returns false when the stream is exhausted; otherwise `timestamp` is the next value.
    MDB_DEV bool next() {
        if (bits.is_empty()) return false;
        int leading_one_bits = 0;
        while (leading_one_bits < 5 && !bits.is_empty() && bits.read(1)) leading_one_bits++;
        if (leading_one_bits != 0 && bits.remaining_bits() < 7) return false;
        uint64_t delta;
        switch (leading_one_bits) {
        case 0: delta = last_delta; break;
        case 1: delta = read_decode_and_compute_delta(bits, 7, last_delta); break;
        case 2: delta = read_decode_and_compute_delta(bits, 9, last_delta); break;
        case 3: delta = read_decode_and_compute_delta(bits, 12, last_delta); break;
        case 4: delta = read_decode_and_compute_delta(bits, 32, last_delta); break;
        default: delta = read_decode_and_compute_delta(bits, 64, last_delta); break;
        }
        timestamp = (int64_t)((uint64_t)timestamp + delta);
        last_delta = delta;
        return true;
    }
};

// models/mod.rs:98-124.  The irregular case decodes the whole stream, as the reference does.
MDB_DEV uint64_t segment_len(int64_t start_time, int64_t end_time, const uint8_t *b, uint64_t n) {
    if (n == 0) return start_time == end_time ? 1 : 2;
    if (are_compressed_timestamps_regular(b, n)) return be_bytes_to_u64(b, n);
    IrregularTimestampDecoder d;
    d.init(start_time, b, n);
    uint64_t count = 2;
    while (d.next()) count++;
    return count;
}

// Irregular encoder (timestamps.rs:113-155) over ts[0..n) (n >= 3), Sink = BitCounter | BitWriter.
template <typename Sink>
MDB_DEV void compress_irregular_residual_timestamps(const int64_t *ts, uint64_t n, Sink &sink) {
    sink.append(1, 1);
    int64_t last_timestamp = ts[0];
    int64_t last_delta = 0;
    for (uint64_t i = 1; i + 1 < n; i++) {
        int64_t t = ts[i];
        int64_t delta = (int64_t)((uint64_t)t - (uint64_t)last_timestamp);
        int64_t dod = (int64_t)((uint64_t)delta - (uint64_t)last_delta);
        if (dod == 0) sink.append(0, 1);
        else if (dod >= -63 && dod <= 64) { sink.append(0b10, 2); sink.append((uint64_t)dod, 7); }
        else if (dod >= -255 && dod <= 256) { sink.append(0b110, 3); sink.append((uint64_t)dod, 9); }
        else if (dod >= -2047 && dod <= 2048) { sink.append(0b1110, 4); sink.append((uint64_t)dod, 12); }
        else if (dod >= -2147483647LL && dod <= 2147483648LL) { sink.append(0b11110, 5); sink.append((uint64_t)dod, 32); }
        else { sink.append(0b11111, 5); sink.append((uint64_t)dod, 64); }
        last_delta = delta;
        last_timestamp = t;
    }
}

// timestamps.rs:77-95
MDB_DEV bool are_uncompressed_timestamps_regular(const int64_t *ts, uint64_t n) {
    if (n < 2) return true;
    int64_t expected = ts[1] - ts[0];
    for (uint64_t i = 2; i < n; i++)
        if (ts[i] - ts[i - 1] != expected) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// MacaqueV (models/macaque_v.rs)
// ---------------------------------------------------------------------------------------------

// Bit reader of the MacaqueV decoder: the stream as aligned 32-bit words through a 64-bit window in registers.  Words
// are fetched four at a time (one 16-byte load) into a small queue, a quad ahead of their use: a thread that owns a
// whole row otherwise waits a memory round trip every few codes, and 4-byte loads scattered over 32 rows cost a full
// sector transaction each.  Bits past the end of the stream read as zero; only bytes of the stream are touched: words
// that are not entirely inside it are assembled byte by byte.
struct WordBitReader {
    const uint32_t *words;     // 16-byte aligned address at or before the first byte of the stream
    uint64_t lo_byte, hi_byte; // the stream is bytes [lo_byte, hi_byte) of that word sequence
    uint64_t full_lo, n_full;  // words [full_lo, full_lo + n_full) lie entirely inside the stream
    uint64_t next;             // index of the first word after the queued quad
    uint64_t buf;              // the next `avail` bits of the stream, from the top; zero below them
    int avail;
    uint32_t q0, q1, q2, q3;   // the current quad's remaining words (big-endian bit order), q0 first
    uint32_t a0, a1, a2, a3;   // the quad after it (as loaded), unless ahead_edge
    bool ahead_edge;
    int q_left;

    static MDB_DEV uint32_t big_endian(uint32_t x) {
#ifdef __CUDA_ARCH__
        return __byte_perm(x, 0, 0x0123);
#else
        return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
    }
    struct Quad {
        uint32_t a, b, c, d;
    };
    static MDB_DEV uint32_t edge_word(const uint32_t *words, uint64_t lo_byte, uint64_t hi_byte, uint64_t w) {
        const uint64_t b0 = 4 * w;
        if (b0 >= hi_byte) return 0;
        if (b0 >= lo_byte && b0 + 4 <= hi_byte) return big_endian(words[w]);
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(words);
        uint32_t x = 0;
        for (uint32_t b = 0; b < 4; b++)
            if (b0 + b >= lo_byte && b0 + b < hi_byte) x |= (uint32_t)bytes[b0 + b] << (24 - 8 * b);
        return x;
    }
    // a quad at the first / last bytes of the stream or past its end (one call, arguments by value: the reader stays in
    // registers and the rare path is not replicated at every refill site)
    static MDB_DEV_NOINLINE Quad load_quad_edge(const uint32_t *words, uint64_t lo_byte, uint64_t hi_byte, uint64_t w) {
        Quad q;
        q.a = edge_word(words, lo_byte, hi_byte, w);
        q.b = edge_word(words, lo_byte, hi_byte, w + 1);
        q.c = edge_word(words, lo_byte, hi_byte, w + 2);
        q.d = edge_word(words, lo_byte, hi_byte, w + 3);
        return q;
    }
    // The quad after the current one, as loaded (little-endian words): nothing reads these registers until the current
    // quad is used up, so the load has a whole quad's worth of codes to complete.  A quad at an edge of the stream is
    // only noted here and assembled when it becomes the current one (its value must not merge into the registers the
    // 16-byte load targets, or the merge would wait for the load).
    MDB_DEV void load_ahead(uint64_t w) { // w is a multiple of 4
        ahead_edge = !(w - full_lo < n_full && w + 3 - full_lo < n_full);
        if (!ahead_edge) {
#ifdef __CUDA_ARCH__
            const uint4 x = __ldg(reinterpret_cast<const uint4 *>(words + w));
            a0 = x.x; a1 = x.y; a2 = x.z; a3 = x.w;
#else
            a0 = words[w]; a1 = words[w + 1]; a2 = words[w + 2]; a3 = words[w + 3];
#endif
        }
    }
    MDB_DEV void next_quad() {
        if (ahead_edge) {
            const Quad q = load_quad_edge(words, lo_byte, hi_byte, next - 4);
            q0 = q.a; q1 = q.b; q2 = q.c; q3 = q.d;
        } else {
            q0 = big_endian(a0); q1 = big_endian(a1); q2 = big_endian(a2); q3 = big_endian(a3);
        }
        load_ahead(next);
        next += 4;
        q_left = 4;
    }
    MDB_DEV uint32_t pop_word() {
        const uint32_t x = q0;
        q0 = q1; q1 = q2; q2 = q3;
        if (--q_left == 0) next_quad();
        return x;
    }
    MDB_DEV void init(const uint8_t *bytes, uint64_t n_bytes) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(bytes);
        const uint32_t skip = (uint32_t)(a & 15);
        words = reinterpret_cast<const uint32_t *>(a - skip);
        lo_byte = skip;
        hi_byte = skip + n_bytes;
        full_lo = (lo_byte + 3) / 4;
        n_full = hi_byte / 4 > full_lo ? hi_byte / 4 - full_lo : 0;
        a0 = a1 = a2 = a3 = 0;
        load_ahead(0);
        next = 4;
        next_quad();
        for (uint32_t i = 0; i < skip / 4; i++) pop_word(); // words before the stream
        const uint32_t sub = 8 * (skip & 3);
        buf = (uint64_t)pop_word() << (32 + sub);
        avail = 32 - (int)sub;
    }
    MDB_DEV void ensure(int n) { // n <= 32
        if (avail < n) {
            buf |= (uint64_t)pop_word() << (32 - avail);
            avail += 32;
        }
    }
    MDB_DEV uint32_t peek32() const { return (uint32_t)(buf >> 32); }
    MDB_DEV void skip(int n) { buf <<= n; avail -= n; } // n <= avail
    MDB_DEV uint32_t read(int n) {                       // n in [0, 32]
        ensure(n);
        const uint32_t value = (uint32_t)((buf >> 1) >> (63 - n));
        skip(n);
        return value;
    }
};

// Decoder state machine of macaque_v.rs:272-323 / :220-265.  next() evaluates the three kinds of code without
// branching (threads of a warp own different rows and would diverge at every code): `0` = the XOR's meaningful bits in
// the window in force, `10` = the same value again, `11` = 5 bits of leading zeros, 6 bits of length, the bits.
struct MacaqueVDecoder {
    WordBitReader bits;
    uint32_t width_in_force;  // payload width of a `0` code: min(32, (32 - leading - trailing) & 0xff), leading = u8::MAX at first
    uint32_t trailing_zeros;
    uint32_t last_value;
    // Positions the decoder; when !has_seed the first value is the raw 32 bits and is returned.
    MDB_DEV void init(const uint8_t *b, uint64_t n, bool has_seed, float seed) {
        bits.init(b, n);
        width_in_force = 32;
        trailing_zeros = 0;
        last_value = has_seed ? __float_as_uint(seed) : bits.read(32);
    }
    MDB_DEV float next() {
        bits.ensure(13);
        const uint32_t head = bits.peek32();
        const bool reuse = !(head & 0x80000000u);
        const bool fresh = (head & 0xC0000000u) == 0xC0000000u;
        const uint32_t leading_zeros = (head >> 25) & 31u;
        const uint32_t stored_len = (head >> 19) & 63u;
        const uint32_t new_trailing = (32u - stored_len - leading_zeros) & 0xffu; // u8 wrapping as in release builds
        const uint32_t new_meaningful = (32u - leading_zeros - new_trailing) & 0xffu;
        trailing_zeros = fresh ? new_trailing : trailing_zeros;
        width_in_force = fresh ? (new_meaningful > 32u ? 32u : new_meaningful) : width_in_force;
        bits.skip(reuse ? 1 : (fresh ? 13 : 2));
        uint32_t value = bits.read((reuse | fresh) ? (int)width_in_force : 0);
        value = trailing_zeros < 32u ? value << trailing_zeros : 0u;
        last_value ^= value;
        return __uint_as_float(last_value);
    }
};

MDB_DEV int32_t f32_as_i32(float x) { // Rust `as i32`: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

// 2f32.powi(e) is exact for base 2 (compiler-rt __powisf2); e in [-127, 127] here.
MDB_DEV float powi2(int32_t e) {
    if (e >= -126) return __uint_as_float((uint32_t)(e + 127) << 23);
    return __uint_as_float(0x00400000u); // 2^-127, subnormal
}

// macaque_v.rs:333-336 with release-mode shift masking
MDB_DEV uint32_t rewrite_bits_by_n(uint32_t bits_to_rewrite, int32_t positions_to_shift) {
    return bits_to_rewrite & (0xFFFFFFFFu << ((uint32_t)positions_to_shift & 31u));
}

// `23 - factorized_epsilon.log2().abs().floor() as i32` (macaque_v.rs:185).  The reference calls
// libm's log2f; this computes log2 in f64 and rounds once to f32, which the CPU test-suite shows is
// the same function of x for every f32 (tests/test_oracle_log2.py) and the GPU tests re-check
// against the oracle on the integer-boundary inputs.
MDB_DEV int32_t rewrite_position(float factorized_epsilon) {
    float l = __double2float_rn(log2((double)factorized_epsilon));
    return 23 - f32_as_i32(floorf(fabsf(l)));
}

// macaque_v.rs:168-196
MDB_DEV float rewrite_least_mantissa_bits(const ErrorBound &eb, float value) {
    if (fabsf(value) == 0.0f || value != value || isinf(value)) return value;
    uint32_t value_as_u32 = __float_as_uint(value);
    float abs_error_bound = __double2float_rn(maximum_allowed_deviation(eb, (double)value));
    int32_t exponent = (int32_t)((value_as_u32 >> 23) & 0xff) - 127;
    float factorized_epsilon = __fdiv_rn(abs_error_bound, powi2(exponent));
    int32_t position = rewrite_position(factorized_epsilon);
    float rewritten = __uint_as_float(rewrite_bits_by_n(value_as_u32, position));
    if (!is_value_within_error_bound(eb, value, rewritten)) {
        position -= 1;
        rewritten = __uint_as_float(rewrite_bits_by_n(value_as_u32, position));
    }
    return rewritten;
}

// Encoder state of macaque_v.rs:39-164; Sink = BitCounter | BitWriter.
struct MacaqueVEncoder {
    float min_value, max_value, last_value;
    uint32_t last_leading_zero_bits, last_trailing_zero_bits;
    MDB_DEV void init() {
        min_value = max_value = __uint_as_float(0x7fc00000u);
        last_value = 0.0f;
        last_leading_zero_bits = 255;
        last_trailing_zero_bits = 0;
    }
    MDB_DEV void update(float value) { // :199-204
        min_value = rust_minf(min_value, value);
        max_value = rust_maxf(max_value, value);
        last_value = value;
    }
    template <typename Sink> MDB_DEV void first_raw(float value, Sink &sink) { // :79-83
        sink.append((uint64_t)__float_as_uint(value), 32);
        update(value);
    }
    template <typename Sink> MDB_DEV void compress_value_xor_last_value(const ErrorBound &eb, float value, Sink &sink) {
        if (eb.kind != KIND_LOSSLESS) {
            if (is_value_within_error_bound(eb, value, last_value)) value = last_value;
            else value = rewrite_least_mantissa_bits(eb, value);
        }
        uint32_t x = __float_as_uint(value) ^ __float_as_uint(last_value);
        if (x == 0) {
            sink.append(0b10, 2);
        } else {
            uint32_t lz = (uint32_t)__clz((int)x);
            uint32_t tz = (uint32_t)(__ffs((int)x) - 1);
            if (lz >= last_leading_zero_bits && tz >= last_trailing_zero_bits) {
                uint32_t meaningful_bits = 32u - last_leading_zero_bits - last_trailing_zero_bits;
                sink.append(0, 1);
                sink.append((uint64_t)(x >> last_trailing_zero_bits), (int)meaningful_bits);
            } else {
                uint32_t meaningful_bits = 32u - lz - tz;
                // `11`, 5 bits of lz, 6 bits of length: 13 header bits in one append
                sink.append((uint64_t)((0b11u << 11) | (lz << 6) | meaningful_bits), 13);
                sink.append((uint64_t)(x >> tz), (int)meaningful_bits);
                last_leading_zero_bits = lz;
                last_trailing_zero_bits = tz;
            }
        }
        update(value);
    }
};

// ---------------------------------------------------------------------------------------------
// Segment rows
// ---------------------------------------------------------------------------------------------

struct SegmentsView { // device mirror of mdbcu_segments_view
    uint64_t n_segments;
    const int8_t *model_type_id;
    const int64_t *start_time;
    const int64_t *end_time;
    const float *min_value;
    const float *max_value;
    const uint64_t *timestamps_off;
    const uint8_t *timestamps_data;
    const uint64_t *values_off;
    const uint8_t *values_data;
    const uint64_t *residuals_off;
    const uint8_t *residuals_data;
};

struct Row {
    int model_type_id;
    int64_t start_time, end_time;
    const uint8_t *timestamps; uint64_t n_timestamps;
    float min_value, max_value;
    const uint8_t *values; uint64_t n_values;
    const uint8_t *residuals; uint64_t n_residuals;
};

// The three offset columns are caller data: a row whose offsets run backwards or past the column's last offset
// (off[n_segments], the byte length the caller states for the data array) is loaded as an empty row of an unknown
// model type, which row_is_well_formed rejects -- no decoder ever sees a length derived from a bad offset.
MDB_DEV Row load_row(const SegmentsView &v, uint64_t i) {
    Row r;
    r.model_type_id = v.model_type_id[i];
    r.start_time = v.start_time[i];
    r.end_time = v.end_time[i];
    const uint64_t S = v.n_segments;
    uint64_t a = v.timestamps_off[i], b = v.timestamps_off[i + 1];
    bool ok = a <= b && b <= v.timestamps_off[S];
    r.timestamps = v.timestamps_data + a; r.n_timestamps = b - a;
    r.min_value = v.min_value[i];
    r.max_value = v.max_value[i];
    a = v.values_off[i]; b = v.values_off[i + 1];
    ok = ok && a <= b && b <= v.values_off[S];
    r.values = v.values_data + a; r.n_values = b - a;
    a = v.residuals_off[i]; b = v.residuals_off[i + 1];
    ok = ok && a <= b && b <= v.residuals_off[S];
    r.residuals = v.residuals_data + a; r.n_residuals = b - a;
    if (!ok) {
        r.model_type_id = -1;
        r.n_timestamps = r.n_values = r.n_residuals = 0;
    }
    return r;
}

// Validation of what the reference would panic on (models/mod.rs:170, :237; types.rs:315-319, :391, :405).
MDB_DEV bool row_is_well_formed(const Row &r) {
    if (r.model_type_id == PMC_MEAN) return r.n_values == 0 || r.n_values == 1 || r.n_values == 4;
    if (r.model_type_id == SWING) {
        if (r.n_values == 5) return r.values[0] <= 3;
        return r.n_values == 0 || r.n_values == 1 || r.n_values == 8;
    }
    if (r.model_type_id == MACAQUE_V) return r.n_values >= 4;
    return false;
}

} // namespace mdb
