// mdb_macaque_warp.cuh -- a MacaqueV stream decoded (macaque_v.rs:272-323) and encoded (macaque_v.rs:39-164) by a whole
// warp; used by k_grid_macaque_warp / k_agg_macaque_warp (mdb_cuda.cu) and k_records_macaque_warp / k_emit_macaque_warp
// (mdb_compress_api.inl), which own the shared-memory stage and say what happens to each batch of 32 values.
//
// MDB_WARP_EMU: tests/emu/warp_emu.h runs this file on the host, the 32 lanes as cooperative fibers (a debugging
// harness for the GPU-less build container; nothing in the product defines it).
#pragma once

#include "mdb_device.cuh"

#include <type_traits>

#if defined(__CUDACC__) || defined(MDB_WARP_EMU)

// 1 (default): the decoder takes whole runs of `0` codes (the XOR in the window in force) and of `10` codes (the same value
// again) in one step instead of walking them one by one.  Exact: every code of a run is verified by its own flag bits.
// Measured on B200 (round 2, 1000 rows of 10^6 values, lossless random walk): grid 72.6 -> 30.5 ms, SUM 87.6 -> 33.5 ms.
// 0 keeps the plain code walk (tests/test_warp_macaque_emulated.py runs both against the oracle).
#ifndef MDB_MACAQUE_SPECULATE_RUNS
#define MDB_MACAQUE_SPECULATE_RUNS 1
#endif

#ifdef MDB_WARP_EMU
#define MDB_WARP_FN inline
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((x << (shift & 31u)) >> 32);
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t, uint32_t selector) { // only the byte swap (0x0123) is used
    (void)selector;
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { // the lanes are cooperative fibers: nothing runs in between
    const uint32_t old = *p;
    *p = old | v;
    return old;
}
#else
#define MDB_WARP_FN __device__ __forceinline__
#endif

namespace mdb {

constexpr uint32_t WIDE_ROW_MIN = 64;
constexpr int WIDE_WARPS = 4;      // warps (rows in flight) per block
constexpr int STAGE_WORDS = 512;   // 2 KiB of the stream per refill (a refill is a synchronous global round trip)

// The row's bytes, staged: STAGE_WORDS big-endian words of the stream starting at word `first_word`.  Bits are
// addressed by their absolute position in the (4-byte aligned) word sequence that contains the stream.
struct WarpBitStage {
    const uint32_t *words; // 4-byte aligned address at or before the first byte of the stream
    uint64_t n_words;
    uint64_t lo_byte, hi_byte; // the stream is bytes [lo_byte, hi_byte) of that word sequence
    uint32_t *stage;       // STAGE_WORDS words of shared memory owned by this warp
    uint64_t first_word;   // stream word held in stage[0]
    uint64_t start_bit;    // position of the stream's first bit

    MDB_WARP_FN void init(const uint8_t *bytes, uint64_t n_bytes, uint32_t *stage_) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(bytes);
        const uint32_t skip = (uint32_t)(a & 3);
        words = reinterpret_cast<const uint32_t *>(a - skip);
        n_words = (skip + n_bytes + 3) / 4;
        lo_byte = skip;
        hi_byte = skip + n_bytes;
        stage = stage_;
        first_word = ~0ull;
        start_bit = 8ull * skip;
    }
    // makes bits [p, p + span) addressable (span <= (STAGE_WORDS - 1) * 32); bits past the stream read as zero
    MDB_WARP_FN void cover(uint64_t p, uint32_t span, int lane) {
        const uint64_t w0 = p >> 5;
        if (first_word != ~0ull && w0 >= first_word && ((p + span + 31) >> 5) < first_word + STAGE_WORDS) return;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < STAGE_WORDS / 32; i++) {
            const uint64_t w = w0 + (uint64_t)(i * 32 + lane);
            uint32_t x = 0;
            if (w < n_words) {
                const uint64_t b0 = 4 * w;
                if (b0 >= lo_byte && b0 + 4 <= hi_byte) {
                    x = __byte_perm(__ldg(words + w), 0, 0x0123); // big-endian bit order
                } else { // the first / last word: only the bytes that belong to the stream are touched
                    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(words);
                    for (uint32_t b = 0; b < 4; b++)
                        if (b0 + b >= lo_byte && b0 + b < hi_byte) x |= (uint32_t)__ldg(bytes + b0 + b) << (24 - 8 * b);
                }
            }
            stage[i * 32 + lane] = x;
        }
        first_word = w0;
        __syncwarp();
    }
    // Software prefetch of the window that starts at bit p: the loads are issued here and land in registers while the caller
    // does something else; prefetch_commit moves them into the stage.  Only whole words of the stream's interior are
    // prefetched (a window that touches the stream's first or last, partial word is left to cover()).
    static constexpr int PRE_WORDS = STAGE_WORDS / 32;
    MDB_WARP_FN void prefetch_issue(uint64_t p, int lane, uint32_t (&pre)[PRE_WORDS], uint64_t &pre_w0) const {
        const uint64_t w0 = p >> 5;
        const bool interior = 4 * w0 >= lo_byte && 4 * (w0 + STAGE_WORDS) <= hi_byte;
        pre_w0 = interior ? w0 : ~0ull;
        if (interior) {
#pragma unroll
            for (int i = 0; i < PRE_WORDS; i++) pre[i] = __ldg(words + w0 + (uint64_t)(i * 32 + lane));
        }
    }
    MDB_WARP_FN void prefetch_commit(int lane, const uint32_t (&pre)[PRE_WORDS], uint64_t pre_w0) {
        if (pre_w0 == ~0ull) return;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PRE_WORDS; i++) stage[i * 32 + lane] = __byte_perm(pre[i], 0, 0x0123);
        first_word = pre_w0;
        __syncwarp();
    }
    // the 32 bits starting at bit `rel` of the stage (rel = absolute position - 32 * first_word; covered)
    MDB_WARP_FN uint32_t peek32(uint32_t rel) const {
        const uint32_t i = rel >> 5;
        return __funnelshift_l(stage[i + 1], stage[i], rel & 31u);
    }
};

// MacaqueVDecoder (mdb_device.cuh, macaque_v.rs:272-323) for one long stream, 32 values per batch.  Where a code
// starts depends on every code before it, so the warp first WALKS the batch's codes serially -- every lane runs the same
// few instructions per code (flag bits, a new window's 11 header bits, the payload's position), nothing else -- and
// lane k keeps the position, width and shift of the k-th payload (packed into one register).  The payloads are then
// extracted by the 32 lanes at once, and the values are an exclusive-or prefix scan over them
// (value k = value k - 1 XOR payload k).  on_batch(k0, value, valid): this lane's value k0 + lane of the stream.
//
// Wide runs.  On high-entropy lossless data the window only ever widens, so after a short warm-up practically every code is a
// `0` code of the same width: the stream is FIXED WIDTH there.  If the code at the cursor is a `0` code, the next 256 codes are
// assumed to be too: lane l looks at the eight slots 8 l .. 8 l + 7 (slot k starts at cursor + k * (1 + width)); a slot whose
// first bit is 0 IS a `0` code provided every slot before it is one, so "all 256 flag bits are zero" proves the whole run, by
// induction over the slots.  Then each lane extracts its eight payloads, XORs them into a lane-local prefix, and ONE warp scan of
// the lanes' totals (instead of eight) turns them into values.  One failed flag anywhere and nothing is used: the ordinary
// 32-value batches take over for a while.  on_wide(k0, v): this lane's values k0 + 8 * lane + 0..7 of the stream.
// Because a run's slots are at known positions, DIFFERENT warps can take different stretches of the same run
// (wide_run_at from any slot boundary): k_grid_macaque_block / k_agg_macaque_block give a long row to a whole block.
struct NoWideRuns {};
constexpr int WIDE_RUN_PER_LANE = 8;
constexpr uint32_t WIDE_RUN = 32 * WIDE_RUN_PER_LANE;

// The 256 slots starting at bit `at` as `0` codes of `width` payload bits: x[j] = XOR of this lane's payloads 0..j (already
// shifted into place).  Returns false (in every lane) if some slot is not a `0` code.
MDB_WARP_FN bool wide_run_at(WarpBitStage &bits, uint64_t at, uint32_t width, uint32_t trailing_zeros, int lane, uint32_t (&x)[WIDE_RUN_PER_LANE]) {
    const uint32_t stride = 1u + width;
    bits.cover(at, WIDE_RUN * 33u + 64u, lane);
    const uint32_t rel = (uint32_t)(at - 32u * bits.first_word);
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < WIDE_RUN_PER_LANE; j++) {
        const uint32_t slot = rel + (uint32_t)(WIDE_RUN_PER_LANE * lane + j) * stride;
        flags |= bits.peek32(slot); // bit 31: the slot's first bit
        uint32_t y = width ? bits.peek32(slot + 1u) >> (32u - width) : 0u;
        y = trailing_zeros < 32u ? y << trailing_zeros : 0u;
        x[j] = j ? x[j - 1] ^ y : y;
    }
    return !__any_sync(0xffffffffu, (flags & 0x80000000u) != 0u);
}

struct WarpMacaqueDecoder {
    WarpBitStage bits;
    uint64_t p;              // bit position of the next code
    uint32_t trailing_zeros;
    uint32_t width_in_force; // payload width of a `0` code: min(32, (32 - leading - trailing) & 0xff), leading = 255 at first
    uint32_t last_value;
    bool first_raw;          // the stream starts with a raw 32-bit value that has not been read yet
#if MDB_MACAQUE_SPECULATE_RUNS
    bool speculate;          // the same in every lane
    int batches_walked;      // batches since the plain walk took over
#endif

    MDB_WARP_FN void init(const uint8_t *bytes, uint64_t n_bytes, bool has_seed, float seed, uint32_t *stage_words) {
        bits.init(bytes, n_bytes, stage_words);
        p = bits.start_bit;
        trailing_zeros = 0;
        width_in_force = 32;
        last_value = has_seed ? __float_as_uint(seed) : 0u; // (the first value is "0 XOR 32 raw bits")
        first_raw = !has_seed;
#if MDB_MACAQUE_SPECULATE_RUNS
        speculate = true;
        batches_walked = 0;
#endif
    }

    // The next `cnt` (<= 32) values; returns this lane's (lane < cnt).
    MDB_WARP_FN uint32_t batch(int cnt, int lane) {
        bits.cover(p, 32u * 45u + 64u, lane);
        const uint32_t rel0 = (uint32_t)(p - 32u * bits.first_word); // positions inside the batch are relative to the stage
        uint32_t rel = rel0;
        uint32_t mine = 0; // payload position (15 bits) | width (6 bits) << 15 | shift (8 bits) << 21
        // Branch-free: the three kinds of code are all evaluated and selected between, so the serial dependency from one
        // code to the next is a shared-memory load, a funnel shift and a handful of dependent integer operations.
        auto walk_one = [&](int k) {
            const uint32_t head = bits.peek32(rel);
            const bool reuse = !(head & 0x80000000u);                  // `0`: the XOR's meaningful bits in the window in force
            const bool fresh = (head & 0xC0000000u) == 0xC0000000u;    // `11`, 5 bits of leading zeros, 6 bits of length, the bits
            const uint32_t leading_zeros = (head >> 25) & 31u;         // (`10`: the same value again, no payload)
            const uint32_t stored_len = (head >> 19) & 63u;
            const uint32_t new_trailing = (32u - stored_len - leading_zeros) & 0xffu; // u8 wrapping as in release builds
            const uint32_t new_meaningful = (32u - leading_zeros - new_trailing) & 0xffu;
            const uint32_t new_width = new_meaningful > 32u ? 32u : new_meaningful;
            trailing_zeros = fresh ? new_trailing : trailing_zeros;
            width_in_force = fresh ? new_width : width_in_force;
            const uint32_t header = reuse ? 1u : (fresh ? 13u : 2u);
            const uint32_t width = (reuse | fresh) ? width_in_force : 0u;
            const uint32_t packed = (reuse | fresh) ? ((rel + header) | (width << 15) | (trailing_zeros << 21)) : 0u;
            rel += header + width;
            if (k == lane) mine = packed;
        };
        int k_begin = 0;
        if (first_raw) { // macaque_v.rs:282-285: the first value is stored raw
            if (lane == 0) mine = rel | (32u << 15);
            rel += 32;
            k_begin = 1;
            first_raw = false;
        }
#if MDB_MACAQUE_SPECULATE_RUNS
        if (speculate) {
            // Runs of equal kinds of code in one step.  The code at `rel` says what kind of run starts there: `0` codes of
            // the width in force (1 + width bits each) or `10` codes (2 bits each); lane l looks at the bits where its
            // code starts IF all codes from k up to it belong to that run, and the first lane whose bits say otherwise
            // ends it (every code before it is thereby verified, and fixes where the next one starts).  A `11` code
            // changes the window and is walked on its own.
            int k = k_begin, steps = 0;
            while (k < cnt) {
                steps++;
                const uint32_t first = bits.peek32(rel) >> 30;
                if (first == 3u) {
                    walk_one(k);
                    k++;
                    continue;
                }
                const bool zero_run = first < 2u;
                const uint32_t stride = zero_run ? 1u + width_in_force : 2u;
                const bool candidate = lane >= k && lane < cnt;
                const uint32_t my_rel = candidate ? rel + (uint32_t)(lane - k) * stride : rel;
                const uint32_t my_bits = bits.peek32(my_rel) >> 30;
                const bool ends_run = candidate && (zero_run ? my_bits >= 2u : my_bits != 2u);
                const unsigned enders = __ballot_sync(0xffffffffu, ends_run);
                const int good = enders ? __ffs((int)enders) - 1 : cnt;
                if (candidate && lane < good) mine = zero_run ? ((my_rel + 1u) | (width_in_force << 15) | (trailing_zeros << 21)) : 0u;
                rel += (uint32_t)(good - k) * stride;
                k = good;
            }
            speculate = steps <= 12; // short runs everywhere: the plain walk is cheaper
            batches_walked = 0;
        } else {
            for (int k = k_begin; k < cnt; k++) walk_one(k);
            speculate = ++batches_walked >= 8; // try runs again every few batches
        }
#else
        if (cnt == 32 && k_begin == 0) {
#pragma unroll
            for (int k = 0; k < 32; k++) walk_one(k);
        } else {
            for (int k = k_begin; k < cnt; k++) walk_one(k);
        }
#endif
        p += rel - rel0;
        uint32_t x = 0;
        const uint32_t my_width = (mine >> 15) & 63u, my_shift = mine >> 21;
        if (lane < cnt && my_width) {
            x = bits.peek32(mine & 0x7fffu) >> (32u - my_width);
            x = my_shift < 32u ? x << my_shift : 0u;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { // inclusive XOR scan
            const uint32_t o = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x ^= o;
        }
        const uint32_t value = last_value ^ x;
        last_value = __shfl_sync(0xffffffffu, value, cnt - 1);
        return value;
    }

    // Tries the next 256 values as one run of `0` codes; on success v holds this lane's eight values and the state has moved on.
    MDB_WARP_FN bool wide(int lane, uint32_t (&v)[WIDE_RUN_PER_LANE]) {
        if (first_raw) return false;
        uint32_t x[WIDE_RUN_PER_LANE];
        if (!wide_run_at(bits, p, width_in_force, trailing_zeros, lane, x)) return false;
        uint32_t t = x[WIDE_RUN_PER_LANE - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { // inclusive XOR scan of the lanes' totals
            const uint32_t o = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t ^= o;
        }
        const uint32_t before = last_value ^ t ^ x[WIDE_RUN_PER_LANE - 1]; // the value before this lane's first one
#pragma unroll
        for (int j = 0; j < WIDE_RUN_PER_LANE; j++) v[j] = before ^ x[j];
        last_value = __shfl_sync(0xffffffffu, v[WIDE_RUN_PER_LANE - 1], 31);
        p += (uint64_t)WIDE_RUN * (1u + width_in_force);
        return true;
    }
};

template <typename OnBatch, typename OnWide = NoWideRuns>
MDB_WARP_FN float warp_macaque_v_decode(const uint8_t *bytes, uint64_t n_bytes, uint32_t count, bool has_seed, float seed,
                                                       uint32_t *stage_words, int lane, OnBatch &&on_batch, OnWide &&on_wide = NoWideRuns()) {
    constexpr bool HAS_WIDE = !std::is_same<typename std::decay<OnWide>::type, NoWideRuns>::value;
    int wide_pause = 0; // 32-value batches to go before the next wide attempt (the same in every lane)
    WarpMacaqueDecoder dec;
    dec.init(bytes, n_bytes, has_seed, seed, stage_words);
    uint32_t k0 = 0;
    while (k0 < count) {
        if constexpr (HAS_WIDE) {
            if (wide_pause == 0 && count - k0 >= WIDE_RUN) {
                uint32_t v[WIDE_RUN_PER_LANE];
                if (dec.wide(lane, v)) {
                    on_wide(k0, v);
                    k0 += WIDE_RUN;
                    continue;
                }
                wide_pause = 8;
            }
            if (wide_pause > 0) wide_pause--;
        }
        const int cnt = (int)(count - k0 < 32u ? count - k0 : 32u);
        const uint32_t value = dec.batch(cnt, lane);
        on_batch(k0, __uint_as_float(value), lane < cnt);
        k0 += 32;
    }
    return __uint_as_float(dec.last_value);
}


// ---- long MacaqueV rows ENCODED by a whole warp -----------------------------------------------------------
// The encoder (macaque_v.rs:39-164) looks serial -- the stored value, the XOR window and the bit position of value
// i all depend on value i - 1 -- but each of the three dependencies is a chain with RESETS whose targets do not
// depend on history, so 32 values are encoded at once, one per lane:
//   * lossy bounds: the stored value stays `last_value` while the input is within the bound of it and is otherwise a
//     function of the input alone (:112-121).  The change points of a batch are found by repeating
//     "every lane tests its value against the current stored value; ballot; the first failing lane starts a new one".
//   * the window (leading, trailing) is kept while the XOR fits into it and is otherwise the XOR's own (:134-156):
//     the same ballot loop, one iteration per new window.
//   * code lengths are then known per lane; a warp scan gives every code its bit offset, and the lanes OR their codes
//     into a zeroed shared-memory stage that leaves in coalesced stores.
// Every decision is the reference's own comparison on the same operands, so the bytes are identical.
// k_records_macaque_warp runs this on a counter to size the row (and to get min / max, an in-order fold),
// k_emit_macaque_warp on the writer at the row's final offset.
struct WarpCodeCounter {
    uint64_t bits = 0;
    MDB_WARP_FN void put(uint64_t, int, uint32_t, uint32_t total_bits) { bits += total_bits; }
    MDB_WARP_FN uint64_t bytes() const { return (bits + 7) >> 3; }
};

struct WarpCodeWriter {
    static constexpr uint32_t STAGE_BITS = STAGE_WORDS * 32;
    uint8_t *out;
    uint32_t *stage; // STAGE_WORDS zeroed words; codes are OR-ed in, MSB first
    uint32_t bitpos; // bits of the stage in use
    int lane;
    MDB_WARP_FN void init(uint8_t *o, uint32_t *stage_, int lane_) {
        out = o; stage = stage_; bitpos = 0; lane = lane_;
        for (int i = lane; i < STAGE_WORDS; i += 32) stage[i] = 0;
        __syncwarp();
    }
    bool word_stores = false; // `out` is 4-byte aligned and stays so (drains write whole words): store words, not bytes
    MDB_WARP_FN void write_bytes(uint32_t n_bytes) {
        uint32_t done = 0;
        if (word_stores) { // the stage holds the stream MSB first: a stored word is the stage word with its bytes reversed
            const uint32_t n_words = n_bytes >> 2;
            uint32_t *out_w = reinterpret_cast<uint32_t *>(out);
            for (uint32_t i = (uint32_t)lane; i < n_words; i += 32) {
                const uint32_t x = stage[i];
                out_w[i] = (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
            }
            done = n_words * 4;
        }
        for (uint32_t i = done + (uint32_t)lane; i < n_bytes; i += 32) out[i] = (uint8_t)(stage[i >> 2] >> (24 - 8 * (i & 3)));
        out += n_bytes;
    }
    MDB_WARP_FN void drain() { // whole words leave; the partial word moves to the front
        __syncwarp();
        const uint32_t n_words = bitpos >> 5;
        write_bytes(n_words * 4);
        __syncwarp();
        const uint32_t partial = stage[n_words & (STAGE_WORDS - 1)];
        __syncwarp();
        for (int i = lane; i < STAGE_WORDS; i += 32) stage[i] = (i == 0 && (bitpos & 31)) ? partial : 0u;
        bitpos &= 31;
        __syncwarp();
    }
    // this lane's code (`len` low bits of `code`, len <= 45) at bit `off` of a batch of `total_bits` bits
    MDB_WARP_FN void put(uint64_t code, int len, uint32_t off, uint32_t total_bits) {
        if (bitpos + total_bits > STAGE_BITS) drain();
        uint32_t p = bitpos + off;
        int remaining = len;
        while (remaining > 0) {
            const int r = (int)(p & 31), take = 32 - r < remaining ? 32 - r : remaining;
            const uint32_t piece = (uint32_t)(code >> (remaining - take)) & (take == 32 ? 0xFFFFFFFFu : ((1u << take) - 1u));
            atomicOr(&stage[p >> 5], piece << (32 - r - take));
            remaining -= take;
            p += (uint32_t)take;
        }
        bitpos += total_bits;
        __syncwarp();
    }
    MDB_WARP_FN void finish() { // zero padding to a whole byte (macaque_v.rs:160-164)
        __syncwarp();
        write_bytes((bitpos + 7) / 8);
    }
};

template <typename Sink>
MDB_WARP_FN void warp_macaque_v_encode(const ErrorBound &eb, const float *__restrict__ values, uint32_t lo, uint32_t hi, Sink &sink,
                                                      int lane, float &min_out, float &max_out) {
    float min_value = __uint_as_float(0x7fc00000u), max_value = min_value; // macaque_v.rs:199-204, folded in order
    float last_value = 0.0f;                                                // stored value before the batch
    uint32_t win_l = 255, win_t = 0;                                        // window before the batch
    for (uint32_t k0 = lo; k0 <= hi; k0 += 32) {
        const int cnt = (int)(hi - k0 + 1 < 32u ? hi - k0 + 1 : 32u);
        const bool in = lane < cnt;
        const float raw = in ? values[k0 + (uint32_t)lane] : 0.0f;
        const bool first_batch = k0 == lo;
        // ---- stored values
        float stored = raw;
        if (eb.kind != KIND_LOSSLESS) {
            float cur = first_batch ? __shfl_sync(0xffffffffu, raw, 0) : last_value; // the first value of a row is stored raw (:79-83)
            int pos = first_batch ? 1 : 0;
            if (first_batch && lane == 0) stored = raw;
            while (true) {
                const bool changes = in && lane >= pos && !is_value_within_error_bound(eb, raw, cur);
                const unsigned m = __ballot_sync(0xffffffffu, changes);
                const int j = m ? __ffs(m) - 1 : 32;
                if (lane >= pos && lane < j) stored = cur;
                if (j == 32) break;
                const float fresh = rewrite_least_mantissa_bits(eb, __shfl_sync(0xffffffffu, raw, j));
                if (lane == j) stored = fresh;
                cur = fresh;
                pos = j + 1;
            }
        }
        // ---- XOR with the previous stored value
        float prev = __shfl_up_sync(0xffffffffu, stored, 1);
        if (lane == 0) prev = last_value;
        const bool is_raw = first_batch && lane == 0; // 32 raw bits, no XOR code
        const uint32_t x = (in && !is_raw) ? (__float_as_uint(stored) ^ __float_as_uint(prev)) : 0u;
        const uint32_t lz = x ? (uint32_t)__clz((int)x) : 32u, tz = x ? (uint32_t)(__ffs((int)x) - 1) : 0u;
        // ---- windows
        uint32_t my_l = win_l, my_t = win_t;
        bool is_new = false;
        {
            int pos = 0;
            while (true) {
                const bool resets = lane >= pos && x != 0u && !(lz >= win_l && tz >= win_t);
                const unsigned m = __ballot_sync(0xffffffffu, resets);
                const int j = m ? __ffs(m) - 1 : 32;
                if (lane >= pos && lane < j) { my_l = win_l; my_t = win_t; }
                if (j == 32) break;
                win_l = __shfl_sync(0xffffffffu, lz, j);
                win_t = __shfl_sync(0xffffffffu, tz, j);
                if (lane == j) { my_l = win_l; my_t = win_t; is_new = true; }
                pos = j + 1;
            }
        }
        // ---- codes
        uint64_t code = 0;
        int len = 0;
        if (in) {
            if (is_raw) {
                code = __float_as_uint(stored);
                len = 32;
            } else if (x == 0u) {
                code = 0b10;
                len = 2;
            } else {
                const uint32_t meaningful = 32u - my_l - my_t;
                if (is_new) {
                    code = ((uint64_t)((0b11u << 11) | (my_l << 6) | meaningful) << meaningful) | (uint64_t)(x >> my_t);
                    len = 13 + (int)meaningful;
                } else {
                    code = (uint64_t)(x >> my_t); // a leading `0` flag, then the meaningful bits
                    len = 1 + (int)meaningful;
                }
            }
        }
        uint32_t end = (uint32_t)len; // inclusive scan of the lengths
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, end, d);
            if (lane >= d) end += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, end, 31);
        sink.put(code, len, end - (uint32_t)len, total);
        // ---- min / max over the stored values, in order (rust_minf / rust_maxf are "leftmost" folds: associative)
        float mn = in ? stored : __uint_as_float(0x7fc00000u), mx = mn;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float on = __shfl_up_sync(0xffffffffu, mn, d), ox = __shfl_up_sync(0xffffffffu, mx, d);
            if (lane >= d) {
                mn = rust_minf(on, mn);
                mx = rust_maxf(ox, mx);
            }
        }
        min_value = rust_minf(min_value, __shfl_sync(0xffffffffu, mn, 31));
        max_value = rust_maxf(max_value, __shfl_sync(0xffffffffu, mx, 31));
        last_value = __shfl_sync(0xffffffffu, stored, cnt - 1);
    }
    min_out = min_value;
    max_out = max_value;
}

} // namespace mdb

#endif // __CUDACC__ || MDB_WARP_EMU
