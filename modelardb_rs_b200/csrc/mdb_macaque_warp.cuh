// mdb_macaque_warp.cuh -- a MacaqueV stream decoded by a whole warp (macaque_v.rs:272-323); used by k_grid_macaque_warp and
// k_agg_macaque_warp (mdb_cuda.cu), which own the shared-memory stage and say what happens to each batch of 32 values.
//
// MDB_WARP_EMU: tests/emu/warp_emu.h runs this file on the host, the 32 lanes as cooperative fibers (a debugging
// harness for the GPU-less build container; nothing in the product defines it).
#pragma once

#include "mdb_device.cuh"

#if defined(__CUDACC__) || defined(MDB_WARP_EMU)

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
template <typename OnBatch>
MDB_WARP_FN float warp_macaque_v_decode(const uint8_t *bytes, uint64_t n_bytes, uint32_t count, bool has_seed, float seed,
                                                       uint32_t *stage_words, int lane, OnBatch &&on_batch) {
    WarpBitStage bits;
    bits.init(bytes, n_bytes, stage_words);
    uint64_t p = bits.start_bit;
    uint32_t trailing_zeros = 0;
    uint32_t width_in_force = 32; // payload width of a `0` code: min(32, (32 - leading - trailing) & 0xff), leading = 255 at first
    uint32_t last_value = has_seed ? __float_as_uint(seed) : 0u; // (the first value is "0 XOR 32 raw bits")
    for (uint32_t k0 = 0; k0 < count; k0 += 32) {
        const int cnt = (int)(count - k0 < 32u ? count - k0 : 32u);
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
        if (!has_seed && k0 == 0) { // macaque_v.rs:282-285: the first value is stored raw
            if (lane == 0) mine = rel | (32u << 15);
            rel += 32;
            k_begin = 1;
        }
        if (cnt == 32 && k_begin == 0) {
#pragma unroll
            for (int k = 0; k < 32; k++) walk_one(k);
        } else {
            for (int k = k_begin; k < cnt; k++) walk_one(k);
        }
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
        on_batch(k0, __uint_as_float(value), lane < cnt);
        last_value = __shfl_sync(0xffffffffu, value, cnt - 1);
    }
    return __uint_as_float(last_value);
}


} // namespace mdb

#endif // __CUDACC__ || MDB_WARP_EMU
