// mdb_cuda.cu -- kernels and C-ABI of libmodelardb_cuda.so (include/modelardb_cuda.h), sm_100a.
//
// Kernel inventory (DESIGN.md has the roofline of each):
//   scan_*                exclusive prefix sums (point offsets, row offsets, byte offsets)
//   k_grid_prepare        one thread per segment row -> SegDesc + point count (+ worklist of serial rows)
//   k_grid_tile_index     first segment row of every 2048-point output tile
//   k_grid_tile           one thread per OUTPUT POINT: streaming, coalesced writes of timestamps/values
//   k_grid_sequential     irregular timestamps, MacaqueV values, residuals (serial per row)
//   k_agg_segments        one thread per row: COUNT and SUM of the row from its model
//   k_agg_partial/final   deterministic in-order tree reduction per group
//   k_spec_chain          one thread per chunk of a unit: greedy PMC-Mean/Swing chain (speculative, see
//                         mdb_compress.cuh); k_spec_propagate / k_spec_finalize stitch chunks per unit
//   k_spec_records        accepted models -> SegRecords in final row order
//   k_compress_gather     row metadata + per-row byte lengths
//   k_compress_emit       one thread per row: MacaqueTS / MacaqueV byte columns
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/modelardb_cuda.h"
#include "mdb_aggregate.cuh"
#include "mdb_compress.cuh"
#include "mdb_fit_warp.cuh"
#include "mdb_fit_lanes.cuh"
#include "mdb_fit_screen.cuh"
#include "mdb_grid.cuh"
#include "mdb_macaque_warp.cuh"

using namespace mdb;

// ------------------------------------------------------------------------------------------------
// errors, context, buffers
// ------------------------------------------------------------------------------------------------

static thread_local std::string g_last_error;

static int fail(const std::string &msg) {
    g_last_error = msg;
    return MDBCU_FAILURE;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(std::string(#expr) + ": " + cudaGetErrorName(e_) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct KernelStat {
    std::string name;
    double ms = 0.0;
    uint64_t launches = 0;
};

struct PendingTiming {
    size_t stat;
    cudaEvent_t start, stop;
};

struct PinnedBlock {
    uint8_t *p = nullptr;
    size_t cap = 0;
};

// Deferred copy out of the bounce ring into pageable caller memory.
struct PendingOut {
    void *dst;
    const uint8_t *src;
    size_t bytes;
};

// Two pinned halves used in turn: the DMA of one half overlaps the host memcpy into / out of the other.
struct Stager {
    static constexpr size_t HALF = 32u << 20;
    uint8_t *base = nullptr;
    int cur = 0;
    size_t head = 0;
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool in_flight[2] = {false, false};
    std::vector<PendingOut> out[2];
};

struct mdbcu_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // a second stream for work that only has to be finished by the end of a call (compress: the regularity check of the
    // timestamps, which is bandwidth bound, beside the chain kernel, which is not), and the two events that order it
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_ready = nullptr, aux_done = nullptr;
    bool overlap_regular_check = true;
    uint64_t launches = 0;
    int sm_count = 148;
    bool fit_wide = false;          // cooperative engine with the 512-point wide steps also when it runs alone (tuning)
    int block_row_warps = 4;        // warps per row of k_macaque_block (2, 4, 8 or 16)
    uint32_t block_row_min = 0;     // MacaqueV values from which a row is decoded by a whole block (set to BLOCK_ROW_MIN at creation)
    uint32_t lane_rows_min = 24576; // long MacaqueV rows per batch from which one thread owns a row (LANE_ROWS_MIN)
    uint32_t chunk_len_override = 0; // 0: choose_chunk_len() decides
    bool grid_tile_scan = false;     // k_grid_tile (rows of a tile by head flags + max-scan) instead of k_grid_tile_search (binary search per quad)
    bool grid_tma_stores = false;    // k_grid_tile_tma (tiles staged in shared memory, stored by the TMA engine) instead of k_grid_tile: measured slower, see there
    bool lane_rounds_by_lanes = false; // repair rounds after the lanes' first pass: by lanes too, or (default) by the cooperative engine
    uint32_t lane_warmup = 4096;     // points a speculative lane chain starts before its chunk (LaneChain, mdb_fit_lanes.cuh)
    uint32_t last_rounds = 0;        // chain rounds of the last mdbcu_compress
    int fit_mode = 0;                // mdbcu_context_set_fit_engine
    // optional per-kernel CUDA-event timing (mdbcu_context_set_profiling)
    bool profiling = false;
    std::vector<KernelStat> stats;
    std::vector<PendingTiming> pending;
    std::vector<cudaEvent_t> event_pool;
    // host <-> device plumbing (see "host transfers" below)
    uint64_t *mailbox = nullptr, *mailbox_dev = nullptr; // mapped pinned words that kernels post scalars into
    Stager stager;                                       // pinned bounce ring for pageable caller memory
    std::vector<PinnedBlock> pinned_cache;               // reusable pinned blocks for library-owned host copies

    size_t stat_index(const char *raw) {
        std::string name(raw); // (a template instance with two arguments is written in parentheses at the launch site)
        if (name.size() >= 2 && name.front() == '(' && name.back() == ')') name = name.substr(1, name.size() - 2);
        for (size_t i = 0; i < stats.size(); i++)
            if (stats[i].name == name) return i;
        stats.push_back(KernelStat{name});
        return stats.size() - 1;
    }
    cudaEvent_t get_event() {
        if (!event_pool.empty()) {
            cudaEvent_t e = event_pool.back();
            event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void resolve_timings() { // caller has synchronised the stream
        for (PendingTiming &p : pending) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, p.start, p.stop) == cudaSuccess) stats[p.stat].ms += ms;
            event_pool.push_back(p.start);
            event_pool.push_back(p.stop);
        }
        pending.clear();
    }
};

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                     \
        if ((ctx)->profiling) {                                              \
            PendingTiming pt_;                                               \
            pt_.stat = (ctx)->stat_index(#kernel);                           \
            pt_.start = (ctx)->get_event();                                  \
            pt_.stop = (ctx)->get_event();                                   \
            cudaEventRecord(pt_.start, (ctx)->stream);                       \
            kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
            cudaEventRecord(pt_.stop, (ctx)->stream);                        \
            (ctx)->stats[pt_.stat].launches++;                               \
            (ctx)->pending.push_back(pt_);                                   \
        } else {                                                             \
            kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        }                                                                    \
        (ctx)->launches++;                                                   \
    } while (0)

// Stream-ordered device buffer (cudaMallocAsync: after warm-up this is a pool hit, not a driver call).
template <typename T> struct DBuf {
    T *p = nullptr;
    cudaStream_t s = nullptr;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    cudaError_t alloc(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        return cudaMallocAsync((void **)&p, (n ? n : 1) * sizeof(T), stream);
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    T *take() { T *q = p; p = nullptr; return q; }
};

struct Status {              // reset before each call (new_status)
    unsigned int bad;        // some row / unit was malformed
    unsigned int n_seq;      // rows in the serial worklist (filled from the front of the worklist array)
    unsigned int first_bad;  // smallest malformed row / unit index
    unsigned int n_wide;     // long MacaqueV rows decoded by a whole warp (filled from the back of the worklist array)
};

static inline unsigned int div_up(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// host transfers
//  * Scalars the host needs between launches (scan totals, the dirty count of a round, the status
//    word) are posted by a one-warp kernel into mapped pinned memory.  A cudaMemcpy of 8 bytes would
//    queue on the device-to-host copy engine behind whatever bulk copy another context has in flight,
//    which serialises contexts that are meant to pipeline.
//  * Bulk copies go straight to / from caller memory when it is pinned.  Pageable memory (Arrow
//    buffers are) is bounced through the context's pinned ring, 32 MiB halves used in turn so that the
//    DMA of one half overlaps the host memcpy of the other.
//  * sync_ctx() is the only way an entry point waits for the stream: it also completes the bounced
//    device-to-host copies.
// ------------------------------------------------------------------------------------------------

constexpr uint32_t MAILBOX_WORDS = 64;
constexpr uint32_t SLOT_STATUS = 60; // Status is two words; entry points use slots below this one freely

__global__ void k_post(uint64_t *mailbox, const uint64_t *src, uint32_t n_words) {
    if (threadIdx.x < n_words) mailbox[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

static cudaError_t post(mdbcu_context *ctx, uint32_t slot, const void *d_src, uint32_t n_words) {
    LAUNCH(ctx, k_post, 1, 32, 0, ctx->mailbox_dev + slot, (const uint64_t *)d_src, n_words);
    return cudaGetLastError();
}

// Pinned (page-locked) host memory is also mapped into the device's address space (unified addressing);
// *device_alias receives the pointer kernels can use for it, or nullptr.
static bool is_pinned(const void *p, void **device_alias = nullptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (device_alias) *device_alias = a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Copies of up to SMALL_COPY bytes between device memory and PINNED host memory are done by a kernel through the
// mapping instead of by a copy engine: a copy engine serves its queue in order, so a 30 MB column would wait behind
// any multi-gigabyte copy another context has queued in the same direction (tens of milliseconds), while loads and
// stores issued by SMs share the link with it.
constexpr size_t SMALL_COPY = 64u << 20;

__global__ void __launch_bounds__(256) k_copy_bytes(uint8_t *dst, const uint8_t *src, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const size_t n16 = n / 16;
        for (size_t k = i; k < n16; k += stride) reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(src)[k];
        for (size_t k = n16 * 16 + i; k < n; k += stride) dst[k] = src[k];
    } else {
        for (size_t k = i; k < n; k += stride) dst[k] = src[k];
    }
}

static cudaError_t copy_by_kernel(mdbcu_context *ctx, void *dst, const void *src, size_t bytes) {
    const unsigned int blocks = (unsigned int)std::min<size_t>((size_t)ctx->sm_count * 4, (bytes / 16 + 255) / 256 + 1);
    LAUNCH(ctx, k_copy_bytes, blocks, 256, 0, (uint8_t *)dst, (const uint8_t *)src, bytes);
    return cudaGetLastError();
}

static cudaError_t stager_retire(mdbcu_context *ctx, int h) {
    Stager &g = ctx->stager;
    if (g.in_flight[h]) {
        cudaError_t e = cudaEventSynchronize(g.done[h]);
        if (e != cudaSuccess) return e;
        g.in_flight[h] = false;
    }
    for (const PendingOut &o : g.out[h]) std::memcpy(o.dst, o.src, o.bytes);
    g.out[h].clear();
    return cudaSuccess;
}

// `bytes` (<= Stager::HALF) of bounce space that stays untouched until the stream has passed this point.
static cudaError_t stager_reserve(mdbcu_context *ctx, size_t bytes, uint8_t **out) {
    Stager &g = ctx->stager;
    if (!g.base) {
        cudaError_t e = cudaHostAlloc((void **)&g.base, 2 * Stager::HALF, cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
        for (int h = 0; h < 2; h++)
            if ((e = cudaEventCreateWithFlags(&g.done[h], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if (g.head + bytes > Stager::HALF) {
        cudaError_t e = cudaEventRecord(g.done[g.cur], ctx->stream);
        if (e != cudaSuccess) return e;
        g.in_flight[g.cur] = true;
        g.cur ^= 1;
        g.head = 0;
        if ((e = stager_retire(ctx, g.cur)) != cudaSuccess) return e;
    }
    *out = g.base + (size_t)g.cur * Stager::HALF + g.head;
    g.head += (bytes + 255) & ~(size_t)255;
    return cudaSuccess;
}

static size_t stager_piece(const mdbcu_context *ctx, size_t bytes) {
    size_t room = Stager::HALF - ctx->stager.head;
    return std::min(bytes, room >= (1u << 20) ? room : Stager::HALF);
}

static cudaError_t h2d_bytes(mdbcu_context *ctx, void *d_dst, const void *h_src, size_t bytes) {
    if (!bytes) return cudaSuccess;
    void *alias = nullptr;
    if (is_pinned(h_src, &alias)) {
        if (alias && bytes <= SMALL_COPY) return copy_by_kernel(ctx, d_dst, alias, bytes);
        return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    }
    while (bytes) {
        size_t n = stager_piece(ctx, bytes);
        uint8_t *b;
        cudaError_t e = stager_reserve(ctx, n, &b);
        if (e != cudaSuccess) return e;
        std::memcpy(b, h_src, n);
        if ((e = cudaMemcpyAsync(d_dst, b, n, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return e;
        d_dst = (uint8_t *)d_dst + n;
        h_src = (const uint8_t *)h_src + n;
        bytes -= n;
    }
    return cudaSuccess;
}

// The bytes are in `h_dst` after the next sync_ctx().
static cudaError_t d2h_bytes(mdbcu_context *ctx, void *h_dst, const void *d_src, size_t bytes) {
    if (!bytes) return cudaSuccess;
    void *alias = nullptr;
    if (is_pinned(h_dst, &alias)) {
        if (alias && bytes <= SMALL_COPY) return copy_by_kernel(ctx, alias, d_src, bytes);
        return cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    }
    while (bytes) {
        size_t n = stager_piece(ctx, bytes);
        uint8_t *b;
        cudaError_t e = stager_reserve(ctx, n, &b);
        if (e != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(b, d_src, n, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return e;
        ctx->stager.out[ctx->stager.cur].push_back(PendingOut{h_dst, b, n});
        h_dst = (uint8_t *)h_dst + n;
        d_src = (const uint8_t *)d_src + n;
        bytes -= n;
    }
    return cudaSuccess;
}

static cudaError_t sync_stream(mdbcu_context *ctx) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    Stager &g = ctx->stager;
    g.in_flight[0] = g.in_flight[1] = false;
    if (e != cudaSuccess) { // the caller's buffers may be gone by the next call: drop, never replay
        g.out[0].clear();
        g.out[1].clear();
    }
    stager_retire(ctx, g.cur ^ 1); // older half first: a copy may continue from it into the current one
    stager_retire(ctx, g.cur);
    g.head = 0;
    return e;
}

static cudaError_t pinned_acquire(mdbcu_context *ctx, size_t bytes, PinnedBlock &b) {
    std::vector<PinnedBlock> &cache = ctx->pinned_cache;
    int best = -1;
    for (size_t i = 0; i < cache.size(); i++)
        if (cache[i].cap >= bytes && (best < 0 || cache[i].cap < cache[best].cap)) best = (int)i;
    if (best >= 0 && cache[best].cap <= 2 * bytes + (1u << 20)) {
        b = cache[best];
        cache.erase(cache.begin() + best);
        return cudaSuccess;
    }
    b.cap = ((bytes + bytes / 4) | ((1u << 20) - 1)) + 1;
    return cudaHostAlloc((void **)&b.p, b.cap, cudaHostAllocPortable);
}

static void pinned_release(mdbcu_context *ctx, PinnedBlock b) {
    if (!b.p) return;
    if (ctx->pinned_cache.size() >= 8) {
        cudaFreeHost(b.p);
        return;
    }
    ctx->pinned_cache.push_back(b);
}

// ------------------------------------------------------------------------------------------------
// exclusive scan: out[i] = sum(in[0..i)), out[n] = total.  Three small kernels; the inputs here are
// per-row / per-unit counters, two to three orders of magnitude smaller than the point data.
// ------------------------------------------------------------------------------------------------

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t x, uint64_t &block_total) {
    __shared__ uint64_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint64_t warp_prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        uint64_t s = warp_sums[w];
        if (w < warp) warp_prefix += s;
        total += s;
    }
    __syncthreads();
    block_total = total;
    return warp_prefix + incl - x;
}

template <typename T> __global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const T *in, uint64_t n, uint64_t *block_sums) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK;
    uint64_t sum = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x; // coalesced
        if (i < n) sum += in[i];
    }
    uint64_t total;
    block_exclusive_scan(sum, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of block_sums[0..nb), total to block_sums[nb]
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint64_t *block_sums, uint64_t nb) {
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nb; base += SCAN_THREADS) {
        uint64_t i = base + threadIdx.x;
        uint64_t x = i < nb ? block_sums[i] : 0;
        uint64_t total;
        uint64_t excl = block_exclusive_scan(x, total);
        uint64_t carry = carry_s;
        if (i < nb) block_sums[i] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nb] = carry_s;
}

template <typename T> __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const T *in, uint64_t n, const uint64_t *block_sums, uint64_t *out) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        v[k] = i < n ? (uint64_t)in[i] : 0;
        sum += v[k];
    }
    uint64_t total;
    uint64_t run = block_sums[blockIdx.x] + block_exclusive_scan(sum, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[gridDim.x];
}

// out must hold n + 1 entries. Leaves the total in out[n] (device).
template <typename T> static int exclusive_scan(mdbcu_context *ctx, const T *in, uint64_t n, uint64_t *out) {
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(uint64_t), ctx->stream));
        return MDBCU_SUCCESS;
    }
    unsigned int nb = div_up(n, SCAN_BLOCK);
    DBuf<uint64_t> block_sums;
    CUDA_TRY(block_sums.alloc((size_t)nb + 1, ctx->stream));
    LAUNCH(ctx, k_scan_block_sums<T>, nb, SCAN_THREADS, 0, in, n, block_sums.p);
    LAUNCH(ctx, k_scan_sums, 1, SCAN_THREADS, 0, block_sums.p, (uint64_t)nb);
    LAUNCH(ctx, k_scan_apply<T>, nb, SCAN_THREADS, 0, in, n, block_sums.p, out);
    CUDA_TRY(cudaGetLastError());
    return MDBCU_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// K2: grid
// ------------------------------------------------------------------------------------------------

constexpr int TILE_THREADS = 256;
constexpr int TILE_POINTS_PER_THREAD = 8;
constexpr int TILE = TILE_THREADS * TILE_POINTS_PER_THREAD; // 2048 output points per block

__device__ __forceinline__ void report_bad(Status *status, uint64_t index) {
    atomicExch(&status->bad, 1u);
    atomicMin(&status->first_bad, (unsigned int)index); // (row and unit counts are below 2^32)
}

// ------------------------------------------------------------------------------------------------
// MacaqueV streams decoded by a whole warp (the decoder itself: mdb_macaque_warp.cuh)
//
// A MacaqueV stream (macaque_v.rs:272-323) is a serial state machine: where a value's code starts and how long it
// is depend on every code before it, so one stream cannot be split.  What a single thread per row loses on a GPU is
// (1) a memory round trip for every few bytes of the stream, (2) divergence between the lanes' three-way code
// branches, and (3) scattered 4-byte stores.  Here one WARP owns one long row: the lanes stage the row's bytes
// into shared memory with coalesced loads, all of them walk the same codes (uniform control flow, no divergence),
// lane l keeps every value whose index is l mod 32, and 32 values are stored by one coalesced instruction.
// Parallelism comes from the rows: thousands of warps are resident at once.  Rows shorter than WIDE_ROW_MIN values stay
// with the one-thread-per-row kernels.
// ------------------------------------------------------------------------------------------------

constexpr uint32_t BLOCK_ROW_MIN = 32768;           // model values from which a row is given to a block (mdbcu_context::block_row_min)
// Rows k_macaque_block takes (the warp kernels skip them): MacaqueV, no residuals (model 2 rows never carry any), long.
__device__ __forceinline__ bool macaque_block_row(const Row &r, uint32_t block_row_min) {
    if (r.model_type_id != MACAQUE_V || r.n_residuals || r.n_values < 4) return false;
    const uint64_t length = segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
    return length >= block_row_min && length <= 0xFFFFFFF0ull;
}

// One warp per wide row (the worklist's back part): the row's MacaqueV values, then its residuals if it has any.
// Timestamps of these rows are regular and were written by the tile kernel.
__global__ void __launch_bounds__(WIDE_WARPS * 32) k_grid_macaque_warp(SegmentsView v, const SegDesc *desc, const uint64_t *point_off,
                                                                        const uint32_t *worklist_back, uint32_t n_wide, uint32_t block_row_min,
                                                                        float *val_out) {
    __shared__ uint32_t stage[WIDE_WARPS][STAGE_WORDS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * WIDE_WARPS + warp;
    if (w >= n_wide) return;
    const uint64_t s = *(worklist_back - w);
    const SegDesc d = desc[s];
    const Row r = load_row(v, s);
    if (macaque_block_row(r, block_row_min)) return; // a whole block decodes it (k_macaque_block)
    const uint64_t base = point_off[s];
    const uint32_t len = (uint32_t)(point_off[s + 1] - base);
    float *row_out = val_out + base;
    const float last = warp_macaque_v_decode(
        r.values, r.n_values, d.model_len, false, 0.0f, stage[warp], lane,
        [&](uint32_t k0, float value, bool valid) {
            if (valid) row_out[k0 + lane] = value; // 32 consecutive values, one store
        },
        [&](uint32_t k0, const uint32_t(&v)[WIDE_RUN_PER_LANE]) { // a verified run of 256 `0` codes: eight consecutive values per lane
            float *o = row_out + k0 + WIDE_RUN_PER_LANE * lane;
            if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                reinterpret_cast<uint4 *>(o)[0] = make_uint4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<uint4 *>(o)[1] = make_uint4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int j = 0; j < WIDE_RUN_PER_LANE; j++) o[j] = __uint_as_float(v[j]);
            }
        });
    if (d.flags & F_HAS_RESIDUALS) { // models/mod.rs:241-249: seeded with the last gridded model value
        const uint64_t res_base = base + d.model_len;
        warp_macaque_v_decode(r.residuals, r.n_residuals - 1, len - d.model_len, true, last, stage[warp], lane,
                              [&](uint32_t k0, float value, bool valid) {
                                  if (valid) val_out[res_base + k0 + lane] = value;
                              });
    }
}

// ---- one BLOCK per very long MacaqueV row ---------------------------------------------------------------------------
// A lossless high-entropy series is ONE row of 10^6 values (BASELINE.json configs[2]: 1000 such rows), and a warp per row is
// 1000 warps on 148 SMs, each walking its stream alone.  Inside a run of `0` codes the stream is fixed width (see
// wide_run_at), so the warps of a block take consecutive stretches of 256 slots each from the row's cursor, verify
// their own flag bits and XOR their own payloads; the block then combines: everything before the first stretch with a
// failed flag is proven (by induction over the slots, stretch after stretch) and is written out, with every stretch's first
// value known from the XOR of the stretches before it.  Where a run breaks (a `11` code widens the window, a value
// repeats) warp 0 walks the next 256 codes the ordinary way and the block carries on from the state it reaches.
// SUM: the f32 additions of a row remain ONE chain in stream order (macaque_v.rs:228-264: one rounding per value); the
// block decodes 4096 values into shared memory and warp 0 adds them in order.

template <bool SUM, int ROW_WARPS>
__global__ void __launch_bounds__(ROW_WARPS * 32) k_macaque_block(SegmentsView v, const uint32_t *list, int list_step, const unsigned int *n_list_ptr,
                                                                        uint32_t n_list_max, uint32_t lane_rows_min, uint32_t block_row_min, const uint64_t *point_off,
                                                                        float *out) {
    __shared__ uint32_t stage[ROW_WARPS][STAGE_WORDS + 1];
    __shared__ uint32_t totals[ROW_WARPS];
    __shared__ uint32_t oks[ROW_WARPS];
    __shared__ float sum_buf[32]; // (SUM: a batch of 32 values decoded the ordinary way; the stretches' values go through `stage`)
    __shared__ uint64_t m_p;                         // the row's cursor: bit position, window, last value, values done
    __shared__ uint32_t m_width, m_tz, m_last, m_k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_list = n_list_ptr ? min(*n_list_ptr, n_list_max) : n_list_max;
    if (n_list >= lane_rows_min) return; // one thread per row then (k_grid_macaque_lanes / k_agg_macaque_lanes)
    for (uint32_t item = blockIdx.x; item < n_list; item += gridDim.x) {
    __syncthreads(); // (the previous row's cursor has been read by everybody)
    const uint64_t s = *(list + (int64_t)list_step * (int64_t)item);
    const Row r = load_row(v, s);
    if (!macaque_block_row(r, block_row_min)) continue; // shorter rows: k_grid_macaque_warp / k_agg_macaque_warp
    const uint32_t count = (uint32_t)segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
    float *row_out = SUM ? nullptr : out + point_off[s];
    float sum = 0.0f; // (warp 0)

    WarpMacaqueDecoder dec; // every warp: its own window onto the row's bytes; warp 0 also decodes the ordinary way with it (the
    dec.init(r.values, r.n_values, false, 0.0f, stage[warp]); // warm-up, the breaks of a run, the tail)
    WarpBitStage &bits = dec.bits;
    auto emit_batch = [&](uint32_t k0, uint32_t value, int cnt) { // warp 0: 32 values decoded the ordinary way
        if (SUM) {
            __syncwarp();
            sum_buf[lane] = __uint_as_float(value);
            __syncwarp();
            for (int j = 0; j < cnt; j++) sum = (k0 == 0 && j == 0) ? sum_buf[0] : __fadd_rn(sum, sum_buf[j]); // the first value STARTS the sum
        } else if (lane < cnt) {
            row_out[k0 + lane] = __uint_as_float(value);
        }
    };
    auto walk = [&](uint32_t k0, uint32_t n_values) { // warp 0: n_values from the cursor, 32 at a time; publishes the new cursor
        for (uint32_t k = 0; k < n_values; k += 32) {
            const int cnt = (int)min(32u, n_values - k);
            emit_batch(k0 + k, dec.batch(cnt, lane), cnt);
        }
        if (lane == 0) {
            m_p = dec.p;
            m_width = dec.width_in_force;
            m_tz = dec.trailing_zeros;
            m_last = dec.last_value;
            m_k = k0 + n_values;
        }
    };
    if (warp == 0) walk(0, min(count, WIDE_RUN)); // warm-up: the raw first value and the first windows
    __syncthreads();
    while (true) {
        const uint32_t k = m_k, width = m_width, tz = m_tz, last = m_last;
        const uint64_t p = m_p;
        if (count - k < (ROW_WARPS * WIDE_RUN)) break;
        const uint32_t stride = 1u + width;
        uint32_t x[WIDE_RUN_PER_LANE];
        const bool ok = wide_run_at(bits, p + (uint64_t)warp * WIDE_RUN * stride, width, tz, lane, x);
        // this warp's stretch of the NEXT step, if the run goes on: requested now, moved into the stage at the end of the step
        uint32_t pre[WarpBitStage::PRE_WORDS];
        uint64_t pre_w0;
        bits.prefetch_issue(p + (uint64_t)(ROW_WARPS + warp) * WIDE_RUN * stride, lane, pre, pre_w0);
        uint32_t t = x[WIDE_RUN_PER_LANE - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { // inclusive XOR scan of the lanes' totals
            const uint32_t o = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t ^= o;
        }
        if (lane == 31) {
            totals[warp] = t;
            oks[warp] = ok ? 1u : 0u;
        }
        __syncthreads();
        int good = 0; // stretches proven: all before the first one with a failed flag
        uint32_t before_warp = last, after_good = last;
        for (int w = 0; w < ROW_WARPS; w++) {
            if (!oks[w]) break;
            good = w + 1;
            if (w < warp) before_warp ^= totals[w];
            after_good ^= totals[w];
        }
        if (warp < good) {
            const uint32_t before = before_warp ^ t ^ x[WIDE_RUN_PER_LANE - 1]; // the value before this lane's first one
            const uint32_t k0 = k + (uint32_t)warp * WIDE_RUN + (uint32_t)(WIDE_RUN_PER_LANE * lane);
            if (SUM) { // the stretch's bits have been used: its values take their place in this warp's stage (reloaded next step anyway)
                __syncwarp();
#pragma unroll
                for (int j = 0; j < WIDE_RUN_PER_LANE; j++) stage[warp][WIDE_RUN_PER_LANE * lane + j] = before ^ x[j];
                bits.first_word = ~0ull;
            } else {
                float *o = row_out + k0;
                if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                    reinterpret_cast<uint4 *>(o)[0] = make_uint4(before ^ x[0], before ^ x[1], before ^ x[2], before ^ x[3]);
                    reinterpret_cast<uint4 *>(o)[1] = make_uint4(before ^ x[4], before ^ x[5], before ^ x[6], before ^ x[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < WIDE_RUN_PER_LANE; j++) o[j] = __uint_as_float(before ^ x[j]);
                }
            }
        }
        __syncthreads(); // totals / oks / cursor have been read by everybody; the values for SUM are in sum_buf
        if (warp == 0) {
            if (SUM) {
                for (int w = 0; w < good; w++) { // one addition chain, in stream order
                    const uint32_t *b = stage[w];
#pragma unroll 16
                    for (int j = 0; j < (int)WIDE_RUN; j++) sum = __fadd_rn(sum, __uint_as_float(b[j]));
                }
            }
            const uint32_t k_new = k + (uint32_t)good * WIDE_RUN;
            if (good < ROW_WARPS) { // the run broke in the next stretch: walk 256 codes the ordinary way from there
                dec.p = p + (uint64_t)good * WIDE_RUN * stride;
                dec.last_value = after_good;
                dec.width_in_force = width;
                dec.trailing_zeros = tz;
                dec.first_raw = false;
                walk(k_new, WIDE_RUN); // (count - k >= (ROW_WARPS * WIDE_RUN): these values exist)
            } else if (lane == 0) {
                m_p = p + (uint64_t)(ROW_WARPS * WIDE_RUN) * stride;
                m_last = after_good;
                m_k = k_new;
            }
        }
        __syncthreads();
        if (good == ROW_WARPS) bits.prefetch_commit(lane, pre, pre_w0); // (after a break the cursor is elsewhere: cover() loads)
    }
    if (warp == 0) { // the tail, the ordinary way
        const uint32_t k = m_k;
        dec.p = m_p;
        dec.last_value = m_last;
        dec.width_in_force = m_width;
        dec.trailing_zeros = m_tz;
        dec.first_raw = false;
        if (k < count) walk(k, count - k);
        if (SUM && lane == 0) out[s] = canonical_nan(sum);
    }
    } // rows
}

// With TENS OF THOUSANDS of long rows in a batch (100 000 series of 10 000 values) the rows themselves are parallelism
// enough, and a warp per row wastes 31 of its 32 issue slots on the serial code walk: from LANE_ROWS_MIN rows on, one
// THREAD owns a row (MacaqueVDecoder: word-wise reads through a register window, branch-free codes).  A thread's loads
// and stores walk its own row, so a warp's accesses are scattered over 32 rows: the stream is read and the values are
// written 16 bytes at a time to keep the number of sector transactions down.  Measured on 10^9 values: a thread
// decodes a code in ~1300 cycles, a warp in ~140 but at most ~27 G codes/s in total; 16 000 rows: 42 ms (threads)
// against 37 ms (warps), 100 000 rows: 7.2 ms against 44.6 ms -- the switch sits where the two cross.
constexpr uint32_t LANE_ROWS_MIN = 24576;

__global__ void __launch_bounds__(128) k_grid_macaque_lanes(SegmentsView v, const SegDesc *desc, const uint64_t *point_off,
                                                            const uint32_t *worklist_back, uint32_t n_wide, float *val_out) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_wide) return;
    const uint64_t s = *(worklist_back - w);
    const SegDesc d = desc[s];
    const Row r = load_row(v, s);
    const uint64_t base = point_off[s];
    const uint32_t len = (uint32_t)(point_off[s + 1] - base);
    MacaqueVDecoder dec;
    float last;
    // the next n values of the stream to out[0 .. n): single stores up to a 16-byte boundary, then four values per store
    auto decode_run = [&](float *out, uint32_t n) {
        uint32_t i = 0;
        for (; i < n && (reinterpret_cast<uintptr_t>(out + i) & 15); i++) out[i] = last = dec.next();
        for (; i + 4 <= n; i += 4) {
            float4 q;
            q.x = dec.next();
            q.y = dec.next();
            q.z = dec.next();
            q.w = last = dec.next();
            *reinterpret_cast<float4 *>(out + i) = q;
        }
        for (; i < n; i++) out[i] = last = dec.next();
    };
    dec.init(r.values, r.n_values, false, 0.0f);
    last = __uint_as_float(dec.last_value);
    val_out[base] = last;
    decode_run(val_out + base + 1, d.model_len - 1);
    if (d.flags & F_HAS_RESIDUALS) { // models/mod.rs:241-249: seeded with the last gridded model value
        dec.init(r.residuals, r.n_residuals - 1, true, last);
        decode_run(val_out + base + d.model_len, len - d.model_len);
    }
}

__global__ void __launch_bounds__(256) k_grid_prepare(SegmentsView v, SegDesc *desc, uint32_t *len, uint32_t *worklist, Status *status) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool serial = false, wide = false;
    if (s < v.n_segments) {
        SegDesc d;
        uint32_t n = grid_prepare_segment(v, s, d);
        desc[s] = d;
        len[s] = n;
        if (d.flags & F_MALFORMED) report_bad(status, s);
        serial = (d.flags & F_SEQUENTIAL) != 0;
        // a long MacaqueV value stream on regular timestamps: decoded by a whole warp (k_grid_macaque_warp)
        wide = serial && (d.flags & F_REGULAR) && (d.flags & F_TYPE_MASK) == MACAQUE_V && d.model_len >= WIDE_ROW_MIN;
        serial = serial && !wide;
    }
    // warp-aggregated appends: serial rows from the front of the worklist array, wide rows from its back
    const int lane = threadIdx.x & 31;
    unsigned int mask = __ballot_sync(0xffffffffu, serial);
    if (mask) {
        int leader = __ffs(mask) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&status->n_seq, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (serial) worklist[base + __popc(mask & ((1u << lane) - 1))] = (uint32_t)s;
    }
    mask = __ballot_sync(0xffffffffu, wide);
    if (mask) {
        int leader = __ffs(mask) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&status->n_wide, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (wide) worklist[v.n_segments - 1 - (base + __popc(mask & ((1u << lane) - 1)))] = (uint32_t)s;
    }
}

// tile_first[t] = the row that contains output point t * TILE
__global__ void __launch_bounds__(256) k_grid_tile_index(const uint64_t *point_off, uint64_t n_segments, uint32_t *tile_first) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segments) return;
    uint64_t a = point_off[s], b = point_off[s + 1];
    for (uint64_t t = (a + TILE - 1) / TILE; t * TILE < b; t++) tile_first[t] = (uint32_t)s;
}

// One block per tile of 2048 consecutive output points.  The rows overlapping the tile are found with
// head flags + a block-wide max-scan in shared memory (no per-point binary search), then every thread
// writes points tid, tid + 256, ... so that each warp store covers 128 B of values / 256 B of timestamps.
__global__ void __launch_bounds__(TILE_THREADS) k_grid_tile(const SegDesc *__restrict__ desc, const uint64_t *__restrict__ point_off,
                                                            const uint32_t *__restrict__ tile_first, uint64_t n_segments,
                                                            uint64_t total, int64_t *__restrict__ ts_out, float *__restrict__ val_out) {
    __shared__ uint64_t po_s[TILE + 2];      // point offsets of the rows overlapping the tile
    __shared__ uint16_t seg_of[TILE];        // local row index of every point of the tile
    __shared__ uint32_t warp_max[TILE_THREADS / 32];

    const uint64_t tile_start = (uint64_t)blockIdx.x * TILE;
    const uint64_t tile_end = min(total, tile_start + TILE);
    const uint64_t s0 = tile_first[blockIdx.x];
    const int tid = threadIdx.x;

    for (int p = tid; p < TILE; p += TILE_THREADS) seg_of[p] = 0;
    __syncthreads();
    // Stage point_off[s0 ..] until it passes the end of the tile (rows have >= 1 point, so at most
    // TILE + 1 rows overlap) and drop a head flag where each row i >= 1 starts inside the tile.
    for (int base = 0; base < TILE + 2; base += TILE_THREADS) {
        int i = base + tid;
        uint64_t x = ~0ull;
        if (i < TILE + 2) {
            if (s0 + i <= n_segments) x = point_off[s0 + i];
            po_s[i] = x;
            if (i >= 1 && x < tile_end) seg_of[x - tile_start] = (uint16_t)i;
        }
        if (!__syncthreads_or(tid == TILE_THREADS - 1 && x < tile_end)) break;
    }
    __syncthreads();
    // inclusive max-scan over seg_of: thread owns 8 consecutive entries
    uint32_t loc[TILE_POINTS_PER_THREAD];
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < TILE_POINTS_PER_THREAD; k++) {
        run = max(run, (uint32_t)seg_of[tid * TILE_POINTS_PER_THREAD + k]);
        loc[k] = run;
    }
    uint32_t incl = run;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = max(incl, y);
    }
    if (lane == 31) warp_max[warp] = incl;
    uint32_t prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev = 0;
    __syncthreads();
    for (int w = 0; w < warp; w++) prev = max(prev, warp_max[w]);
#pragma unroll
    for (int k = 0; k < TILE_POINTS_PER_THREAD; k++) seg_of[tid * TILE_POINTS_PER_THREAD + k] = (uint16_t)max(prev, loc[k]);
    __syncthreads();

    // Four consecutive points per thread: they nearly always lie in one row, so the row is looked up once and the
    // points leave in 16-byte stores (a warp store covers 512 contiguous bytes).  Quads that straddle rows, touch
    // the end of the output, reach into a row's residual part or belong to rows the serial kernels write fall back
    // to the point-wise code, as do outputs that are not 16-byte aligned.
    const bool vector_ok = ((reinterpret_cast<uintptr_t>(val_out) | reinterpret_cast<uintptr_t>(ts_out)) & 15) == 0;
#pragma unroll
    for (int k = 0; k < TILE_POINTS_PER_THREAD / 4; k++) {
        const int p = 4 * (tid + k * TILE_THREADS);
        const uint64_t gp = tile_start + p;
        if (gp >= tile_end) continue;
        const uint32_t i = seg_of[p];
        if (vector_ok && gp + 3 < tile_end && seg_of[p + 3] == i) {
            const SegDesc d = desc[s0 + i];
            if (!(d.flags & F_REGULAR)) continue; // timestamps and values of this row come from k_grid_sequential
            const uint32_t j = (uint32_t)(gp - po_s[i]);
            const int64_t t0 = d.start + (int64_t)j * d.interval;
            const int64_t t1 = t0 + d.interval, t2 = t1 + d.interval, t3 = t2 + d.interval;
            reinterpret_cast<longlong2 *>(ts_out + gp)[0] = make_longlong2(t0, t1);
            reinterpret_cast<longlong2 *>(ts_out + gp)[1] = make_longlong2(t2, t3);
            if (d.flags & F_TILE_VALUES) {
                if (j + 3 < d.model_len) {
                    float4 v;
                    if ((d.flags & F_TYPE_MASK) == PMC_MEAN) {
                        v.x = v.y = v.z = v.w = (float)d.a;                                  // pmc_mean.rs:104-108
                    } else {
                        v.x = swing_value(d.a, d.b, t0);                                     // swing.rs:304-319
                        v.y = swing_value(d.a, d.b, t1);
                        v.z = swing_value(d.a, d.b, t2);
                        v.w = swing_value(d.a, d.b, t3);
                    }
                    *reinterpret_cast<float4 *>(val_out + gp) = v;
                } else { // the model part ends inside the quad: the rest are residual values (k_grid_sequential)
                    for (int q = 0; q < 4; q++)
                        if (j + q < d.model_len)
                            val_out[gp + q] = (d.flags & F_TYPE_MASK) == PMC_MEAN ? (float)d.a : swing_value(d.a, d.b, t0 + q * d.interval);
                }
            }
            continue;
        }
        for (int q = 0; q < 4; q++) {
            if (gp + q < tile_end) {
                const uint32_t iq = seg_of[p + q];
                grid_point(desc[s0 + iq], (uint32_t)(gp + q - po_s[iq]), ts_out, val_out, gp + q);
            }
        }
    }
}


// Tiles with their rows found by BINARY SEARCH.  k_grid_tile's head flags + block-wide max-scan give every point its row in
// O(1), but cost ~40 of its 64 thread instructions per point, and every tile pays three dependent global round trips
// (first row, point offsets, descriptors) with little work to hide them behind (ncu, round 2: issue slots half used,
// long_scoreboard the top stall).  Here a block takes TILE_GROUP consecutive tiles at once: the rows overlapping them are
// tile_first[b] .. tile_first[b + TILE_GROUP] (about forty on the benchmark); their point offsets are staged in shared
// memory once, each QUAD of four consecutive points finds its row with log2(rows) probes, and a thread's
// TILE_GROUP * 2 quads have their descriptor loads and stores in flight together.
#ifndef MDB_TILE_GROUP
#define MDB_TILE_GROUP 4
#endif
constexpr int TILE_GROUP = MDB_TILE_GROUP;
__global__ void __launch_bounds__(TILE_THREADS) k_grid_tile_search(const SegDesc *__restrict__ desc, const uint64_t *__restrict__ point_off,
                                                                   const uint32_t *__restrict__ tile_first, uint64_t n_segments, uint32_t n_tiles,
                                                                   uint64_t total, int64_t *__restrict__ ts_out, float *__restrict__ val_out) {
    __shared__ uint64_t po_s[TILE + 2]; // point offsets of the rows overlapping the span, and the end of the last one
    const int tid = threadIdx.x;
    const uint32_t tile0 = blockIdx.x * TILE_GROUP;
    const bool vector_ok = ((reinterpret_cast<uintptr_t>(val_out) | reinterpret_cast<uintptr_t>(ts_out)) & 15) == 0;
    uint32_t t_lo = tile0, t_hi = min(tile0 + (uint32_t)TILE_GROUP, n_tiles); // the tiles [t_lo, t_hi) are done together
    while (t_lo < t_hi) {
        const uint64_t s0 = tile_first[t_lo];
        uint64_t s_last = t_hi < n_tiles ? (uint64_t)tile_first[t_hi] : n_segments - 1; // (may start at the next span's first point)
        if (s_last - s0 + 1 > (uint64_t)TILE + 1) { // more rows than the staging area holds (rows of a few points): one tile at a time
            t_hi = t_lo + 1;
            s_last = t_hi < n_tiles ? (uint64_t)tile_first[t_hi] : n_segments - 1; // (rows have >= 1 point: at most TILE + 1 now)
        }
        const int nr = (int)(s_last - s0) + 1; // rows s0 .. s0 + nr - 1
        const uint64_t span_start = (uint64_t)t_lo * TILE, span_end = min(total, (uint64_t)t_hi * TILE);
        __syncthreads(); // (the previous span's offsets have been read by everybody)
        for (int i = tid; i <= nr; i += TILE_THREADS) po_s[i] = point_off[s0 + i];
        __syncthreads();
        for (uint64_t gp = span_start + 4 * (uint64_t)tid; gp < span_end; gp += 4 * TILE_THREADS) {
            int i = 0, hi = nr; // largest i in [0, nr) with po_s[i] <= gp
            while (hi - i > 1) {
                const int mid = (i + hi) >> 1;
                if (po_s[mid] <= gp) i = mid;
                else hi = mid;
            }
            if (vector_ok && gp + 3 < span_end && gp + 3 < po_s[i + 1]) { // the four points lie in row i
                const SegDesc d = desc[s0 + i];
                if (!(d.flags & F_REGULAR)) continue; // timestamps and values of this row come from k_grid_sequential
                const uint32_t j = (uint32_t)(gp - po_s[i]);
                const int64_t t0 = d.start + (int64_t)j * d.interval;
                const int64_t t1 = t0 + d.interval, t2 = t1 + d.interval, t3 = t2 + d.interval;
                reinterpret_cast<longlong2 *>(ts_out + gp)[0] = make_longlong2(t0, t1);
                reinterpret_cast<longlong2 *>(ts_out + gp)[1] = make_longlong2(t2, t3);
                if (d.flags & F_TILE_VALUES) {
                    if (j + 3 < d.model_len) {
                        float4 v;
                        if ((d.flags & F_TYPE_MASK) == PMC_MEAN) {
                            v.x = v.y = v.z = v.w = (float)d.a;                                  // pmc_mean.rs:104-108
                        } else {
                            v.x = swing_value(d.a, d.b, t0);                                     // swing.rs:304-319
                            v.y = swing_value(d.a, d.b, t1);
                            v.z = swing_value(d.a, d.b, t2);
                            v.w = swing_value(d.a, d.b, t3);
                        }
                        *reinterpret_cast<float4 *>(val_out + gp) = v;
                    } else { // the model part ends inside the quad: the rest are residual values (k_grid_sequential)
                        for (int q = 0; q < 4; q++)
                            if (j + q < d.model_len)
                                val_out[gp + q] = (d.flags & F_TYPE_MASK) == PMC_MEAN ? (float)d.a : swing_value(d.a, d.b, t0 + q * d.interval);
                    }
                }
                continue;
            }
            for (int q = 0; q < 4; q++) {
                if (gp + q < span_end) {
                    while (gp + q >= po_s[i + 1]) i++; // (rows have >= 1 point: at most three steps)
                    grid_point(desc[s0 + i], (uint32_t)(gp + q - po_s[i]), ts_out, val_out, gp + q);
                }
            }
        }
        t_lo = t_hi;
        t_hi = min(tile0 + (uint32_t)TILE_GROUP, n_tiles);
    }
}

// ---- the same tile, staged in shared memory and written by the TMA engine ------------------------------------------
// k_grid_tile's stores are issued by the threads themselves: three 16-byte STG per four points, each waiting in the LSU
// pipe behind the descriptor loads of the same thread.  Here a block assembles the whole tile (2048 timestamps = 16 KiB,
// 2048 values = 8 KiB) in shared memory and ONE thread hands the two arrays to the TMA engine as bulk copies
// (cp.async.bulk.global.shared::cta, SASS UBLKCP): the stores leave the SM as full 128-byte lines without occupying
// issue slots, and the block goes on with its next tile while they drain (the kernel is persistent: a block walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... and only waits for the engine to have READ the buffer before refilling it).
// The descriptors of the rows overlapping the tile are staged in shared memory once (a tile overlaps ~10 rows on the
// benchmark) instead of being fetched per quad.  Positions of the tile that belong to the serial kernels (irregular
// rows, MacaqueV values, residual values) are written as zeros here and overwritten by those kernels, which run after
// this one on the same stream.  Needs 16-byte aligned outputs.
//
// MEASURED (B200, round 2, 10^9 points of PMC-Mean / Swing rows): k_grid_tile 2.86 ms; this kernel 3.9 ms as first written,
// 3.57 ms with the binary-search row lookup, 3.64 ms with the metadata pipeline added, 4.04 ms with the loops rolled.  ncu
// (profiles/r02_grid_tile_tma_ncu_full.txt): both kernels are instruction-issue bound (2.0e9 against 2.4e9-2.8e9 warp
// instructions per 10^9 points at IPC 2.4-2.7), neither is near the store bandwidth, and this one adds a pass through
// shared memory and two more barriers per tile (top stall: barrier, 5.4 cycles per instruction).  The plain kernel
// therefore stays the default; this one is selected with mdbcu_context_set_option("grid_tma_stores", 1) and is covered
// by the same tests.
constexpr int TILE_DESC_CACHE = 48;

__device__ __forceinline__ void bulk_store(void *gmem, const void *smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(bytes)
                 : "memory");
}

// Rows of a tile are found by a binary search over the tile's row starts (a tile overlaps ~10 rows on the benchmark, at most
// TILE + 1): k_grid_tile's head flags + block-wide max-scan cost ~40 of its 64 thread instructions per point.
//
// Software pipeline.  What a tile needs before it can compute -- its first row (tile_first), that row's and the following
// rows' point offsets, their descriptors -- is three DEPENDENT global round trips, and under several TB/s of store traffic a
// round trip takes microseconds.  They are therefore taken off the critical path: while tile k is computed, the point offsets
// and descriptors of the block's NEXT tile are copied into a second staging buffer with cp.async (LDGSTS, 8-byte pieces: the
// arrays are only 8-byte aligned at s0), and the first row of the tile after that is loaded into a register.
constexpr int TILE_PO_WINDOW = 256; // point offsets staged per tile; tiles that overlap more rows read the rest from global memory

__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(TILE_THREADS) k_grid_tile_tma(const SegDesc *__restrict__ desc, const uint64_t *__restrict__ point_off,
                                                                const uint32_t *__restrict__ tile_first, uint64_t n_segments, uint64_t total,
                                                                uint32_t n_tiles, int64_t *__restrict__ ts_out, float *__restrict__ val_out) {
    __shared__ __align__(128) int64_t ts_tile[TILE];
    __shared__ __align__(128) float val_tile[TILE];
    __shared__ uint16_t po_rel[TILE + 2];   // po_rel[i], i >= 1: start of local row i relative to the tile (< TILE); po_rel[0] = 0
    __shared__ __align__(8) uint64_t po_stage[2][TILE_PO_WINDOW];                       // point_off[s0 .. s0 + 256) of this / the next tile
    __shared__ __align__(8) uint64_t desc_stage[2][TILE_DESC_CACHE * sizeof(SegDesc) / 8]; // desc[s0 .. s0 + 48) likewise
    static_assert(sizeof(SegDesc) == 40 && TILE_DESC_CACHE * 5 <= TILE_THREADS, "one 8-byte piece of a descriptor per thread");

    const int tid = threadIdx.x;
    // stage the metadata of the tile whose first row is s0 into buffer `buf` (asynchronously; one commit group)
    auto prefetch = [&](int buf, uint64_t s0) {
        if (s0 + tid <= n_segments) cp_async_8(&po_stage[buf][tid], point_off + s0 + tid);
        if (tid < TILE_DESC_CACHE * 5 && s0 + (uint64_t)(tid / 5) < n_segments)
            cp_async_8(&desc_stage[buf][tid], reinterpret_cast<const uint64_t *>(desc + s0) + tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    uint64_t s0 = tile_first[tile];
    prefetch(0, s0);
    uint64_t s0_next = tile + gridDim.x < n_tiles ? tile_first[tile + gridDim.x] : 0;
    for (int buf = 0; tile < n_tiles; buf ^= 1) {
        const uint32_t next_tile = tile + gridDim.x;
        const uint64_t tile_start = (uint64_t)tile * TILE;
        const uint64_t tile_end = min(total, tile_start + TILE);
        const uint32_t tile_n = (uint32_t)(tile_end - tile_start);
        // this tile's metadata has landed (every thread waits for its own pieces; the barrier below publishes them) ...
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // ... and the previous tile's bulk stores have READ the tile buffers before they are refilled
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        // the next tile's metadata, and the first row of the one after it
        uint64_t s0_after = 0;
        if (next_tile < n_tiles) {
            prefetch(buf ^ 1, s0_next);
            const uint32_t after = next_tile + gridDim.x;
            if (after < n_tiles) s0_after = tile_first[after];
        }
        // rows s0 .. that start inside the tile: from the staged window, then (rarely) 256 at a time from global memory
        const uint64_t *po_w = po_stage[buf];
        uint32_t n_rows = 1;
        for (int base = 0; base < TILE + 2; base += TILE_THREADS) {
            const int i = base + tid;
            uint64_t x = ~0ull;
            if (i < TILE + 2 && s0 + i <= n_segments) x = base == 0 ? po_w[tid] : point_off[s0 + i];
            const bool inside = i >= 1 && x < tile_end;
            if (i == 0) po_rel[0] = 0;
            if (inside) po_rel[i] = (uint16_t)(x - tile_start);
            const int cnt = __syncthreads_count(inside);
            n_rows += (uint32_t)cnt;
            if (cnt < TILE_THREADS - (base == 0 ? 1 : 0)) break; // (row starts increase: the first entry outside ends the list)
        }
        const uint64_t po0 = po_w[0]; // first point of local row 0 (it may start before the tile)
        const SegDesc *desc_w = reinterpret_cast<const SegDesc *>(desc_stage[buf]);
        const bool staged = n_rows <= (uint32_t)TILE_DESC_CACHE; // every descriptor of the tile is in shared memory (the usual case)

        // Four consecutive points per thread and pass; lane l of a warp owns quads l, l + 32 of the warp's 256 points, so that a
        // warp's 16-byte shared stores are contiguous.  The row of the first quad is found by binary search, the row of the
        // second by walking on from it.
        uint32_t row = 0;
#pragma unroll 1
        for (int k = 0; k < TILE_POINTS_PER_THREAD / 4; k++) {
            const uint32_t p = 4u * (uint32_t)(tid + k * TILE_THREADS);
            if (p >= tile_n) break;
            if (k == 0) {
                uint32_t lo = 0, hi = n_rows; // the last local row that starts at or before p
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (po_rel[mid] <= p) lo = mid;
                    else hi = mid;
                }
                row = lo;
            } else {
                while (row + 1 < n_rows && po_rel[row + 1] <= p) row++;
            }
            const uint32_t next_start = row + 1 < n_rows ? po_rel[row + 1] : 0xFFFFu;
            longlong2 t01 = make_longlong2(0, 0), t23 = make_longlong2(0, 0);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p + 3 < next_start && p + 3 < tile_n) { // the quad lies in one row
                const SegDesc d = staged ? desc_w[row] : desc[s0 + row];
                if (d.flags & F_REGULAR) {
                    const uint32_t j = row == 0 ? (uint32_t)(tile_start + p - po0) : p - po_rel[row];
                    const int64_t t0 = d.start + (int64_t)j * d.interval;
                    const int64_t t1 = t0 + d.interval, t2 = t1 + d.interval, t3 = t2 + d.interval;
                    t01 = make_longlong2(t0, t1);
                    t23 = make_longlong2(t2, t3);
                    if (d.flags & F_TILE_VALUES) {
                        if ((d.flags & F_TYPE_MASK) == PMC_MEAN) {
                            v.x = v.y = v.z = v.w = (float)d.a;                       // pmc_mean.rs:104-108
                        } else {
                            v.x = swing_value(d.a, d.b, t0);                          // swing.rs:304-319
                            v.y = swing_value(d.a, d.b, t1);
                            v.z = swing_value(d.a, d.b, t2);
                            v.w = swing_value(d.a, d.b, t3);
                        }
                        if (j + 3 >= d.model_len) { // the model part ends inside the quad: the rest are residual values (k_grid_sequential)
                            if (j + 0 >= d.model_len) v.x = 0.f;
                            if (j + 1 >= d.model_len) v.y = 0.f;
                            if (j + 2 >= d.model_len) v.z = 0.f;
                            v.w = 0.f;
                        }
                    }
                }
            } else { // the quad straddles rows or the end of the output: point by point
                int64_t tt[4] = {0, 0, 0, 0};
                float vv[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t iq = row;
#pragma unroll 1
                for (uint32_t q = 0; q < 4 && p + q < tile_n; q++) {
                    while (iq + 1 < n_rows && po_rel[iq + 1] <= p + q) iq++;
                    const SegDesc d = staged ? desc_w[iq] : desc[s0 + iq];
                    if (!(d.flags & F_REGULAR)) continue;
                    const uint32_t j = iq == 0 ? (uint32_t)(tile_start + p + q - po0) : p + q - po_rel[iq];
                    const int64_t t = d.start + (int64_t)j * d.interval;
                    float x = 0.f;
                    if ((d.flags & F_TILE_VALUES) && j < d.model_len) x = (d.flags & F_TYPE_MASK) == PMC_MEAN ? (float)d.a : swing_value(d.a, d.b, t);
                    // (q is a loop variable: select the slot without indexing a register array dynamically)
                    if (q == 0) { tt[0] = t; vv[0] = x; }
                    if (q == 1) { tt[1] = t; vv[1] = x; }
                    if (q == 2) { tt[2] = t; vv[2] = x; }
                    if (q == 3) { tt[3] = t; vv[3] = x; }
                }
                t01 = make_longlong2(tt[0], tt[1]);
                t23 = make_longlong2(tt[2], tt[3]);
                v = make_float4(vv[0], vv[1], vv[2], vv[3]);
            }
            reinterpret_cast<longlong2 *>(ts_tile + p)[0] = t01;
            reinterpret_cast<longlong2 *>(ts_tile + p)[1] = t23;
            *reinterpret_cast<float4 *>(val_tile + p) = v;
        }
        // shared-memory writes of the generic proxy, made visible to the async proxy (the TMA engine) before the copy is issued
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t bulk_n = tile_n & ~3u; // sizes are multiples of 16 bytes; the last tile's tail (<= 3 points) is stored directly
            if (bulk_n) {
                bulk_store(ts_out + tile_start, ts_tile, bulk_n * 8u);
                bulk_store(val_out + tile_start, val_tile, bulk_n * 4u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            for (uint32_t q = bulk_n; q < tile_n; q++) {
                ts_out[tile_start + q] = ts_tile[q];
                val_out[tile_start + q] = val_tile[q];
            }
        }
        tile = next_tile;
        s0 = s0_next;
        s0_next = s0_after;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // the copies are complete before the block retires
}

__global__ void __launch_bounds__(128) k_grid_sequential(SegmentsView v, const SegDesc *desc, const uint64_t *point_off, const uint32_t *worklist,
                                                         uint32_t n_work, int64_t *ts_out, float *val_out) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_work) return;
    uint64_t s = worklist[w];
    uint64_t base = point_off[s];
    grid_sequential_segment(v, s, desc[s], base, (uint32_t)(point_off[s + 1] - base), ts_out, val_out);
}

// ------------------------------------------------------------------------------------------------
// K3: aggregates
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) k_agg_segments(SegmentsView v, uint64_t *seg_count, float *seg_sum, uint32_t *wide_list, Status *status) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool wide = false;
    if (s < v.n_segments) {
        uint64_t c;
        float sum;
        if (!aggregate_segment(v, s, c, sum, WIDE_ROW_MIN, &wide)) report_bad(status, s);
        if (seg_count) seg_count[s] = c;
        if (!wide) seg_sum[s] = sum;
    }
    // long MacaqueV rows: their SUM is computed by k_agg_macaque_warp
    const unsigned int mask = __ballot_sync(0xffffffffu, wide);
    if (mask) {
        const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&status->n_wide, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (wide) wide_list[base + __popc(mask & ((1u << lane) - 1))] = (uint32_t)s;
    }
}

// One warp per long MacaqueV row: the f32 sum in stream order (macaque_v.rs:220-265), as aggregate_segment computes it.
// The 32 values of a batch are decoded in parallel; their additions stay a serial chain (one rounding per value).
__global__ void __launch_bounds__(WIDE_WARPS * 32) k_agg_macaque_warp(SegmentsView v, const uint32_t *wide_list, const unsigned int *n_wide_ptr,
                                                                       uint32_t lane_rows_min, uint32_t block_row_min, float *seg_sum) {
    __shared__ uint32_t stage[WIDE_WARPS][STAGE_WORDS + 1];
    __shared__ float batch[WIDE_WARPS][32];
    __shared__ __align__(16) float wide_batch[WIDE_WARPS][WIDE_RUN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_wide = *n_wide_ptr;
    if (n_wide >= lane_rows_min) return; // k_agg_macaque_lanes has them
    for (uint32_t w = blockIdx.x * WIDE_WARPS + warp; w < n_wide; w += gridDim.x * WIDE_WARPS) {
        const uint64_t s = wide_list[w];
        const Row r = load_row(v, s);
        if (macaque_block_row(r, block_row_min)) continue; // a whole block decodes it (k_macaque_block)
        const uint64_t res_len = r.n_residuals ? r.residuals[r.n_residuals - 1] : 0;
        const uint64_t length = segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
        const uint32_t model_length = (uint32_t)(length - res_len);
        float sum = 0.0f;
        bool first = true; // the first value STARTS the model's sum (macaque_v.rs:228-233); a seeded stream starts from 0
        auto add_in_order = [&](uint32_t k0, float value, bool, uint32_t total) {
            __syncwarp();
            batch[warp][lane] = value;
            __syncwarp();
            const int cnt = (int)min(32u, total - k0);
            for (int j = 0; j < cnt; j++) {
                const float x = batch[warp][j];
                sum = first ? x : __fadd_rn(sum, x);
                first = false;
            }
        };
        warp_macaque_v_decode(
            r.values, r.n_values, model_length, false, 0.0f, stage[warp], lane,
            [&](uint32_t k0, float value, bool valid) { add_in_order(k0, value, valid, model_length); },
            [&](uint32_t, const uint32_t(&v)[WIDE_RUN_PER_LANE]) { // 256 values at once: still ONE addition chain in stream order
                __syncwarp();
                uint4 *slot = reinterpret_cast<uint4 *>(&wide_batch[warp][WIDE_RUN_PER_LANE * lane]);
                slot[0] = make_uint4(v[0], v[1], v[2], v[3]);
                slot[1] = make_uint4(v[4], v[5], v[6], v[7]);
                __syncwarp();
                const float4 *b4 = reinterpret_cast<const float4 *>(wide_batch[warp]);
#pragma unroll 8
                for (int j = 0; j < (int)WIDE_RUN / 4; j++) { // (never the stream's first value: wide runs start after it)
                    const float4 q = b4[j];
                    sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, q.x), q.y), q.z), q.w);
                }
            });
        if (r.n_residuals) { // models/mod.rs:173-183: seeded with the "last value" a MacaqueV model reports there, NaN
            const float model_sum = sum;
            sum = 0.0f;
            first = false;
            warp_macaque_v_decode(r.residuals, r.n_residuals - 1, (uint32_t)res_len, true, __uint_as_float(0x7fc00000u), stage[warp], lane,
                                  [&](uint32_t k0, float value, bool valid) { add_in_order(k0, value, valid, (uint32_t)res_len); });
            sum = __fadd_rn(model_sum, sum);
        }
        if (lane == 0) seg_sum[s] = canonical_nan(sum);
    }
}

// The same rows with one thread per row when the batch holds at least lane_rows_min of them (see k_grid_macaque_lanes);
// the count is only known on the device here, so both kernels are launched and one of them returns at once.
__global__ void __launch_bounds__(128) k_agg_macaque_lanes(SegmentsView v, const uint32_t *wide_list, const unsigned int *n_wide_ptr,
                                                           uint32_t lane_rows_min, float *seg_sum) {
    const uint32_t n_wide = *n_wide_ptr;
    if (n_wide < lane_rows_min) return;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_wide; w += gridDim.x * blockDim.x) {
        const uint64_t s = wide_list[w];
        uint64_t count;
        float sum;
        aggregate_segment(v, s, count, sum); // well-formed: k_agg_segments checked the row before deferring it
        seg_sum[s] = sum;
    }
}

// ---- the fused aggregate: no per-row round trip through global memory --------------------------------------------
// k_agg_find_wide  one thread per row: only MacaqueV rows are looked at further (1 byte per row otherwise); long ones are
//                  listed for the warp / lane decoders, which leave their f32 sums in seg_sum (sparse)
// k_agg_groups     block (part, group): every thread parses rows part_lo + tid, + 256, ... of its slice (coalesced column
//                  reads), computes COUNT and SUM of each from the model in registers (aggregate_segment) and folds them;
//                  a fixed shuffle tree then combines the block.  Nothing per row is written.
// The shape of the reduction depends only on (rows of the group, parts), and parts only on the batch: results do not
// depend on the device.  COUNT / MIN / MAX are exact; the f64 SUM is a fixed tree over f32 row sums (the reference adds
// them row by row, model_simple_aggregates.rs:481-511: 1e-12 relative, see modelardb_cuda.h).
__global__ void __launch_bounds__(256) k_agg_find_wide(SegmentsView v, uint32_t *wide_list, Status *status) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool wide = false;
    if (s < v.n_segments && v.model_type_id[s] == MACAQUE_V) {
        uint64_t c;
        float sum;
        aggregate_segment(v, s, c, sum, WIDE_ROW_MIN, &wide, /*count_only=*/true);
    }
    const unsigned int mask = __ballot_sync(0xffffffffu, wide);
    if (mask) {
        const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(&status->n_wide, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (wide) wide_list[base + __popc(mask & ((1u << lane) - 1))] = (uint32_t)s;
    }
}

constexpr int AGG_THREADS = 256;

// In-order tree reduction of a block's per-thread partials (thread t holds rows EARLIER than t + 1).
__device__ __forceinline__ GroupAgg block_reduce_in_order(GroupAgg a) {
    __shared__ GroupAgg warp_part[AGG_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        GroupAgg o;
        o.count = __shfl_down_sync(0xffffffffu, a.count, d);
        o.min = __shfl_down_sync(0xffffffffu, a.min, d);
        o.max = __shfl_down_sync(0xffffffffu, a.max, d);
        o.sum = __shfl_down_sync(0xffffffffu, a.sum, d);
        if (lane + d < 32) a = group_agg_combine(a, o);
    }
    if (lane == 0) warp_part[warp] = a;
    __syncthreads();
    GroupAgg r = warp_part[0];
    if (threadIdx.x == 0)
        for (int w = 1; w < AGG_THREADS / 32; w++) r = group_agg_combine(r, warp_part[w]);
    __syncthreads();
    return r; // valid in thread 0
}

#ifndef MDB_AGG_SLICE
#define MDB_AGG_SLICE 2048
#endif
// Block (part, group) of the fused aggregate: see k_agg_find_wide above.
__global__ void __launch_bounds__(AGG_THREADS) k_agg_groups(SegmentsView v, const uint64_t *group_off, uint64_t n_groups, uint32_t parts,
                                                            const float *wide_sum, GroupAgg *partial, Status *status) {
    for (uint64_t g = blockIdx.y; g < n_groups; g += gridDim.y) {
        uint64_t lo = group_off ? group_off[g] : 0, hi = group_off ? group_off[g + 1] : v.n_segments;
        hi = min(hi, v.n_segments); // never read past the batch, whatever the caller passed
        lo = min(lo, hi);
        const uint64_t rows = hi - lo, part = blockIdx.x;
        const uint64_t plo = lo + rows * part / parts, phi = lo + rows * (part + 1) / parts;
        GroupAgg a = group_agg_identity();
        for (uint64_t s = plo + threadIdx.x; s < phi; s += AGG_THREADS) {
            GroupAgg row;
            uint64_t c;
            float sum;
            bool wide = false;
            if (!aggregate_segment(v, s, c, sum, WIDE_ROW_MIN, &wide)) report_bad(status, s);
            if (wide) sum = wide_sum[s]; // a long MacaqueV row: decoded by k_agg_macaque_warp / _lanes before this kernel
            row.count = (int64_t)c;
            row.min = v.min_value[s];
            row.max = v.max_value[s];
            row.sum = (double)sum;
            a = group_agg_combine(a, row);
        }
        a = block_reduce_in_order(a);
        if (threadIdx.x == 0) partial[g * parts + part] = a;
    }
}

// The same with one WARP per group, for batches of many small groups (GROUP BY series over 100 000 short series): lane l folds
// rows lo + l, lo + l + 32, ...; a fixed shuffle tree combines the lanes.
__global__ void __launch_bounds__(AGG_THREADS) k_agg_groups_warp(SegmentsView v, const uint64_t *group_off, uint64_t n_groups, const float *wide_sum,
                                                                 int64_t *count, float *mn, float *mx, double *sum, Status *status) {
    const int lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * (AGG_THREADS / 32) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    uint64_t lo = group_off ? group_off[g] : 0, hi = group_off ? group_off[g + 1] : v.n_segments;
    hi = min(hi, v.n_segments);
    lo = min(lo, hi);
    GroupAgg a = group_agg_identity();
    for (uint64_t s = lo + lane; s < hi; s += 32) {
        GroupAgg row;
        uint64_t c;
        float rs;
        bool wide = false;
        if (!aggregate_segment(v, s, c, rs, WIDE_ROW_MIN, &wide)) report_bad(status, s);
        if (wide) rs = wide_sum[s];
        row.count = (int64_t)c;
        row.min = v.min_value[s];
        row.max = v.max_value[s];
        row.sum = (double)rs;
        a = group_agg_combine(a, row);
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        GroupAgg o;
        o.count = __shfl_down_sync(0xffffffffu, a.count, d);
        o.min = __shfl_down_sync(0xffffffffu, a.min, d);
        o.max = __shfl_down_sync(0xffffffffu, a.max, d);
        o.sum = __shfl_down_sync(0xffffffffu, a.sum, d);
        if (lane + d < 32) a = group_agg_combine(a, o);
    }
    if (lane == 0) {
        count[g] = a.count;
        mn[g] = a.min;
        mx[g] = a.max;
        sum[g] = a.sum;
    }
}

// Block (part, group): folds the part-th contiguous slice of the group's rows.
__global__ void __launch_bounds__(AGG_THREADS) k_agg_partial(const uint64_t *group_off, uint64_t n_groups, uint64_t n_rows_all, uint32_t parts,
                                                             const uint64_t *seg_count, const float *seg_sum, const float *min_value,
                                                             const float *max_value, GroupAgg *partial) {
    for (uint64_t g = blockIdx.y; g < n_groups; g += gridDim.y) {
        uint64_t lo = group_off ? group_off[g] : 0, hi = group_off ? group_off[g + 1] : n_rows_all;
        hi = min(hi, n_rows_all); // never read past the batch, whatever the caller passed
        lo = min(lo, hi);
        uint64_t rows = hi - lo;
        uint64_t part = blockIdx.x;
        uint64_t plo = lo + rows * part / parts, phi = lo + rows * (part + 1) / parts;
        uint64_t prow = phi - plo;
        // thread t folds the t-th contiguous chunk, in row order
        uint64_t tlo = plo + prow * threadIdx.x / AGG_THREADS, thi = plo + prow * (threadIdx.x + 1) / AGG_THREADS;
        GroupAgg a = group_agg_identity();
        for (uint64_t s = tlo; s < thi; s++) {
            GroupAgg row;
            row.count = (int64_t)seg_count[s];
            row.min = min_value[s];
            row.max = max_value[s];
            row.sum = (double)seg_sum[s];
            a = group_agg_combine(a, row);
        }
        a = block_reduce_in_order(a);
        if (threadIdx.x == 0) partial[g * parts + part] = a;
    }
}

__global__ void __launch_bounds__(AGG_THREADS) k_agg_final(const GroupAgg *partial, uint64_t n_groups, uint32_t parts, int64_t *count, float *mn,
                                                           float *mx, double *sum) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    GroupAgg a = partial[g * parts];
    for (uint32_t p = 1; p < parts; p++) a = group_agg_combine(a, partial[g * parts + p]);
    count[g] = a.count;
    mn[g] = a.min;
    mx[g] = a.max;
    sum[g] = a.sum;
}

// ------------------------------------------------------------------------------------------------
// K1: compress
// ------------------------------------------------------------------------------------------------

// (the compress kernels and mdbcu_compress live in mdb_compress_api.inl, included at the end of this file)

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

struct mdbcu_segments {
    mdbcu_context *ctx = nullptr;
    uint64_t n_segments = 0, n_units = 0;
    // device columns
    int8_t *model_type_id = nullptr;
    int64_t *start_time = nullptr, *end_time = nullptr;
    float *min_value = nullptr, *max_value = nullptr;
    uint64_t *ts_off = nullptr, *val_off = nullptr, *res_off = nullptr, *unit_seg_off = nullptr;
    uint8_t *ts_data = nullptr, *val_data = nullptr, *res_data = nullptr;
    uint64_t ts_bytes = 0, val_bytes = 0, res_bytes = 0;
    // lazily made host copy: every column in one pinned block, so the copy runs at link speed and the
    // views can be handed back to grid / aggregate without being bounced
    bool have_host = false;
    PinnedBlock host_block;
    mdbcu_segments_view host_view{};
    const uint64_t *h_unit_seg_off = nullptr;
};

static SegmentsView to_device_view(const mdbcu_segments_view *v) {
    SegmentsView d;
    d.n_segments = v->n_segments;
    d.model_type_id = v->model_type_id;
    d.start_time = v->start_time;
    d.end_time = v->end_time;
    d.min_value = v->min_value;
    d.max_value = v->max_value;
    d.timestamps_off = v->timestamps_off;
    d.timestamps_data = v->timestamps_data;
    d.values_off = v->values_off;
    d.values_data = v->values_data;
    d.residuals_off = v->residuals_off;
    d.residuals_data = v->residuals_data;
    return d;
}

// Device copy of a HOST-space segment batch.
struct StagedSegments {
    DBuf<int8_t> model_type_id;
    DBuf<int64_t> start_time, end_time;
    DBuf<float> min_value, max_value;
    DBuf<uint64_t> ts_off, val_off, res_off;
    DBuf<uint8_t> ts_data, val_data, res_data;
    SegmentsView view;
};

template <typename T> static cudaError_t upload(mdbcu_context *ctx, DBuf<T> &dst, const T *src, size_t n) {
    cudaError_t e = dst.alloc(n, ctx->stream);
    if (e != cudaSuccess) return e;
    return h2d_bytes(ctx, dst.p, src, n * sizeof(T));
}

static int stage_segments(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *v, StagedSegments &st) {
    if (!v) return fail("segments view is null");
    uint64_t S = v->n_segments;
    if (S > 0xFFFFFFF0ull) return fail("more than 2^32 segment rows in one batch");
    if (space == MDBCU_DEVICE) {
        st.view = to_device_view(v);
        return MDBCU_SUCCESS;
    }
    if (S && (!v->timestamps_off || !v->values_off || !v->residuals_off)) return fail("offset column is null");
    uint64_t tb = S ? v->timestamps_off[S] : 0, vb = S ? v->values_off[S] : 0, rb = S ? v->residuals_off[S] : 0;
    CUDA_TRY(upload(ctx, st.model_type_id, v->model_type_id, S));
    CUDA_TRY(upload(ctx, st.start_time, v->start_time, S));
    CUDA_TRY(upload(ctx, st.end_time, v->end_time, S));
    CUDA_TRY(upload(ctx, st.min_value, v->min_value, S));
    CUDA_TRY(upload(ctx, st.max_value, v->max_value, S));
    CUDA_TRY(upload(ctx, st.ts_off, v->timestamps_off, S ? S + 1 : 0));
    CUDA_TRY(upload(ctx, st.val_off, v->values_off, S ? S + 1 : 0));
    CUDA_TRY(upload(ctx, st.res_off, v->residuals_off, S ? S + 1 : 0));
    CUDA_TRY(upload(ctx, st.ts_data, v->timestamps_data, tb));
    CUDA_TRY(upload(ctx, st.val_data, v->values_data, vb));
    CUDA_TRY(upload(ctx, st.res_data, v->residuals_data, rb));
    st.view.n_segments = S;
    st.view.model_type_id = st.model_type_id.p;
    st.view.start_time = st.start_time.p;
    st.view.end_time = st.end_time.p;
    st.view.min_value = st.min_value.p;
    st.view.max_value = st.max_value.p;
    st.view.timestamps_off = st.ts_off.p;
    st.view.timestamps_data = st.ts_data.p;
    st.view.values_off = st.val_off.p;
    st.view.values_data = st.val_data.p;
    st.view.residuals_off = st.res_off.p;
    st.view.residuals_data = st.res_data.p;
    return MDBCU_SUCCESS;
}

static int check_ctx(mdbcu_context *ctx) {
    if (!ctx) return fail("context is null");
    CUDA_TRY(cudaSetDevice(ctx->device));
    Stager &g = ctx->stager;
    if (g.head || g.in_flight[0] || g.in_flight[1] || !g.out[0].empty() || !g.out[1].empty()) {
        // an earlier call failed midway: none of its bounced copies may land in memory the caller has since reused
        cudaStreamSynchronize(ctx->stream);
        g.out[0].clear();
        g.out[1].clear();
        g.in_flight[0] = g.in_flight[1] = false;
        g.head = 0;
    }
    return MDBCU_SUCCESS;
}

// Waits for the stream (and with it for every transfer queued so far).
static int read_status(mdbcu_context *ctx, const Status *d_status, Status &h, const char *what) {
    static_assert(STAGE_WORDS * 32 <= 0x8000, "payload positions are packed into 15 bits");
    static_assert(sizeof(Status) == 16, "Status is posted as two words");
    CUDA_TRY(post(ctx, SLOT_STATUS, d_status, 2));
    CUDA_TRY(sync_stream(ctx));
    std::memcpy(&h, ctx->mailbox + SLOT_STATUS, sizeof(Status));
    if (h.bad) return fail(std::string("malformed ") + what + " " + std::to_string(h.first_bad));
    return MDBCU_SUCCESS;
}

static int new_status(mdbcu_context *ctx, DBuf<Status> &st) {
    CUDA_TRY(st.alloc(1, ctx->stream));
    // bad = n_seq = n_wide = 0, first_bad = ~0: two memsets, no copy engine involved
    CUDA_TRY(cudaMemsetAsync(st.p, 0, 16, ctx->stream));
    CUDA_TRY(cudaMemsetAsync((uint8_t *)st.p + 8, 0xFF, 4, ctx->stream));
    return MDBCU_SUCCESS;
}

// Warps per row of k_macaque_block: more warps finish a row sooner, fewer warps keep more rows resident (mdbcu_context::block_row_warps).
template <bool SUM>
static void launch_macaque_block(mdbcu_context *ctx, unsigned int blocks, SegmentsView v, const uint32_t *list, int step, const unsigned int *n_ptr,
                                 uint32_t n_max, const uint64_t *point_off, float *out) {
    switch (ctx->block_row_warps) {
    case 16: LAUNCH(ctx, (k_macaque_block<SUM, 16>), blocks, 16 * 32, 0, v, list, step, n_ptr, n_max, ctx->lane_rows_min, ctx->block_row_min, point_off, out); break;
    case 8: LAUNCH(ctx, (k_macaque_block<SUM, 8>), blocks, 8 * 32, 0, v, list, step, n_ptr, n_max, ctx->lane_rows_min, ctx->block_row_min, point_off, out); break;
    case 2: LAUNCH(ctx, (k_macaque_block<SUM, 2>), blocks, 2 * 32, 0, v, list, step, n_ptr, n_max, ctx->lane_rows_min, ctx->block_row_min, point_off, out); break;
    default: LAUNCH(ctx, (k_macaque_block<SUM, 4>), blocks, 4 * 32, 0, v, list, step, n_ptr, n_max, ctx->lane_rows_min, ctx->block_row_min, point_off, out); break;
    }
}

extern "C" {

const char *mdbcu_last_error(void) { return g_last_error.c_str(); }

const char *mdbcu_version(void) { return "0.1.0"; }

int mdbcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mdbcu_context_create(int device, mdbcu_context **out) {
    if (!out) return fail("out is null");
    *out = nullptr;
    int n = mdbcu_device_count();
    if (n == 0) return fail("no CUDA device: libmodelardb_cuda has no CPU implementation");
    if (device < 0 || device >= n) return fail("device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    mdbcu_context *ctx = new mdbcu_context();
    ctx->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    ctx->own_stream = true;
    e = cudaHostAlloc((void **)&ctx->mailbox, MAILBOX_WORDS * sizeof(uint64_t), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&ctx->mailbox_dev, ctx->mailbox, 0);
    if (e != cudaSuccess) {
        mdbcu_context_destroy(ctx);
        return fail(std::string("mailbox allocation: ") + cudaGetErrorString(e));
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    ctx->block_row_min = BLOCK_ROW_MIN;
    ctx->lane_rows_min = LANE_ROWS_MIN; // (tuning and tests change it with mdbcu_context_set_option: the library reads no environment variables)
    // keep freed blocks in the pool: steady-state calls then never reach the driver allocator
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    *out = ctx;
    return MDBCU_SUCCESS;
}

void mdbcu_context_destroy(mdbcu_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->aux_stream) {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
    }
    if (ctx->aux_ready) cudaEventDestroy(ctx->aux_ready);
    if (ctx->aux_done) cudaEventDestroy(ctx->aux_done);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    if (ctx->stager.base) cudaFreeHost(ctx->stager.base);
    for (cudaEvent_t ev : ctx->stager.done)
        if (ev) cudaEventDestroy(ev);
    for (PinnedBlock &b : ctx->pinned_cache) cudaFreeHost(b.p);
    for (cudaEvent_t ev : ctx->event_pool) cudaEventDestroy(ev);
    delete ctx;
}

int mdbcu_context_set_stream(mdbcu_context *ctx, void *cuda_stream) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return MDBCU_SUCCESS;
}

void *mdbcu_context_stream(mdbcu_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

uint64_t mdbcu_context_launch_count(const mdbcu_context *ctx) { return ctx ? ctx->launches : 0; }

int mdbcu_context_set_profiling(mdbcu_context *ctx, int enabled) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->resolve_timings();
    ctx->profiling = enabled != 0;
    if (enabled) ctx->stats.clear();
    return MDBCU_SUCCESS;
}

int mdbcu_context_kernel_stat(mdbcu_context *ctx, uint32_t index, const char **name, double *total_ms, uint64_t *launches) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->resolve_timings();
    if (index >= ctx->stats.size()) return fail("kernel stat index out of range");
    if (name) *name = ctx->stats[index].name.c_str();
    if (total_ms) *total_ms = ctx->stats[index].ms;
    if (launches) *launches = ctx->stats[index].launches;
    return MDBCU_SUCCESS;
}

// ---- K2 ------------------------------------------------------------------------------------------

// Shared front half of grid_count / grid: descriptors, point offsets, worklist. Leaves h_status/total on host.
struct GridPlan {
    DBuf<SegDesc> desc;
    DBuf<uint32_t> len, worklist;
    DBuf<uint64_t> point_off;
    DBuf<Status> status;
    Status h_status;
    uint64_t total = 0;
};

static int grid_plan(mdbcu_context *ctx, const SegmentsView &v, GridPlan &pl) {
    uint64_t S = v.n_segments;
    cudaStream_t s = ctx->stream;
    CUDA_TRY(pl.desc.alloc(S, s));
    CUDA_TRY(pl.len.alloc(S, s));
    CUDA_TRY(pl.worklist.alloc(S, s));
    CUDA_TRY(pl.point_off.alloc(S + 1, s));
    if (new_status(ctx, pl.status)) return MDBCU_FAILURE;
    if (S) LAUNCH(ctx, k_grid_prepare, div_up(S, 256), 256, 0, v, pl.desc.p, pl.len.p, pl.worklist.p, pl.status.p);
    if (exclusive_scan<uint32_t>(ctx, pl.len.p, S, pl.point_off.p)) return MDBCU_FAILURE;
    CUDA_TRY(post(ctx, 0, pl.point_off.p + S, 1));
    if (read_status(ctx, pl.status.p, pl.h_status, "segment row")) return MDBCU_FAILURE;
    pl.total = ctx->mailbox[0];
    return MDBCU_SUCCESS;
}

int mdbcu_grid_count(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments, uint64_t *point_off, uint64_t *total) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    StagedSegments st;
    if (stage_segments(ctx, space, segments, st)) return MDBCU_FAILURE;
    GridPlan pl;
    if (grid_plan(ctx, st.view, pl)) return MDBCU_FAILURE;
    if (point_off) {
        const size_t bytes = (st.view.n_segments + 1) * sizeof(uint64_t);
        if (space == MDBCU_HOST)
            CUDA_TRY(d2h_bytes(ctx, point_off, pl.point_off.p, bytes));
        else
            CUDA_TRY(cudaMemcpyAsync(point_off, pl.point_off.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(sync_stream(ctx));
    }
    if (total) *total = pl.total;
    return MDBCU_SUCCESS;
}

// The launches of grid(): every point of every row of `v` (device memory) into d_ts / d_val (device memory, pl.total entries).
static int grid_on_device(mdbcu_context *ctx, const SegmentsView &v, GridPlan &pl, int64_t *d_ts, float *d_val) {
    cudaStream_t s = ctx->stream;
    const uint64_t S = v.n_segments;
    unsigned int n_tiles = div_up(pl.total, TILE);
    DBuf<uint32_t> tile_first;
    CUDA_TRY(tile_first.alloc(n_tiles, s));
    LAUNCH(ctx, k_grid_tile_index, div_up(S, 256), 256, 0, pl.point_off.p, S, tile_first.p);
    if ((((uintptr_t)d_ts | (uintptr_t)d_val) & 15) == 0 && ctx->grid_tma_stores) {
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_tile_tma, TILE_THREADS, 0));
        const unsigned int blocks = std::min<unsigned int>(n_tiles, (unsigned int)(ctx->sm_count * std::max(per_sm, 1)));
        LAUNCH(ctx, k_grid_tile_tma, blocks, TILE_THREADS, 0, pl.desc.p, pl.point_off.p, tile_first.p, S, pl.total, n_tiles, d_ts, d_val);
    } else if (ctx->grid_tile_scan) {
        LAUNCH(ctx, k_grid_tile, n_tiles, TILE_THREADS, 0, pl.desc.p, pl.point_off.p, tile_first.p, S, pl.total, d_ts, d_val);
    } else {
        LAUNCH(ctx, k_grid_tile_search, div_up(n_tiles, TILE_GROUP), TILE_THREADS, 0, pl.desc.p, pl.point_off.p, tile_first.p, S, n_tiles, pl.total, d_ts,
               d_val);
    }
    if (pl.h_status.n_seq)
        LAUNCH(ctx, k_grid_sequential, div_up(pl.h_status.n_seq, 128), 128, 0, v, pl.desc.p, pl.point_off.p, pl.worklist.p,
               (uint32_t)pl.h_status.n_seq, d_ts, d_val);
    if (pl.h_status.n_wide >= ctx->lane_rows_min)
        LAUNCH(ctx, k_grid_macaque_lanes, div_up(pl.h_status.n_wide, 128), 128, 0, v, pl.desc.p, pl.point_off.p, pl.worklist.p + (S - 1),
               (uint32_t)pl.h_status.n_wide, d_val);
    else if (pl.h_status.n_wide) {
        LAUNCH(ctx, k_grid_macaque_warp, div_up(pl.h_status.n_wide, WIDE_WARPS), WIDE_WARPS * 32, 0, v, pl.desc.p, pl.point_off.p,
               pl.worklist.p + (S - 1), (uint32_t)pl.h_status.n_wide, ctx->block_row_min, d_val);
        launch_macaque_block<false>(ctx, std::min<unsigned int>(pl.h_status.n_wide, (unsigned int)ctx->sm_count * 16), v, pl.worklist.p + (S - 1), -1,
                                    (const unsigned int *)nullptr, (uint32_t)pl.h_status.n_wide, pl.point_off.p, d_val);
    }
    return MDBCU_SUCCESS;
}

int mdbcu_grid(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments, int64_t *timestamps_out, float *values_out,
               uint64_t capacity, uint64_t *n_points) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    StagedSegments st;
    if (stage_segments(ctx, space, segments, st)) return MDBCU_FAILURE;
    GridPlan pl;
    if (grid_plan(ctx, st.view, pl)) return MDBCU_FAILURE;
    if (n_points) *n_points = pl.total;
    if (pl.total > capacity) return fail("grid: batch holds " + std::to_string(pl.total) + " data points, capacity is " + std::to_string(capacity));
    if (pl.total == 0) return MDBCU_SUCCESS;
    if (!timestamps_out || !values_out) return fail("grid: output pointer is null");
    cudaStream_t s = ctx->stream;

    int64_t *d_ts = timestamps_out;
    float *d_val = values_out;
    DBuf<int64_t> ts_buf;
    DBuf<float> val_buf;
    if (space == MDBCU_HOST) {
        CUDA_TRY(ts_buf.alloc(pl.total, s));
        CUDA_TRY(val_buf.alloc(pl.total, s));
        d_ts = ts_buf.p;
        d_val = val_buf.p;
    }
    if (grid_on_device(ctx, st.view, pl, d_ts, d_val)) return MDBCU_FAILURE;
    CUDA_TRY(cudaGetLastError());
    if (space == MDBCU_HOST) {
        CUDA_TRY(d2h_bytes(ctx, timestamps_out, d_ts, pl.total * sizeof(int64_t)));
        CUDA_TRY(d2h_bytes(ctx, values_out, d_val, pl.total * sizeof(float)));
    }
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}


// ---- grid with the query's time predicate pushed in -------------------------------------------------------------------
// The reference reconstructs every data point of every segment and then prunes by the predicate (grid_exec.rs:366-387: "for
// simplicity, all data points are reconstructed and then pruned by time"); segments are selected by start_time / end_time
// before that (time_series_table.rs:290-373).  Here both happen on the device and before any point leaves it:
//   k_range_select   one thread per row: does [start_time, end_time] meet [t_lo, t_hi]?          -> scan -> kept row ids
//   k_range_lengths / k_range_gather   the kept rows as a dense batch of their own (byte columns copied: ~30 bytes a row)
//   grid_plan + grid_on_device         the ordinary K2 over that batch, into scratch
//   k_range_clip     one thread per kept row: timestamps within a row ascend, so the points inside the range are one index
//                    range, found by two binary searches in the row's reconstructed timestamps  -> scan -> output offsets
//   k_range_copy     one warp per kept row: that index range to its place in the output
// Rows outside the range cost one comparison; only the rows straddling t_lo or t_hi are reconstructed beyond what is returned.
__global__ void __launch_bounds__(256) k_range_select(SegmentsView v, int64_t t_lo, int64_t t_hi, uint32_t *keep) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= v.n_segments) return;
    keep[s] = (v.end_time[s] >= t_lo && v.start_time[s] <= t_hi) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_range_lengths(SegmentsView v, const uint32_t *keep, const uint64_t *kidx, uint32_t *rows, uint32_t *tl,
                                                       uint32_t *vl, uint32_t *rl, Status *status) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= v.n_segments || !keep[s]) return;
    const uint64_t k = kidx[s];
    rows[k] = (uint32_t)s;
    const uint64_t a = v.timestamps_off[s + 1] - v.timestamps_off[s], b = v.values_off[s + 1] - v.values_off[s],
                   c = v.residuals_off[s + 1] - v.residuals_off[s];
    // (offsets that run backwards, or a row of more than 4 GiB: load_row would refuse it too)
    if (v.timestamps_off[s + 1] < v.timestamps_off[s] || v.values_off[s + 1] < v.values_off[s] || v.residuals_off[s + 1] < v.residuals_off[s] ||
        (a | b | c) > 0xFFFFFFF0ull)
        report_bad(status, s);
    tl[k] = (uint32_t)a;
    vl[k] = (uint32_t)b;
    rl[k] = (uint32_t)c;
}
__global__ void __launch_bounds__(128) k_range_gather(SegmentsView v, const uint32_t *rows, uint64_t n_kept, int8_t *type, int64_t *start, int64_t *end,
                                                      float *mn, float *mx, const uint64_t *t_off, uint8_t *t_data, const uint64_t *v_off, uint8_t *v_data,
                                                      const uint64_t *r_off, uint8_t *r_data) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_kept) return;
    const uint64_t s = rows[k];
    type[k] = v.model_type_id[s];
    start[k] = v.start_time[s];
    end[k] = v.end_time[s];
    mn[k] = v.min_value[s];
    mx[k] = v.max_value[s];
    const uint8_t *src = v.timestamps_data + v.timestamps_off[s];
    for (uint64_t i = 0, n = t_off[k + 1] - t_off[k]; i < n; i++) t_data[t_off[k] + i] = src[i];
    src = v.values_data + v.values_off[s];
    if (v_off[k + 1] - v_off[k] < 4096) // (longer ones: k_range_gather_long)
        for (uint64_t i = 0, n = v_off[k + 1] - v_off[k]; i < n; i++) v_data[v_off[k] + i] = src[i];
    src = v.residuals_data + v.residuals_off[s];
    for (uint64_t i = 0, n = r_off[k + 1] - r_off[k]; i < n; i++) r_data[r_off[k] + i] = src[i];
}
// Long rows (a MacaqueV row can hold a whole series) have their bytes copied by a block each.
__global__ void __launch_bounds__(256) k_range_gather_long(SegmentsView v, const uint32_t *rows, uint64_t n_kept, const uint64_t *v_off, uint8_t *v_data) {
    for (uint64_t k = blockIdx.x; k < n_kept; k += gridDim.x) {
        const uint64_t n = v_off[k + 1] - v_off[k];
        if (n < 4096) continue;
        const uint8_t *src = v.values_data + v.values_off[rows[k]];
        for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) v_data[v_off[k] + i] = src[i];
    }
}
__global__ void __launch_bounds__(256) k_range_clip(const uint64_t *point_off, const int64_t *ts, uint64_t n_kept, int64_t t_lo, int64_t t_hi, uint32_t *first,
                                                    uint32_t *count) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_kept) return;
    const int64_t *row = ts + point_off[k];
    const uint32_t n = (uint32_t)(point_off[k + 1] - point_off[k]);
    uint32_t lo = 0, hi = n; // first index with row[i] >= t_lo
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (row[mid] < t_lo) lo = mid + 1;
        else hi = mid;
    }
    const uint32_t a = lo;
    hi = n; // first index with row[i] > t_hi
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (row[mid] <= t_hi) lo = mid + 1;
        else hi = mid;
    }
    first[k] = a;
    count[k] = lo - a;
}
constexpr uint64_t RANGE_LONG_ROW = 16384;
__global__ void __launch_bounds__(128) k_range_copy(const uint64_t *point_off, const uint32_t *first, const uint64_t *out_off, uint64_t n_kept,
                                                    const int64_t *ts, const float *val, int64_t *ts_out, float *val_out) {
    const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_kept) return;
    const int lane = threadIdx.x & 31;
    const uint64_t src = point_off[k] + first[k], dst = out_off[k], n = min(out_off[k + 1] - dst, RANGE_LONG_ROW); // (the rest: k_range_copy_long)
    for (uint64_t i = lane; i < n; i += 32) {
        ts_out[dst + i] = ts[src + i];
        val_out[dst + i] = val[src + i];
    }
}
// Rows of more than RANGE_LONG_ROW points: the rest is copied by a block (one warp would take a 10^6-point row alone).
__global__ void __launch_bounds__(256) k_range_copy_long(const uint64_t *point_off, const uint32_t *first, const uint64_t *out_off, uint64_t n_kept,
                                                         const int64_t *ts, const float *val, int64_t *ts_out, float *val_out) {
    for (uint64_t k = blockIdx.x; k < n_kept; k += gridDim.x) {
        const uint64_t src = point_off[k] + first[k], dst = out_off[k], n = out_off[k + 1] - dst;
        if (n <= RANGE_LONG_ROW) continue;
        for (uint64_t i = RANGE_LONG_ROW + threadIdx.x; i < n; i += blockDim.x) {
            ts_out[dst + i] = ts[src + i];
            val_out[dst + i] = val[src + i];
        }
    }
}
__global__ void __launch_bounds__(256) k_range_row_counts(const uint32_t *rows, const uint32_t *count, uint64_t n_kept, uint32_t *per_row) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_kept) per_row[rows[k]] = count[k];
}

int mdbcu_grid_range(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments, int64_t t_lo, int64_t t_hi, uint64_t *point_off,
                     int64_t *timestamps_out, float *values_out, uint64_t capacity, uint64_t *n_points) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    StagedSegments st;
    if (stage_segments(ctx, space, segments, st)) return MDBCU_FAILURE;
    cudaStream_t s = ctx->stream;
    const uint64_t S = st.view.n_segments;
    if (n_points) *n_points = 0;
    auto give_point_off = [&](const uint64_t *d_off) -> int { // (d_off == nullptr: all zero)
        if (!point_off) return MDBCU_SUCCESS;
        const size_t bytes = (S + 1) * sizeof(uint64_t);
        if (!d_off) {
            if (space == MDBCU_HOST) std::memset(point_off, 0, bytes);
            else CUDA_TRY(cudaMemsetAsync(point_off, 0, bytes, s));
        } else if (space == MDBCU_HOST) {
            CUDA_TRY(d2h_bytes(ctx, point_off, d_off, bytes));
        } else {
            CUDA_TRY(cudaMemcpyAsync(point_off, d_off, bytes, cudaMemcpyDeviceToDevice, s));
        }
        return MDBCU_SUCCESS;
    };
    if (S == 0 || t_lo > t_hi) {
        if (give_point_off(nullptr)) return MDBCU_FAILURE;
        CUDA_TRY(sync_stream(ctx));
        return MDBCU_SUCCESS;
    }
    // ---- the rows that meet the range, as a batch of their own
    DBuf<uint32_t> keep, rows, tl, vl, rl;
    DBuf<uint64_t> kidx;
    DBuf<Status> status;
    CUDA_TRY(keep.alloc(S, s));
    CUDA_TRY(kidx.alloc(S + 1, s));
    if (new_status(ctx, status)) return MDBCU_FAILURE;
    LAUNCH(ctx, k_range_select, div_up(S, 256), 256, 0, st.view, t_lo, t_hi, keep.p);
    if (exclusive_scan<uint32_t>(ctx, keep.p, S, kidx.p)) return MDBCU_FAILURE;
    CUDA_TRY(post(ctx, 0, kidx.p + S, 1));
    CUDA_TRY(sync_stream(ctx));
    const uint64_t K = ctx->mailbox[0];
    if (K == 0) {
        if (give_point_off(nullptr)) return MDBCU_FAILURE;
        CUDA_TRY(sync_stream(ctx));
        return MDBCU_SUCCESS;
    }
    CUDA_TRY(rows.alloc(K, s));
    CUDA_TRY(tl.alloc(K, s));
    CUDA_TRY(vl.alloc(K, s));
    CUDA_TRY(rl.alloc(K, s));
    LAUNCH(ctx, k_range_lengths, div_up(S, 256), 256, 0, st.view, keep.p, kidx.p, rows.p, tl.p, vl.p, rl.p, status.p);
    StagedSegments sub;
    CUDA_TRY(sub.model_type_id.alloc(K, s));
    CUDA_TRY(sub.start_time.alloc(K, s));
    CUDA_TRY(sub.end_time.alloc(K, s));
    CUDA_TRY(sub.min_value.alloc(K, s));
    CUDA_TRY(sub.max_value.alloc(K, s));
    CUDA_TRY(sub.ts_off.alloc(K + 1, s));
    CUDA_TRY(sub.val_off.alloc(K + 1, s));
    CUDA_TRY(sub.res_off.alloc(K + 1, s));
    if (exclusive_scan<uint32_t>(ctx, tl.p, K, sub.ts_off.p)) return MDBCU_FAILURE;
    if (exclusive_scan<uint32_t>(ctx, vl.p, K, sub.val_off.p)) return MDBCU_FAILURE;
    if (exclusive_scan<uint32_t>(ctx, rl.p, K, sub.res_off.p)) return MDBCU_FAILURE;
    CUDA_TRY(post(ctx, 0, sub.ts_off.p + K, 1));
    CUDA_TRY(post(ctx, 1, sub.val_off.p + K, 1));
    CUDA_TRY(post(ctx, 2, sub.res_off.p + K, 1));
    {
        Status h;
        if (read_status(ctx, status.p, h, "segment row (offsets)")) return MDBCU_FAILURE; // (synchronises the stream)
    }
    const uint64_t tb = ctx->mailbox[0], vb = ctx->mailbox[1], rb = ctx->mailbox[2];
    CUDA_TRY(sub.ts_data.alloc(tb ? tb : 1, s));
    CUDA_TRY(sub.val_data.alloc(vb ? vb : 1, s));
    CUDA_TRY(sub.res_data.alloc(rb ? rb : 1, s));
    LAUNCH(ctx, k_range_gather, div_up(K, 128), 128, 0, st.view, rows.p, K, sub.model_type_id.p, sub.start_time.p, sub.end_time.p, sub.min_value.p,
           sub.max_value.p, sub.ts_off.p, sub.ts_data.p, sub.val_off.p, sub.val_data.p, sub.res_off.p, sub.res_data.p);
    LAUNCH(ctx, k_range_gather_long, (unsigned int)std::min<uint64_t>(K, (uint64_t)ctx->sm_count * 8), 256, 0, st.view, rows.p, K, sub.val_off.p,
           sub.val_data.p);
    sub.view.n_segments = K;
    sub.view.model_type_id = sub.model_type_id.p;
    sub.view.start_time = sub.start_time.p;
    sub.view.end_time = sub.end_time.p;
    sub.view.min_value = sub.min_value.p;
    sub.view.max_value = sub.max_value.p;
    sub.view.timestamps_off = sub.ts_off.p;
    sub.view.timestamps_data = sub.ts_data.p;
    sub.view.values_off = sub.val_off.p;
    sub.view.values_data = sub.val_data.p;
    sub.view.residuals_off = sub.res_off.p;
    sub.view.residuals_data = sub.res_data.p;
    // ---- their points (scratch), the index range of each row inside [t_lo, t_hi], and those ranges to the output
    GridPlan pl;
    if (grid_plan(ctx, sub.view, pl)) return MDBCU_FAILURE;
    DBuf<int64_t> ts_s;
    DBuf<float> val_s;
    CUDA_TRY(ts_s.alloc(pl.total ? pl.total : 1, s));
    CUDA_TRY(val_s.alloc(pl.total ? pl.total : 1, s));
    if (pl.total && grid_on_device(ctx, sub.view, pl, ts_s.p, val_s.p)) return MDBCU_FAILURE;
    DBuf<uint32_t> first, count, per_row;
    DBuf<uint64_t> out_off, row_off;
    CUDA_TRY(first.alloc(K, s));
    CUDA_TRY(count.alloc(K, s));
    CUDA_TRY(out_off.alloc(K + 1, s));
    LAUNCH(ctx, k_range_clip, div_up(K, 256), 256, 0, pl.point_off.p, ts_s.p, K, t_lo, t_hi, first.p, count.p);
    if (exclusive_scan<uint32_t>(ctx, count.p, K, out_off.p)) return MDBCU_FAILURE;
    CUDA_TRY(post(ctx, 0, out_off.p + K, 1));
    if (point_off) {
        CUDA_TRY(per_row.alloc(S, s));
        CUDA_TRY(row_off.alloc(S + 1, s));
        CUDA_TRY(cudaMemsetAsync(per_row.p, 0, S * sizeof(uint32_t), s));
        LAUNCH(ctx, k_range_row_counts, div_up(K, 256), 256, 0, rows.p, count.p, K, per_row.p);
        if (exclusive_scan<uint32_t>(ctx, per_row.p, S, row_off.p)) return MDBCU_FAILURE;
        if (give_point_off(row_off.p)) return MDBCU_FAILURE;
    }
    CUDA_TRY(sync_stream(ctx));
    const uint64_t N = ctx->mailbox[0];
    if (n_points) *n_points = N;
    if (!timestamps_out && !values_out) return MDBCU_SUCCESS; // a count
    if (N > capacity) return fail("grid_range: " + std::to_string(N) + " data points lie in the range, capacity is " + std::to_string(capacity));
    if (N == 0) return MDBCU_SUCCESS;
    if (!timestamps_out || !values_out) return fail("grid_range: output pointer is null");
    int64_t *d_ts = timestamps_out;
    float *d_val = values_out;
    DBuf<int64_t> ts_buf;
    DBuf<float> val_buf;
    if (space == MDBCU_HOST) {
        CUDA_TRY(ts_buf.alloc(N, s));
        CUDA_TRY(val_buf.alloc(N, s));
        d_ts = ts_buf.p;
        d_val = val_buf.p;
    }
    LAUNCH(ctx, k_range_copy, div_up(K * 32, 128), 128, 0, pl.point_off.p, first.p, out_off.p, K, ts_s.p, val_s.p, d_ts, d_val);
    if (N > RANGE_LONG_ROW)
        LAUNCH(ctx, k_range_copy_long, (unsigned int)std::min<uint64_t>(K, (uint64_t)ctx->sm_count * 8), 256, 0, pl.point_off.p, first.p, out_off.p, K, ts_s.p,
               val_s.p, d_ts, d_val);
    CUDA_TRY(cudaGetLastError());
    if (space == MDBCU_HOST) {
        CUDA_TRY(d2h_bytes(ctx, timestamps_out, d_ts, N * sizeof(int64_t)));
        CUDA_TRY(d2h_bytes(ctx, values_out, d_val, N * sizeof(float)));
    }
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}

// ---- K3 ------------------------------------------------------------------------------------------

int mdbcu_segment_sums(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments, float *sums_out) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    StagedSegments st;
    if (stage_segments(ctx, space, segments, st)) return MDBCU_FAILURE;
    uint64_t S = st.view.n_segments;
    if (S == 0) return MDBCU_SUCCESS;
    if (!sums_out) return fail("segment_sums: output pointer is null");
    cudaStream_t s = ctx->stream;
    DBuf<Status> status;
    if (new_status(ctx, status)) return MDBCU_FAILURE;
    DBuf<float> sum_buf;
    float *d_sum = sums_out;
    if (space == MDBCU_HOST) {
        CUDA_TRY(sum_buf.alloc(S, s));
        d_sum = sum_buf.p;
    }
    DBuf<uint32_t> wide_list;
    CUDA_TRY(wide_list.alloc(S, s));
    LAUNCH(ctx, k_agg_segments, div_up(S, 128), 128, 0, st.view, (uint64_t *)nullptr, d_sum, wide_list.p, status.p);
    LAUNCH(ctx, k_agg_macaque_warp, std::min<unsigned int>(div_up(S, WIDE_WARPS), (unsigned int)ctx->sm_count * 8), WIDE_WARPS * 32, 0, st.view,
           wide_list.p, &status.p->n_wide, ctx->lane_rows_min, ctx->block_row_min, d_sum);
    launch_macaque_block<true>(ctx, std::min<unsigned int>((unsigned int)std::min<uint64_t>(S, 1u << 20), (unsigned int)ctx->sm_count * 16), st.view, wide_list.p, 1,
                               &status.p->n_wide, (uint32_t)std::min<uint64_t>(S, 0xFFFFFFFFull), (const uint64_t *)nullptr, d_sum);
    if (S >= ctx->lane_rows_min)
        LAUNCH(ctx, k_agg_macaque_lanes, std::min<unsigned int>(div_up(S, 128), (unsigned int)ctx->sm_count * 16), 128, 0, st.view, wide_list.p,
               &status.p->n_wide, ctx->lane_rows_min, d_sum);
    CUDA_TRY(cudaGetLastError());
    if (space == MDBCU_HOST) CUDA_TRY(d2h_bytes(ctx, sums_out, d_sum, S * sizeof(float)));
    Status h;
    return read_status(ctx, status.p, h, "segment row");
}

int mdbcu_aggregate(mdbcu_context *ctx, mdbcu_space space, const mdbcu_segments_view *segments, const uint64_t *group_off, uint64_t n_groups,
                    int64_t *count, float *min, float *max, double *sum) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (!group_off) n_groups = 1;
    if (n_groups == 0) return MDBCU_SUCCESS;
    if (!count || !min || !max || !sum) return fail("aggregate: output pointer is null");
    StagedSegments st;
    if (stage_segments(ctx, space, segments, st)) return MDBCU_FAILURE;
    uint64_t S = st.view.n_segments;
    cudaStream_t s = ctx->stream;

    DBuf<uint64_t> group_off_buf;
    const uint64_t *d_group_off = group_off;
    if (group_off && space == MDBCU_HOST) {
        if (group_off[n_groups] > S) return fail("aggregate: group_off exceeds the number of rows");
        CUDA_TRY(upload(ctx, group_off_buf, group_off, n_groups + 1));
        d_group_off = group_off_buf.p;
    }
    DBuf<Status> status;
    if (new_status(ctx, status)) return MDBCU_FAILURE;
    DBuf<float> wide_sum; // f32 sums of the long MacaqueV rows only (sparse; every other row is folded in registers)
    DBuf<uint32_t> wide_list;
    CUDA_TRY(wide_sum.alloc(S, s));
    CUDA_TRY(wide_list.alloc(S, s));
    if (S) {
        LAUNCH(ctx, k_agg_find_wide, div_up(S, 256), 256, 0, st.view, wide_list.p, status.p);
        LAUNCH(ctx, k_agg_macaque_warp, std::min<unsigned int>(div_up(S, WIDE_WARPS), (unsigned int)ctx->sm_count * 8), WIDE_WARPS * 32, 0,
               st.view, wide_list.p, &status.p->n_wide, ctx->lane_rows_min, ctx->block_row_min, wide_sum.p);
        launch_macaque_block<true>(ctx, std::min<unsigned int>((unsigned int)std::min<uint64_t>(S, 1u << 20), (unsigned int)ctx->sm_count * 16), st.view, wide_list.p, 1,
                                   &status.p->n_wide, (uint32_t)std::min<uint64_t>(S, 0xFFFFFFFFull), (const uint64_t *)nullptr, wide_sum.p);
        if (S >= ctx->lane_rows_min)
            LAUNCH(ctx, k_agg_macaque_lanes, std::min<unsigned int>(div_up(S, 128), (unsigned int)ctx->sm_count * 16), 128, 0, st.view,
                   wide_list.p, &status.p->n_wide, ctx->lane_rows_min, wide_sum.p);
    }

    // Shape of the reduction: a function of the batch alone (never of the device), so that the tree -- and with it the last
    // bits of SUM -- is the same everywhere.  Small groups: one warp each.  Large groups: blocks of 256 threads over slices of
    // ~MDB_AGG_SLICE rows (parts per group), then a fold of the parts.
    const uint64_t avg_rows = S / n_groups + 1;
    const bool warp_groups = avg_rows < 512;
    const uint32_t parts = warp_groups ? 1u : (uint32_t)std::min<uint64_t>(1024, (avg_rows + MDB_AGG_SLICE - 1) / MDB_AGG_SLICE);
    DBuf<GroupAgg> partial;
    if (!warp_groups) {
        CUDA_TRY(partial.alloc(n_groups * parts, s));
        dim3 grid(parts, (unsigned int)std::min<uint64_t>(n_groups, 65535));
        LAUNCH(ctx, k_agg_groups, grid, AGG_THREADS, 0, st.view, d_group_off, n_groups, parts, wide_sum.p, partial.p, status.p);
    }

    DBuf<int64_t> count_buf;
    DBuf<float> min_buf, max_buf;
    DBuf<double> sum_buf;
    int64_t *d_count = count;
    float *d_min = min, *d_max = max;
    double *d_sum = sum;
    if (space == MDBCU_HOST) {
        CUDA_TRY(count_buf.alloc(n_groups, s));
        CUDA_TRY(min_buf.alloc(n_groups, s));
        CUDA_TRY(max_buf.alloc(n_groups, s));
        CUDA_TRY(sum_buf.alloc(n_groups, s));
        d_count = count_buf.p; d_min = min_buf.p; d_max = max_buf.p; d_sum = sum_buf.p;
    }
    if (warp_groups)
        LAUNCH(ctx, k_agg_groups_warp, div_up(n_groups, AGG_THREADS / 32), AGG_THREADS, 0, st.view, d_group_off, n_groups, wide_sum.p, d_count, d_min,
               d_max, d_sum, status.p);
    else
        LAUNCH(ctx, k_agg_final, div_up(n_groups, AGG_THREADS), AGG_THREADS, 0, partial.p, n_groups, parts, d_count, d_min, d_max, d_sum);
    CUDA_TRY(cudaGetLastError());
    if (space == MDBCU_HOST) {
        CUDA_TRY(d2h_bytes(ctx, count, d_count, n_groups * sizeof(int64_t)));
        CUDA_TRY(d2h_bytes(ctx, min, d_min, n_groups * sizeof(float)));
        CUDA_TRY(d2h_bytes(ctx, max, d_max, n_groups * sizeof(float)));
        CUDA_TRY(d2h_bytes(ctx, sum, d_sum, n_groups * sizeof(double)));
    }
    Status h;
    return read_status(ctx, status.p, h, "segment row");
}

// ---- K1 ------------------------------------------------------------------------------------------

uint64_t mdbcu_segments_len(const mdbcu_segments *segments) { return segments ? segments->n_segments : 0; }

void mdbcu_segments_free(mdbcu_segments *sg) {
    if (!sg) return;
    if (sg->ctx) {
        cudaSetDevice(sg->ctx->device);
        cudaStream_t s = sg->ctx->stream;
        void *ptrs[] = {sg->model_type_id, sg->start_time, sg->end_time, sg->min_value, sg->max_value, sg->ts_off, sg->val_off,
                        sg->res_off, sg->unit_seg_off, sg->ts_data, sg->val_data, sg->res_data};
        for (void *p : ptrs)
            if (p) cudaFreeAsync(p, s);
        pinned_release(sg->ctx, sg->host_block);
    }
    delete sg;
}

int mdbcu_segments_get(mdbcu_segments *sg, mdbcu_space space, mdbcu_segments_view *view, const uint64_t **unit_seg_off) {
    if (!sg || !view) return fail("segments or view is null");
    if (check_ctx(sg->ctx)) return MDBCU_FAILURE;
    uint64_t S = sg->n_segments;
    if (space == MDBCU_DEVICE) {
        view->n_segments = S;
        view->model_type_id = sg->model_type_id;
        view->start_time = sg->start_time;
        view->end_time = sg->end_time;
        view->min_value = sg->min_value;
        view->max_value = sg->max_value;
        view->timestamps_off = sg->ts_off;
        view->timestamps_data = sg->ts_data;
        view->values_off = sg->val_off;
        view->values_data = sg->val_data;
        view->residuals_off = sg->res_off;
        view->residuals_data = sg->res_data;
        if (unit_seg_off) *unit_seg_off = sg->unit_seg_off;
        return MDBCU_SUCCESS;
    }
    if (!sg->have_host) {
        mdbcu_context *ctx = sg->ctx;
        // one pinned block, columns at 64-byte aligned offsets
        const void *src[12] = {sg->model_type_id, sg->start_time, sg->end_time, sg->min_value, sg->max_value, sg->ts_off,
                               sg->val_off, sg->res_off, sg->unit_seg_off, sg->ts_data, sg->val_data, sg->res_data};
        const size_t bytes[12] = {S, S * 8, S * 8, S * 4, S * 4, (S + 1) * 8, (S + 1) * 8, (S + 1) * 8, (sg->n_units + 1) * 8,
                                  sg->ts_bytes, sg->val_bytes, sg->res_bytes};
        size_t at[12], total = 0;
        for (int c = 0; c < 12; c++) {
            at[c] = total;
            total += (bytes[c] + 63) & ~(size_t)63;
        }
        CUDA_TRY(pinned_acquire(ctx, total + 64, sg->host_block));
        uint8_t *base = sg->host_block.p;
        for (int c = 0; c < 12; c++) CUDA_TRY(d2h_bytes(ctx, base + at[c], src[c], bytes[c]));
        CUDA_TRY(sync_stream(ctx));
        mdbcu_segments_view &h = sg->host_view;
        h.n_segments = S;
        h.model_type_id = (const int8_t *)(base + at[0]);
        h.start_time = (const int64_t *)(base + at[1]);
        h.end_time = (const int64_t *)(base + at[2]);
        h.min_value = (const float *)(base + at[3]);
        h.max_value = (const float *)(base + at[4]);
        h.timestamps_off = (const uint64_t *)(base + at[5]);
        h.values_off = (const uint64_t *)(base + at[6]);
        h.residuals_off = (const uint64_t *)(base + at[7]);
        sg->h_unit_seg_off = (const uint64_t *)(base + at[8]);
        h.timestamps_data = base + at[9];
        h.values_data = base + at[10];
        h.residuals_data = base + at[11];
        sg->have_host = true;
    }
    *view = sg->host_view;
    if (unit_seg_off) *unit_seg_off = sg->h_unit_seg_off;
    return MDBCU_SUCCESS;
}

} // extern "C"

#include "mdb_compress_api.inl"
#include "mdb_comm.inl"
#include "mdb_sort.inl"
