// mdb_compress_api.inl -- K1: compress kernels and mdbcu_compress (included at the end of mdb_cuda.cu).
//
// Flow (all on the context's stream; the host syncs each read back a handful of scalars through the mailbox):
//   k_unit_chunks      validate units / error bounds, chunks per unit            -> scan -> chunk_base
//   k_spec_init        ChunkState of every chunk, chunk -> unit map, list sizes  -> scan -> list_base
//   chains, either     k_sched_count / k_sched_fill / k_sched_units + k_spec_async: a device-side work queue of
//                      chunks served by persistent warps, every unit advancing its own exact frontier
//                      (sched_advance in mdb_compress.cuh; the default)
//              or      repeat { k_spec_chain[_warp] over the dirty chunks ; k_spec_propagate per unit } in global
//                      rounds until no chunk is dirty (fit engines 1 and 2)
//   k_spec_finalize    skipped chunks, residual-run ends, leading MacaqueV rows, regularity
//   k_swing_finish     the order-dependent MSE sums of the accepted Swing models (one lane per model)
//   k_spec_count_rows  rows per chunk                                            -> scan -> row_base
//   k_spec_records     SegRecord of every row in final order (byte lengths by running the encoders on a counter);
//                      long MacaqueV rows are only listed ...
//   k_records_macaque_warp   ... and encoded ONCE by a whole warp, 32 values at a time (warp_macaque_v_encode), into a staging area
//   k_compress_gather  row metadata columns + per-row byte lengths               -> 3 scans -> offsets
//   k_compress_emit    MacaqueTS / MacaqueV byte columns at their final offsets (short rows, one thread each)
//   k_copy_macaque_warp      the value bytes of the long MacaqueV rows from the staging area to their final offsets

struct CompressCounters { // device-resident, read back once per round
    unsigned int dirty;
    unsigned int pad;
};

__global__ void __launch_bounds__(256) k_unit_chunks(const uint64_t *unit_off, uint64_t n_units, const uint8_t *eb_kind, const float *eb_value,
                                                     uint32_t chunk_len, uint32_t *unit_chunks, Status *status) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    uint64_t a = unit_off[u], b = unit_off[u + 1];
    bool ok = b >= a && (b - a) < 0xFFFFFF00ull;
    uint8_t kind = eb_kind[u];
    float value = eb_value[u];
    // ErrorBound::try_new_absolute / try_new_relative (modelardb_types/src/types.rs:312-334)
    if (kind == KIND_ABSOLUTE) ok = ok && value > 0.0f && !isinf(value) && value == value;
    else if (kind == KIND_RELATIVE) ok = ok && value > 0.0f && value <= 100.0f;
    else if (kind != KIND_LOSSLESS) ok = false;
    if (!ok) report_bad(status, u);
    unit_chunks[u] = ok ? (uint32_t)((b - a + chunk_len - 1) / chunk_len) : 0;
}

// One thread per unit: initial state of its chunks (round 0 runs every chunk from its first index).
__global__ void __launch_bounds__(128) k_spec_init(const uint64_t *unit_off, uint64_t n_units, const uint64_t *chunk_base, uint32_t chunk_len,
                                                   ChunkState *st, uint32_t *chunk_unit, uint32_t *list_cap) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    uint64_t n = unit_off[u + 1] - unit_off[u];
    uint64_t g0 = chunk_base[u], C = chunk_base[u + 1] - g0;
    for (uint64_t c = 0; c < C; c++) {
        ChunkState s;
        s.entry = IDX_NONE;
        s.exit = IDX_NONE;
        s.truncated_at = 0;
        s.n_models = 0;
        s.new_entry = (uint32_t)(c * chunk_len);
        s.next_start = 0;
        s.lead_end = IDX_NONE;
        s.rows = 0;
        s.dirty = 1;
        s.exact = c == 0;
        s.buf = 0;
        s.skipped = 0;
        s.irregular = 0;
        s.pad[0] = s.pad[1] = s.pad[2] = 0;
        s.first_start = IDX_NONE;
        s.phase = PH_QUEUED;
        st[g0 + c] = s;
        chunk_unit[g0 + c] = (uint32_t)u;
        uint64_t rest = n - c * chunk_len;
        uint64_t len = rest < chunk_len ? rest : chunk_len;
        list_cap[g0 + c] = 2 * models_per_chunk((uint32_t)len); // two buffers
    }
}

// One thread per chunk; a block is one warp of which the first `lanes` lanes carry a chain (few chains:
// one per warp, no divergence and every SM busy; many chains: lanes fill up).
__global__ void __launch_bounds__(32) k_spec_chain(const int64_t *__restrict__ ts, const float *__restrict__ values, const uint64_t *__restrict__ unit_off,
                                                   const uint8_t *__restrict__ eb_kind, const float *__restrict__ eb_value,
                                                   const uint64_t *__restrict__ chunk_base, const uint32_t *__restrict__ chunk_unit,
                                                   const uint32_t *__restrict__ worklist, uint64_t n_work, uint32_t lanes, uint32_t chunk_len,
                                                   ChunkState *st, FittedModel *lists, const uint64_t *__restrict__ list_base,
                                                   const uint32_t *__restrict__ list_cap) {
    if (threadIdx.x >= lanes) return;
    uint64_t w = (uint64_t)blockIdx.x * lanes + threadIdx.x;
    if (w >= n_work) return;
    uint64_t g = worklist ? worklist[w] : w; // round 0: every chunk
    ChunkState s = st[g];
    if (!s.dirty) return;
    uint32_t u = chunk_unit[g];
    uint64_t a = unit_off[u];
    uint32_t n = (uint32_t)(unit_off[u + 1] - a);
    uint32_t c = (uint32_t)(g - chunk_base[u]);
    uint32_t chunk_start = c * chunk_len;
    uint32_t chunk_end = (uint64_t)chunk_start + chunk_len < n ? chunk_start + chunk_len : n;
    ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
    ScalarFit fitter(eb, ts + a, values + a, n);
    spec_chain(fitter, 0u, 1u, n, chunk_end, chunk_len, s, lists + list_base[g], list_cap[g] / 2);
    st[g] = s;
}

// One WARP per chunk: the chain's control flow is uniform across the lanes and every fit is done by the
// 32 lanes together (mdb_fit_warp.cuh).  This is the default engine: it does not need tens of
// thousands of independent chains to fill the GPU, and a warp's loads of 32 consecutive points are
// fully coalesced (128 B of values, 256 B of timestamps per step).
constexpr int CHAIN_WARPS = 4;
// three blocks of four warps per SM (168 registers per thread; measured best of 2 / 3 / 4 on B200)
#ifndef MDB_CHAIN_MIN_BLOCKS
#define MDB_CHAIN_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(CHAIN_WARPS * 32, MDB_CHAIN_MIN_BLOCKS) k_spec_chain_warp(const int64_t *__restrict__ ts, const float *__restrict__ values,
                                                                      const uint64_t *__restrict__ unit_off, const uint8_t *__restrict__ eb_kind,
                                                                      const float *__restrict__ eb_value, const uint64_t *__restrict__ chunk_base,
                                                                      const uint32_t *__restrict__ chunk_unit, const uint32_t *__restrict__ worklist,
                                                                      uint64_t n_work, uint32_t chunk_len, ChunkState *st, FittedModel *lists,
                                                                      const uint64_t *__restrict__ list_base, const uint32_t *__restrict__ list_cap) {
    __shared__ double smem[CHAIN_WARPS][WarpFit::SMEM_DOUBLES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t w = (uint64_t)blockIdx.x * CHAIN_WARPS + warp;
    if (w >= n_work) return; // whole warps leave together
    uint64_t g = worklist ? worklist[w] : w; // round 0: every chunk
    ChunkState s = st[g];
    if (!s.dirty) return;
    uint32_t u = chunk_unit[g];
    uint64_t a = unit_off[u];
    uint32_t n = (uint32_t)(unit_off[u + 1] - a);
    uint32_t c = (uint32_t)(g - chunk_base[u]);
    uint32_t chunk_start = c * chunk_len;
    uint32_t chunk_end = (uint64_t)chunk_start + chunk_len < n ? chunk_start + chunk_len : n;
    ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
    WarpFit fitter(eb, ts + a, values + a, n, smem[warp]);
    spec_chain(fitter, (uint32_t)lane, 32u, n, chunk_end, chunk_len, s, lists + list_base[g], list_cap[g] / 2);
    __syncwarp();
    s.phase = PH_DONE; // (only read by the asynchronous scheduler, when these rounds run before it: the chunk has a chain now)
    if (lane == 0) st[g] = s;
}

// ---- one lane per chain (mdb_fit_lanes.cuh): the bulk of the chains, before the exact stitching ------------------
//
// k_lanes_units    one thread per unit: LaneUnit (first timestamp, interval, bound constants) and whether the unit
//                  qualifies at all (at least two points, positive interval, |timestamps| < 2^53, exact relative test)
// k_lanes_regular  one block per chunk, coalesced: every interval of the chunk equals the unit's first one, else the
//                  unit is marked irregular (the lanes then leave it alone)
// k_spec_lanes     persistent warps; every LANE claims chunks from a counter and runs the chunk's chain from its first
//                  index (LaneChain), reading its values through a 16-value ring in shared memory that is refilled a
//                  16-byte quad at a time, one quad ahead of the cursor
// k_sched_kick     one thread per unit: the first sched_advance, which finds the first chunk whose chain did not start
//                  at the exact entry and queues its re-run for k_spec_async

__global__ void __launch_bounds__(128) k_lanes_units(const int64_t *__restrict__ ts, const uint64_t *__restrict__ unit_off, uint64_t n_units,
                                                     const uint8_t *__restrict__ eb_kind, const float *__restrict__ eb_value, LaneUnit *info,
                                                     unsigned int *kind_units) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const uint64_t a = unit_off[u], n = unit_off[u + 1] - a;
    const ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
    const LaneUnit lu = lane_unit_init(ts + a, n, eb);
    info[u] = lu;
    if (lu.ok) atomicAdd(&kind_units[eb.kind], 1u);
}

constexpr int REGULAR_THREADS = 256;
// Units that passed lane_unit_init but turned out irregular: when the check below runs BESIDE the chain kernel (see
// mdbcu_compress), chains of such a unit may have been screened under a wrong assumption and the chains are redone.
__global__ void __launch_bounds__(256) k_lanes_late(const LaneUnit *info, uint64_t n_units, unsigned int *count) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_units && info[u].ok && info[u].irregular) atomicAdd(count, 1u);
}

__global__ void __launch_bounds__(REGULAR_THREADS) k_lanes_regular(const int64_t *__restrict__ ts, const uint64_t *__restrict__ unit_off,
                                                                   const uint64_t *__restrict__ chunk_base, const uint32_t *__restrict__ chunk_unit,
                                                                   uint32_t chunk_len, LaneUnit *info) {
    const uint64_t g = blockIdx.x;
    const uint32_t u = chunk_unit[g];
    if (!info[u].ok) return;
    const uint64_t a = unit_off[u];
    const uint32_t n = (uint32_t)(unit_off[u + 1] - a);
    const uint32_t c = (uint32_t)(g - chunk_base[u]);
    const uint32_t lo = c * chunk_len, hi = (uint64_t)lo + chunk_len < n ? lo + chunk_len : n;
    const int64_t *uts = ts + a;
    const int64_t t0 = uts[0], delta = uts[1] - t0;
    bool bad = false;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += REGULAR_THREADS) bad |= uts[i] != t0 + (int64_t)i * delta;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(&info[u].irregular, 1u);
}

constexpr uint32_t LANE_ROUNDS = 4; // repair rounds after the first pass (see mdbcu_compress)
constexpr int LANES_WARPS = 4;
constexpr int LANES_RING = 16; // values per lane: four 16-byte quads (the one before the cursor's, the cursor's, one or two ahead)
#ifndef MDB_LANES_MIN_BLOCKS
#define MDB_LANES_MIN_BLOCKS 5
#endif

template <int KIND>
__global__ void __launch_bounds__(LANES_WARPS * 32, MDB_LANES_MIN_BLOCKS) k_spec_lanes(const float *__restrict__ values, uint64_t n_total,
                                                                  const uint64_t *__restrict__ unit_off, const LaneUnit *__restrict__ info,
                                                                  const uint64_t *__restrict__ chunk_base, const uint32_t *__restrict__ chunk_unit,
                                                                  uint32_t chunk_len, uint32_t warmup, const uint32_t *__restrict__ worklist,
                                                                  uint32_t n_chunks, ChunkState *st, FittedModel *lists,
                                                                  const uint64_t *__restrict__ list_base, const uint32_t *__restrict__ list_cap,
                                                                  unsigned int *next_chunk) {
    __shared__ float ring_s[LANES_WARPS][LANES_RING][32]; // [slot][lane]: a lane's slot k is in bank `lane` whatever k is
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float(*ring)[32] = ring_s[warp];
    // the values array in 16-byte quads: element e of the quad space is values[e - skew]
    const uint32_t skew = (uint32_t)((reinterpret_cast<uintptr_t>(values) >> 2) & 3u);
    const float4 *quads = reinterpret_cast<const float4 *>(values - skew);

    bool have = false;        // this lane owns a chunk
    uint32_t g = 0;           // the chunk
    const float4 *uq = quads; // the quad that holds the unit's point 0 ...
    uint32_t sk = 0;          // ... at element sk of it: point i of the unit is element sk + i from there
    uint32_t vec_lo = 0, vec_n = 0; // quads [vec_lo, vec_lo + vec_n) (relative to uq) lie entirely inside the values array
    uint64_t eb4 = 0;         // quad-space element index of uq[0]
    LaneUnit lu;
    LaneChain chain;
    FittedModel *list = nullptr;
    // Ring: the elements [win_lo, win_lo + win_n) (relative to uq[0]) are in the ring, element e in slot e & 15.  Quads are
    // loaded in order, one per refill, `next_q` next; `pend` is the quad requested at the previous refill.
    uint32_t win_lo = 0, win_n = 0, next_q = 0;
    bool pending = false;
    float4 pend = make_float4(0.f, 0.f, 0.f, 0.f);
    bool exhausted = false;   // the chunk counter has run out

    auto load_quad = [&](uint32_t q) -> float4 {
        if (q - vec_lo < vec_n) return __ldg(uq + q);
        const uint64_t e = eb4 + (uint64_t)q * 4; // first / last quad of the array: the elements that exist
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        float *rp = reinterpret_cast<float *>(&r);
        for (int k = 0; k < 4; k++)
            if (e + k >= skew && e + k < n_total + skew) rp[k] = __ldg(values + (e + k - skew));
        return r;
    };

    while (true) {
        // ---- claim a chunk if this lane has none
        if (!have && !exhausted) {
            while (true) {
                const unsigned int w = atomicAdd(next_chunk, 1u);
                if (w >= n_chunks) {
                    exhausted = true;
                    break;
                }
                // worklist == nullptr: the first pass, every chunk from (a warm-up before) its first index.  Otherwise a round:
                // the listed chunks are re-run from the entry k_spec_propagate gave them until they meet their old chain.
                g = worklist ? worklist[w] : w;
                const uint32_t u = chunk_unit[g];
                lu = info[u];
                if (!lu.ok || lu.irregular || lu.kind != KIND) continue;
                const ChunkState s0 = st[g];
                if (worklist && (!s0.dirty || s0.entry == IDX_NONE || (s0.new_entry == s0.entry && s0.exit == IDX_NONE)))
                    continue; // never run, or a cut chain to be resumed: the cooperative engine's
                const uint64_t a = unit_off[u];
                const uint32_t n = (uint32_t)(unit_off[u + 1] - a);
                const uint32_t c = (uint32_t)(g - chunk_base[u]);
                const uint32_t lo = c * chunk_len;
                const uint32_t chunk_end = (uint64_t)lo + chunk_len < n ? lo + chunk_len : n;
                const uint32_t limit = (uint64_t)chunk_end + chunk_len < n ? chunk_end + chunk_len : n;
                const uint32_t cap = list_cap[g] / 2;
                list = lists + list_base[g] + (size_t)(s0.buf ^ 1) * cap; // the buffer the chunk's current chain is not in
                if (worklist) {
                    chain.begin(s0.new_entry, lo, chunk_end, limit, n);
                    chain.set_old_chain(lists + list_base[g] + (size_t)s0.buf * cap, s0.n_models, s0.entry, s0.exit, s0.truncated_at);
                    if (chain.splice_at(s0.new_entry, list)) { // the old chain started a fit at the new entry: nothing to run
                        ChunkState s = s0;
                        lane_chain_publish(chain, s);
                        s.phase = PH_DONE;
                        st[g] = s;
                        continue;
                    }
                } else {
                    chain.begin(lo > warmup ? lo - warmup : 0u, lo, chunk_end, limit, n); // (see LaneChain: warm-up)
                }
                sk = (uint32_t)((a + skew) & 3u);
                eb4 = (a + skew) & ~(uint64_t)3;
                uq = quads + (eb4 >> 2);
                // quads of this unit's span that can be read with one 16-byte load (all but, at most, the array's first and last)
                const uint64_t q_first = eb4 >= skew ? 0 : 1;                       // quad 0 starts before values[0]
                const uint64_t q_end = (n_total + skew - eb4) / 4;                   // first quad that reaches past the array
                vec_lo = (uint32_t)q_first;
                vec_n = q_end > q_first ? (q_end - q_first < 0xFFFFFFF0ull ? (uint32_t)(q_end - q_first) : 0xFFFFFFF0u) : 0u;
                have = true;
                win_n = 0;
                pending = false;
                next_q = (sk + chain.fit.idx) >> 2;
                win_lo = next_q * 4;
                break;
            }
        }
        if (!__any_sync(FULL_MASK, have)) break;
        // ---- ring refill: land the quad requested at the previous refill, then request the next one
        if (pending) {
            const uint32_t s4 = ((next_q - 1u) & 3u) * 4u;
            ring[s4 + 0][lane] = pend.x;
            ring[s4 + 1][lane] = pend.y;
            ring[s4 + 2][lane] = pend.z;
            ring[s4 + 3][lane] = pend.w;
            win_n += 4;
            if (win_n > 16) { // the quad just landed took the slot of the oldest one
                win_lo += 4;
                win_n = 16;
            }
            pending = false;
        }
        if (have) {
            const uint32_t e = sk + chain.fit.idx;
            if (e - win_lo > win_n) { // the cursor is outside the ring and not at its end (a jump back, or a new chunk): start over there
                next_q = e >> 2;
                win_lo = next_q * 4;
                win_n = 0;
            }
            // keep two quads ahead of the cursor's
            if (next_q <= (e >> 2) + 2 && next_q * 4 < sk + chain.limit) {
                pend = load_quad(next_q);
                next_q += 1;
                pending = true;
            }
        }
#pragma unroll
        for (int it = 0; it < 4; it++) {
            if (have) {
                const uint32_t e = sk + chain.fit.idx;
                if (e - win_lo < win_n) {
                    const float v = ring[e & 15u][lane];
                    if (chain.template step<KIND>(lu, v, list)) {
                        if (!chain.bailed) {
                            ChunkState s = st[g];
                            lane_chain_publish(chain, s);
                            s.phase = PH_DONE;
                            st[g] = s;
                        }
                        have = false;
                    }
                }
            }
        }
    }
}

// ---- asynchronous scheduling (mdb_compress.cuh: sched_advance) ------------------------------------------
struct SchedQueue {
    uint32_t head;       // oldest queued slot
    uint32_t tail;       // next free slot of `items`
    uint32_t finished;   // every live unit is final
    uint32_t units_done;
    uint32_t live_units; // units with at least one chunk
    uint32_t capacity;   // slots in `items` (a zero slot has not been written yet; chunk g is stored as g + 1)
    uint32_t init_head;  // next slot of the pre-filled part [0, n_chunks); `head` serves the pushed part behind it
    uint32_t pad;
};

// Initial queue order: chunk 0 of every unit, then chunk 1 of every unit, ...: the exact frontiers (chunk 0 is
// exact by definition) start moving in the first wave, and later chunks are still unstarted -- and can be
// re-aimed at their exact entry -- when the frontier reaches them.
// (Chunks whose chain a lane has already run -- phase PH_DONE -- are not queued: the frontier reaches them through
// sched_advance, which re-queues those that turn out to have started from the wrong entry.)
__global__ void __launch_bounds__(256) k_sched_count(const uint64_t *chunk_base, const uint32_t *chunk_unit, uint64_t n_chunks, const ChunkState *st,
                                                     uint32_t *per_index) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks || st[g].phase != PH_QUEUED) return;
    atomicAdd(&per_index[(uint32_t)(g - chunk_base[chunk_unit[g]])], 1u);
}
__global__ void __launch_bounds__(256) k_sched_fill(const uint64_t *chunk_base, const uint32_t *chunk_unit, uint64_t n_chunks, const ChunkState *st,
                                                    const uint64_t *index_base, uint32_t *cursor, uint32_t *items) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks || st[g].phase != PH_QUEUED) return;
    const uint32_t c = (uint32_t)(g - chunk_base[chunk_unit[g]]);
    items[index_base[c] + atomicAdd(&cursor[c], 1u)] = (uint32_t)g + 1u;
}
__global__ void __launch_bounds__(128) k_sched_units(const uint64_t *chunk_base, uint64_t n_units, uint64_t n_chunks, uint32_t capacity, uint32_t n_initial,
                                                     UnitSched *units, SchedQueue *q) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u == 0) {
        q->init_head = n_initial;
        q->head = (uint32_t)n_chunks;
        q->tail = (uint32_t)n_chunks;
        q->units_done = 0;
        q->capacity = capacity;
    }
    if (u >= n_units) return;
    UnitSched s;
    s.lock = 0;
    s.next_c = 0;
    s.entry = 0;
    s.finished = 0;
    units[u] = s;
    if (chunk_base[u + 1] > chunk_base[u]) atomicAdd(&q->live_units, 1u);
}

// The first sched_advance of every unit after the lanes: walks the chunks whose chains started at the exact entry and queues
// the re-run of the first one that did not (or re-aims it, if no chain has run it: a chunk the lanes left alone).
__global__ void __launch_bounds__(128) k_sched_kick(const uint64_t *__restrict__ unit_off, uint64_t n_units, const uint64_t *__restrict__ chunk_base,
                                                    uint32_t chunk_len, ChunkState *st, UnitSched *units, SchedQueue *q, uint32_t *items) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const uint64_t g0 = chunk_base[u];
    const uint32_t C = (uint32_t)(chunk_base[u + 1] - g0);
    if (C == 0) return;
    const uint32_t n = (uint32_t)(unit_off[u + 1] - unit_off[u]);
    const bool unit_final = sched_advance(units[u], n, chunk_len, C, st + g0, [&](uint32_t cc) {
        __threadfence();
        const uint32_t slot = atomicAdd(&q->tail, 1u);
        if (slot < sync_load(&q->capacity)) sync_store(&items[slot], (uint32_t)(g0 + cc) + 1u);
        else sync_store(&q->finished, 2u);
    });
    if (unit_final && atomicAdd(&q->units_done, 1u) + 1u == sync_load(&q->live_units)) {
        __threadfence();
        atomicCAS(&q->finished, 0u, 1u);
    }
}

// Workers: one warp = one worker.  Worker w starts with queue slot w; afterwards it pops the oldest queued chunk,
// claims it, runs its chain (the same spec_chain as the round scheme) and advances the unit.  A worker that finds
// the queue empty EXITS: every later push is made by a worker that is still alive and pops right afterwards, so
// nothing is ever stranded, and the SM slots of a draining kernel become free for other streams.
// The screened engine (mdb_fit_screen.cuh) needs far fewer registers per step than the exact one, whose code it only calls
// on the rare fits it cannot decide (that code spills under this bound, which is the price of those fits): four blocks per SM (measured: 13.2 ms against 15.5 ms with five, whose 96 registers make the step itself spill).
using WarpFitScreen = WarpFitScreenT<MDB_FIT_POINTS_PER_LANE>;
#ifndef MDB_SCREEN_MIN_BLOCKS
#define MDB_SCREEN_MIN_BLOCKS 4
#endif
// info: the units as the pre-pass (k_lanes_units, k_lanes_regular) saw them, or nullptr (only the screened engine looks)
template <typename Fit, int MIN_BLOCKS>
__global__ void __launch_bounds__(CHAIN_WARPS * 32, MIN_BLOCKS) k_spec_async(const int64_t *__restrict__ ts, const float *__restrict__ values,
                                                                 const uint64_t *__restrict__ unit_off, const uint8_t *__restrict__ eb_kind,
                                                                 const float *__restrict__ eb_value, const uint64_t *__restrict__ chunk_base,
                                                                 const uint32_t *__restrict__ chunk_unit, uint32_t chunk_len, ChunkState *st,
                                                                 FittedModel *lists, const uint64_t *__restrict__ list_base,
                                                                 const uint32_t *__restrict__ list_cap, UnitSched *units, SchedQueue *q, uint32_t *items,
                                                                 uint32_t n_initial, uint32_t n_chunks, const LaneUnit *__restrict__ info) {
    __shared__ double smem[CHAIN_WARPS][Fit::SMEM_DOUBLES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    bool first = true, initial_phase = true;
    uint32_t carry = 0; // (lane 0) the chunk this worker's own sched_advance queued last: taken next, without the queue
    while (true) {
        uint32_t item = 0;
        if (lane == 0) {
            const uint32_t w = blockIdx.x * CHAIN_WARPS + warp;
            if (carry) { // stay with the unit: its frontier advances one chunk after the other, and a queue shared by
                item = carry; // every worker serialises at about a microsecond per item
                carry = 0;
            } else if (first && w < n_initial) { // (init_head starts at n_initial)
                item = sync_load(&items[w]);
            } else {
                // the pushed part first: those are frontier chunks, the critical path of their unit.  Popped exactly
                // (a worker must never hold a ticket for a slot that nobody may ever fill).
                while (true) {
                    const uint32_t h = sync_load(&q->head), t = min(sync_load(&q->tail), sync_load(&q->capacity));
                    if (h >= t) break;
                    if (atomicCAS(&q->head, h, h + 1u) == h) {
                        while ((item = sync_load(&items[h])) == 0) __nanosleep(20); // its pusher has the slot and is writing it
                        break;
                    }
                }
                // then the pre-filled part: a plain ticket counter (overshooting it is harmless)
                if (item == 0 && initial_phase) {
                    const uint32_t h = atomicAdd(&q->init_head, 1u);
                    if (h < n_chunks) item = sync_load(&items[h]);
                    else initial_phase = false;
                }
                // (both empty: this worker is surplus and leaves)
            }
        }
        first = false;
        item = __shfl_sync(FULL_MASK, item, 0);
        if (item == 0) return;
        const uint64_t g = item - 1;
        const uint32_t u = chunk_unit[g];
        const uint64_t g0 = chunk_base[u];
        const uint32_t c = (uint32_t)(g - g0), C = (uint32_t)(chunk_base[u + 1] - g0);
        int run = 1;
        if (lane == 0) {
            while (true) { // claim: PH_QUEUED -> PH_RUNNING (PH_LOCKED: sched_advance is re-aiming it right now)
                const uint32_t old = atomicCAS(&st[g].phase, PH_QUEUED, PH_RUNNING);
                if (old == PH_QUEUED) break;
                if (old != PH_LOCKED) { run = 0; break; }
                __nanosleep(50);
            }
            if (run && c < sync_load(&units[u].next_c)) { // the frontier is already past it: a model spans this chunk
                atomicExch(&st[g].phase, PH_DONE);
                run = 0;
            }
            __threadfence();
        }
        run = __shfl_sync(FULL_MASK, run, 0);
        if (!run) continue;
        const uint64_t a = unit_off[u];
        const uint32_t n = (uint32_t)(unit_off[u + 1] - a);
        const uint32_t chunk_start = c * chunk_len;
        const uint32_t chunk_end = (uint64_t)chunk_start + chunk_len < n ? chunk_start + chunk_len : n;
        ChunkState s = load_shared_record(st + g);
        ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
        Fit fitter(eb, ts + a, values + a, n, smem[warp], info ? info + u : nullptr);
#ifdef MDB_FIT_COUNTERS
        const long long tc0 = clock64();
#endif
        spec_chain(fitter, (uint32_t)lane, 32u, n, chunk_end, chunk_len, s, lists + list_base[g], list_cap[g] / 2);
#ifdef MDB_FIT_COUNTERS
        const long long tc1 = clock64();
        if (lane == 0) {
            atomicAdd(&g_fit_counters[13], 1ull);
            atomicAdd(&g_fit_counters[14], (unsigned long long)(tc1 - tc0));
        }
#endif
        __threadfence(); // this lane's list writes, before lane 0 publishes the chunk
        __syncwarp();
        if (lane == 0) {
            s.phase = PH_RUNNING;
            st[g] = s;
            __threadfence();
            atomicExch(&st[g].phase, PH_DONE);
            const bool unit_final = sched_advance(units[u], n, chunk_len, C, st + g0, [&](uint32_t cc) {
                __threadfence();
                if (!carry) {
                    carry = (uint32_t)(g0 + cc) + 1u;
                    return;
                }
                const uint32_t slot = atomicAdd(&q->tail, 1u);
                if (slot < sync_load(&q->capacity)) sync_store(&items[slot], (uint32_t)(g0 + cc) + 1u);
                else sync_store(&q->finished, 2u); // cannot happen (a chunk is queued at most three times); stops the workers
            });
            if (unit_final && atomicAdd(&q->units_done, 1u) + 1u == sync_load(&q->live_units)) {
                __threadfence();
                atomicCAS(&q->finished, 0u, 1u);
            }
#ifdef MDB_FIT_COUNTERS
            atomicAdd(&g_fit_counters[15], (unsigned long long)(clock64() - tc1));
#endif
        }
        __syncwarp();
    }
}

// ---- completing the pending Swing models (swing_finish) ---------------------------------------------------
// The two MSE sums of swing.rs:212-228 must be accumulated in point order, one rounding per point, so a model
// is a serial chain of additions -- but the models are independent of each other.  One LANE therefore sums one
// model (32 chains advance per instruction), reading its points with 16-byte loads; on a regular unit the time
// differences are k * interval and the timestamps are not read at all.  Models longer than
// SWING_FINISH_LONG points would leave the other 31 lanes waiting: those are summed by the whole warp instead
// (coalesced loads, terms in parallel, the additions chained through shared memory).
#include "mdb_swing_sums.cuh" // swing_sums_one_lane
#ifndef MDB_SWING_FINISH_LONG
#define MDB_SWING_FINISH_LONG 4096
#endif
constexpr uint32_t SWING_FINISH_LONG = MDB_SWING_FINISH_LONG;

// (Used for LONG models only.  For the short ones of k_swing_finish four accumulators per lane were measured on B200 and gained
// nothing -- 1.5 ms for the order-free pass against 1.95 ms for the serial one per 10^9 points: a lane per model is bound by
// its scattered 16-byte loads, not by the two addition chains -- while every model that failed the test paid twice.)
// When are the two error sums independent of the ORDER of their additions?  When every partial sum, in any order, is exactly
// representable -- then no addition rounds.  On a regular unit a numerator term is (v_k - v0) * dt_k with dt_k = k * delta an
// integer <= D = K * delta (K = points after the first), and every value is a multiple of q = 2^(emin - 23), emin the smallest
// exponent among the model's non-zero values: so v_k - v0 is a multiple of q below 2^(emax + 2), the product is a multiple of q
// below 2^(emax + 2) D, and any sum of up to K of them stays below 2^(bits(K) + emax + 2 + bits(D)): all of these are exact
// doubles iff that exponent exceeds emin - 23 by at most 53.  A denominator term is the integer dt_k^2 <= D^2: sums of up to K
// of them are exact iff bits(K) + 2 bits(D) <= 53.  (umax / umin1: the largest |value| and the smallest non-zero |value| minus
// one as bit patterns, v0 included; zeros are multiples of anything.)
__device__ __forceinline__ bool swing_sums_order_free(unsigned umax, unsigned umin1, uint32_t K, double delta_d) {
    if (umax >= 0x7f800000u) return false; // (a NaN or an infinity: never in a pending model, but then nothing is exact)
    const double D = __dmul_rn((double)K, delta_d);
    if (!(D >= 1.0)) return false;
    const int bits_d = ((__double2hiint(D) >> 20) & 0x7ff) - 1022; // floor(log2 D) + 1 (a rounded-up D only errs on the safe side)
    const int bits_k = 32 - __clz((int)K);
    if (bits_k + 2 * bits_d > 53) return false;
    if (umax == 0u) return true; // every value is zero: every term is skipped
    const int emax = max((int)(umax >> 23), 1) - 127, emin = max((int)((umin1 + 1u) >> 23), 1) - 127;
    return bits_k + emax + 2 + bits_d - (emin - 23) <= 53;
}

__global__ void __launch_bounds__(128) k_swing_finish(const int64_t *__restrict__ ts, const float *__restrict__ values,
                                                      const uint64_t *__restrict__ unit_off, const uint32_t *__restrict__ chunk_unit,
                                                      uint64_t n_chunks, const ChunkState *st, FittedModel *lists,
                                                      const uint64_t *__restrict__ list_base, const uint32_t *__restrict__ list_cap,
                                                      const uint8_t *__restrict__ unit_irregular, uint2 *long_models, unsigned int *n_long) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t g = (uint64_t)blockIdx.x * 4 + warp;
    if (g >= n_chunks) return;
    const ChunkState s = st[g];
    if (s.skipped || s.n_models == 0) return;
    const uint32_t u = chunk_unit[g];
    const uint64_t a = unit_off[u];
    const uint32_t n = (uint32_t)(unit_off[u + 1] - a);
    const int64_t *uts = ts + a;
    const float *uval = values + a;
    const int64_t delta0 = n >= 2 ? uts[1] - uts[0] : 0;
    const bool regular = !unit_irregular[u] && delta0 >= 0 && delta0 < (1ll << 31);
    const double delta_d = (double)delta0;
    FittedModel *list = lists + list_base[g] + (size_t)s.buf * (list_cap[g] / 2);
    for (uint32_t k0 = 0; k0 < s.n_models; k0 += 32) {
        const uint32_t k = k0 + (uint32_t)lane;
        FittedModel m;
        bool mine = false;
        if (k < s.n_models) {
            m = list[k];
            mine = m.pending != 0;
        }
        const bool is_long = mine && (m.end_index - m.start_index) >= SWING_FINISH_LONG;
        if (mine && !is_long) {
            double num, den;
            swing_sums_one_lane(uts, uval, regular, delta_d, m.start_index, m.end_index, num, den);
            swing_finish_from_sums(m, num, den, uts, uval);
            list[k] = m;
        }
        if (is_long) { // summed by k_swing_finish_long: one block per long model
            const unsigned int slot = atomicAdd(n_long, 1u);
            long_models[slot] = make_uint2((uint32_t)g, k);
        }
    }
}

// A LONG pending Swing model whose sums are order free (swing_sums_order_free): one block adds the terms in parallel --
// coalesced loads, a partial sum per thread, a shuffle tree -- instead of a chain of 10^6 dependent additions.  Models whose
// sums are not order free (or units with irregular timestamps) are left pending for k_swing_finish_long below.
__global__ void __launch_bounds__(256) k_swing_finish_long_par(const int64_t *__restrict__ ts, const float *__restrict__ values,
                                                               const uint64_t *__restrict__ unit_off, const uint32_t *__restrict__ chunk_unit,
                                                               const ChunkState *st, FittedModel *lists, const uint64_t *__restrict__ list_base,
                                                               const uint32_t *__restrict__ list_cap, const uint8_t *__restrict__ unit_irregular,
                                                               const uint2 *__restrict__ long_models, const unsigned int *n_long) {
    __shared__ double s_num[8], s_den[8];
    __shared__ unsigned s_max[8], s_min[8];
    for (unsigned int item_i = blockIdx.x; item_i < *n_long; item_i += gridDim.x) {
        const uint2 item = long_models[item_i];
        const uint64_t g = item.x;
        const ChunkState s = st[g];
        const uint32_t u = chunk_unit[g];
        const uint64_t a = unit_off[u];
        const uint32_t n = (uint32_t)(unit_off[u + 1] - a);
        const int64_t *uts = ts + a;
        const float *uval = values + a;
        const int64_t delta0 = n >= 2 ? uts[1] - uts[0] : 0;
        const bool regular = !unit_irregular[u] && delta0 >= 0 && delta0 < (1ll << 31);
        if (!regular) continue; // (uniform: the whole block skips the model)
        FittedModel *slot = lists + list_base[g] + (size_t)s.buf * (list_cap[g] / 2) + item.y;
        FittedModel m = *slot;
        const double delta_d = (double)delta0, v0 = (double)uval[m.start_index];
        // what is known before a value is read: the denominator's condition, and the numerator's with the narrowest
        // conceivable values (a spread of 16 bits between the largest difference and the quantum).  With microsecond
        // timestamps a millisecond apart a model of 4096 points already fails it, and the block leaves at no cost.
        {
            const uint32_t K = m.end_index - m.start_index;
            const double D = __dmul_rn((double)K, delta_d);
            const int bits_d = D >= 1.0 ? ((__double2hiint(D) >> 20) & 0x7ff) - 1022 : 64, bits_k = 32 - __clz((int)K);
            if (bits_k + 2 * bits_d > 53 || bits_k + bits_d + 16 > 53) continue; // (uniform)
        }
        double num = 0.0, den = 0.0;
        unsigned umax = 0u, umin1 = 0xffffffffu;
        for (uint32_t i = m.start_index + threadIdx.x; i <= m.end_index; i += blockDim.x) {
            const float vf = uval[i];
            const unsigned b = __float_as_uint(vf) & 0x7fffffffu;
            umax = max(umax, b);
            umin1 = min(umin1, b - 1u);
            if (i >= m.start_index + 2) { // the first two points add no term
                const double v = (double)vf, dt = __dmul_rn((double)(i - m.start_index), delta_d);
                if (!(v0 == v)) {
                    num = __dadd_rn(num, __dmul_rn(__dsub_rn(v, v0), dt));
                    den = __dadd_rn(den, __dmul_rn(dt, dt));
                }
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            num = __dadd_rn(num, __shfl_xor_sync(0xffffffffu, num, d));
            den = __dadd_rn(den, __shfl_xor_sync(0xffffffffu, den, d));
        }
        umax = __reduce_max_sync(0xffffffffu, umax);
        umin1 = __reduce_min_sync(0xffffffffu, umin1);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads(); // (the previous model's partials have been read)
        if (lane == 0) { s_num[warp] = num; s_den[warp] = den; s_max[warp] = umax; s_min[warp] = umin1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; w++) {
                num = __dadd_rn(num, s_num[w]);
                den = __dadd_rn(den, s_den[w]);
                umax = max(umax, s_max[w]);
                umin1 = min(umin1, s_min[w]);
            }
            if (swing_sums_order_free(umax, umin1, m.end_index - m.start_index, delta_d)) {
                swing_finish_from_sums(m, num, den, uts, uval);
                *slot = m; // (pending is cleared: k_swing_finish_long skips it)
            }
        }
    }
}

// A LONG pending Swing model (>= SWING_FINISH_LONG points, up to a whole unit): its two sums are chains of 10^6 dependent
// f64 additions, and a dependent DADD takes 8.3 cycles on B200 (tools/microbench/fp64.cu), so the floor is ~4.4 ms per
// 10^6 points however the work is arranged.  To get near it the chain must be nothing but the additions: warp 1 of the block
// PRODUCES the terms (loads, conversions, two multiplications per point) for chunk c + 1 into one half of a shared buffer
// while warp 0 CONSUMES chunk c from the other half -- 16-byte shared loads hoisted ahead of the chain, two DADD chains
// interleaved -- one __syncthreads per 512 points.  (The former arrangement, one warp doing both in turn, took 31 cycles per
// point.)  Order and operands of every addition are the reference's (swing.rs:180-193, 212-228).
constexpr int SWING_LONG_CHUNK = 512;

__global__ void __launch_bounds__(64) k_swing_finish_long(const int64_t *__restrict__ ts, const float *__restrict__ values,
                                                          const uint64_t *__restrict__ unit_off, const uint32_t *__restrict__ chunk_unit, const ChunkState *st,
                                                          FittedModel *lists, const uint64_t *__restrict__ list_base, const uint32_t *__restrict__ list_cap,
                                                          const uint2 *__restrict__ long_models, const unsigned int *n_long) {
    __shared__ __align__(16) double sx[2][SWING_LONG_CHUNK], sy[2][SWING_LONG_CHUNK];
    if (blockIdx.x >= *n_long) return;
    const uint2 item = long_models[blockIdx.x];
    const uint64_t g = item.x;
    const ChunkState s = st[g];
    const uint32_t u = chunk_unit[g];
    const int64_t *uts = ts + unit_off[u];
    const float *uval = values + unit_off[u];
    FittedModel *slot = lists + list_base[g] + (size_t)s.buf * (list_cap[g] / 2) + item.y;
    FittedModel m = *slot;
    if (!m.pending) return; // its sums were order free: k_swing_finish_long_par has added them
    const int64_t t0 = uts[m.start_index];
    const double v0 = (double)uval[m.start_index];
    const uint32_t first = m.start_index + 2, last = m.end_index; // the first two points add no term
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_terms = last >= first ? last - first + 1 : 0;
    const uint32_t n_chunks = (n_terms + SWING_LONG_CHUNK - 1) / SWING_LONG_CHUNK;
    double num = 0.0, den = 0.0;
    auto produce = [&](uint32_t c) { // (warp 1) terms of chunk c into half c & 1; entries past the model's end are not read back
        const uint32_t base = first + c * SWING_LONG_CHUNK;
        constexpr int PER_LANE = SWING_LONG_CHUNK / 32;
        int64_t t[PER_LANE];
        float v[PER_LANE];
#pragma unroll
        for (int q = 0; q < PER_LANE; q++) { // every load of the chunk is in flight before the first one is used
            const uint32_t i = base + (uint32_t)(q * 32 + lane);
            t[q] = i <= last ? uts[i] : t0;
            v[q] = i <= last ? uval[i] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < PER_LANE; q++) {
            double x, y;
            swing_mse_terms(t0, v0, t[q], (double)v[q], x, y);
            sx[c & 1][q * 32 + lane] = x;
            sy[c & 1][q * 32 + lane] = y;
        }
    };
    if (warp == 1 && n_chunks) produce(0);
    __syncthreads();
    for (uint32_t c = 0; c < n_chunks; c++) {
        if (warp == 1) {
            if (c + 1 < n_chunks) produce(c + 1);
        } else {
            const int cnt = (int)min((uint32_t)SWING_LONG_CHUNK, n_terms - c * SWING_LONG_CHUNK);
            const double2 *bx = reinterpret_cast<const double2 *>(sx[c & 1]), *by = reinterpret_cast<const double2 *>(sy[c & 1]);
            int j = 0;
#pragma unroll 8
            for (; j + 2 <= cnt; j += 2) { // every lane runs the same chain (uniform shared loads are broadcasts)
                const double2 xx = bx[j >> 1], yy = by[j >> 1];
                num = __dadd_rn(__dadd_rn(num, xx.x), xx.y);
                den = __dadd_rn(__dadd_rn(den, yy.x), yy.y);
            }
            if (j < cnt) { // (adding a padding zero could flip the sign of a zero sum)
                num = __dadd_rn(num, sx[c & 1][j]);
                den = __dadd_rn(den, sy[c & 1][j]);
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
        swing_finish_from_sums(m, num, den, uts, uval);
        if (lane == 0) *slot = m;
    }
}

__global__ void __launch_bounds__(128) k_spec_propagate(const uint64_t *unit_off, uint64_t n_units, const uint64_t *chunk_base, uint32_t chunk_len,
                                                        ChunkState *st, int allow_optimistic, uint2 *unit_resume, uint32_t *worklist,
                                                        CompressCounters *counters) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    uint32_t n = (uint32_t)(unit_off[u + 1] - unit_off[u]);
    uint64_t g0 = chunk_base[u];
    uint32_t C = (uint32_t)(chunk_base[u + 1] - g0);
    if (C == 0) return;
    uint2 r = unit_resume[u]; // (chunk, entry) up to which this unit is final
    if (r.x >= C) return;
    spec_propagate_unit(n, chunk_len, C, st + g0, allow_optimistic != 0, r.x, r.y, [&](uint32_t c) {
        worklist[atomicAdd(&counters->dirty, 1u)] = (uint32_t)(g0 + c); // the next round runs exactly these chunks
    });
    unit_resume[u] = r;
}

__global__ void __launch_bounds__(128) k_spec_finalize(const uint64_t *unit_off, uint64_t n_units, const uint64_t *chunk_base, uint32_t chunk_len,
                                                       ChunkState *st, uint8_t *unit_irregular) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    uint32_t n = (uint32_t)(unit_off[u + 1] - unit_off[u]);
    uint64_t g0 = chunk_base[u];
    uint32_t C = (uint32_t)(chunk_base[u + 1] - g0);
    uint8_t irregular = 0;
    if (C) spec_finalize_unit(n, chunk_len, C, st + g0, irregular);
    unit_irregular[u] = irregular;
}

__global__ void __launch_bounds__(128) k_spec_count_rows(ChunkState *st, uint64_t n_chunks, const FittedModel *lists, const uint64_t *list_base,
                                                         const uint32_t *list_cap, uint32_t *rows) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks) return;
    const ChunkState s = st[g];
    rows[g] = spec_count_rows(s, lists + list_base[g] + (size_t)s.buf * (list_cap[g] / 2));
}

__global__ void __launch_bounds__(256) k_unit_seg_off(const uint64_t *chunk_base, uint64_t n_units, const uint64_t *row_base, uint64_t *unit_seg_off) {
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u > n_units) return;
    unit_seg_off[u] = row_base[chunk_base[u]]; // chunk_base[n_units] == n_chunks, row_base[n_chunks] == total rows
}

// One warp per chunk, one lane per accepted model (spec_records in mdb_compress.cuh is the one-thread form).  A model
// yields one row, or two when more than 255 residual points follow it (compression.rs:310-362); an exclusive scan over
// the lanes gives every model its row position, so the rows come out in the chain's order.
__global__ void __launch_bounds__(128) k_spec_records(const int64_t *__restrict__ ts, const float *__restrict__ values, const uint64_t *__restrict__ unit_off,
                                                      const uint8_t *__restrict__ eb_kind, const float *__restrict__ eb_value,
                                                      const uint32_t *__restrict__ chunk_unit, uint64_t n_chunks, const ChunkState *st,
                                                      const FittedModel *lists, const uint64_t *__restrict__ list_base, const uint32_t *__restrict__ list_cap,
                                                      const uint8_t *__restrict__ unit_irregular, const uint64_t *__restrict__ row_base, SegRecord *recs,
                                                      uint32_t *row_unit, uint32_t *wide_rows, unsigned int *n_wide) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t g = (uint64_t)blockIdx.x * 4 + warp;
    if (g >= n_chunks) return;
    const ChunkState s = st[g];
    const uint64_t r0 = row_base[g];
    if (row_base[g + 1] == r0) return; // (skipped chunks have no rows)
    const uint32_t u = chunk_unit[g];
    const uint64_t a = unit_off[u];
    const ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
    const int64_t *uts = ts + a;
    const float *uval = values + a;
    const bool regular = unit_irregular[u] == 0;
    const FittedModel *list = lists + list_base[g] + (size_t)s.buf * (list_cap[g] / 2);
    auto publish = [&](uint64_t row) { // long MacaqueV rows go on to the warp kernels below
        row_unit[row] = u;
        if (recs[row].wide) wide_rows[atomicAdd(n_wide, 1u)] = (uint32_t)row;
    };
    uint32_t r = 0; // rows written so far (uniform)
    if (s.lead_end != IDX_NONE) { // compression.rs:350-361: leading residuals of the unit
        if (lane == 0) {
            record_macaque_v_segment(eb, 0, s.lead_end, uts, uval, regular, recs[r0], WIDE_ENCODE_MIN);
            publish(r0);
        }
        r = 1;
    }
    for (uint32_t k0 = 0; k0 < s.n_models; k0 += 32) {
        const uint32_t k = k0 + (uint32_t)lane;
        const bool have = k < s.n_models;
        FittedModel m;
        uint32_t res_end = 0, n_rows = 0;
        if (have) {
            m = list[k];
            const uint32_t next_start = k + 1 < s.n_models ? list[k + 1].start_index : s.next_start;
            res_end = next_start - 1;
            n_rows = (res_end - m.end_index <= RESIDUAL_VALUES_MAX_LENGTH) ? 1u : 2u;
        }
        uint32_t incl = n_rows;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        if (have) {
            const uint64_t row = r0 + r + (incl - n_rows);
            store_segments(eb, true, m, res_end, uts, uval, regular, recs + row, WIDE_ENCODE_MIN);
            row_unit[row] = u;
            if (n_rows == 2) publish(row + 1);
        }
        r += __shfl_sync(FULL_MASK, incl, 31);
    }
}

// ---- long MacaqueV rows encoded by a whole warp: warp_macaque_v_encode, WarpCodeCounter, WarpCodeWriter (mdb_macaque_warp.cuh)

// A long MacaqueV row is encoded ONCE (round 2; before, once on a bit counter to size it and once more to write it): its bytes
// go to a staging area whose slot for the row follows from where the row's points lie in the input -- a code is at most 45
// bits, so the row of the points [p, q] needs fewer than 6 (q - p + 1) bytes (rows here have at least 256 values) and gets the
// bytes from round_up(6 p, 16) on -- and k_copy_macaque_warp moves them to their final offset once that is known.
__device__ __forceinline__ uint64_t wide_slot(uint64_t first_point) { return (6 * first_point + 15) & ~(uint64_t)15; }

__global__ void __launch_bounds__(WIDE_WARPS * 32) k_records_macaque_warp(const float *__restrict__ values, const uint64_t *__restrict__ unit_off,
                                                                           const uint8_t *__restrict__ eb_kind, const float *__restrict__ eb_value,
                                                                           const uint32_t *__restrict__ row_unit, const uint32_t *__restrict__ wide_rows,
                                                                           const unsigned int *n_wide_ptr, SegRecord *recs, uint8_t *wide_tmp) {
    __shared__ uint32_t stage[WIDE_WARPS][STAGE_WORDS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_wide = *n_wide_ptr;
    for (uint32_t w = blockIdx.x * WIDE_WARPS + warp; w < n_wide; w += gridDim.x * WIDE_WARPS) {
        const uint32_t r = wide_rows[w];
        const uint32_t u = row_unit[r];
        const ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
        const uint32_t lo = recs[r].start_index, hi = recs[r].res_end_index;
        uint8_t *slot = wide_tmp + wide_slot(unit_off[u] + lo);
        WarpCodeWriter sink;
        sink.init(slot, stage[warp], lane);
        sink.word_stores = true;
        float mn, mx;
        warp_macaque_v_encode(eb, values + unit_off[u], lo, hi, sink, lane, mn, mx);
        sink.finish();
        if (lane == 0) {
            recs[r].val_len = (uint32_t)(sink.out - slot);
            recs[r].min_value = mn;
            recs[r].max_value = mx;
        }
        __syncwarp();
    }
}

// The staged bytes of every long MacaqueV row to their final place (arbitrary byte alignment): a block per row (a row can be
// megabytes: a lossless series is one row), whole words assembled from two aligned source words by a funnel shift.
constexpr int COPY_THREADS = 256;
__global__ void __launch_bounds__(COPY_THREADS) k_copy_macaque_warp(const uint64_t *__restrict__ unit_off, const uint32_t *__restrict__ row_unit,
                                                                        const uint32_t *__restrict__ wide_rows, const unsigned int *n_wide_ptr,
                                                                        const SegRecord *__restrict__ recs, const uint8_t *__restrict__ wide_tmp,
                                                                        const uint64_t *__restrict__ val_off, uint8_t *val_data) {
    const uint32_t lane = threadIdx.x; // (of the block)
    const uint32_t n_wide = *n_wide_ptr;
    for (uint32_t w = blockIdx.x; w < n_wide; w += gridDim.x) {
        const uint32_t r = wide_rows[w];
        const uint8_t *src = wide_tmp + wide_slot(unit_off[row_unit[r]] + recs[r].start_index);
        uint8_t *dst = val_data + val_off[r];
        const uint32_t n = recs[r].val_len;
        const uint32_t head = min(n, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3)); // bytes up to the first aligned word of dst
        if (lane < head) dst[lane] = src[lane];
        const uint32_t n_words = (n - head) >> 2;
        const uint32_t *src_w = reinterpret_cast<const uint32_t *>(src); // (16-byte aligned)
        uint32_t *dst_w = reinterpret_cast<uint32_t *>(dst + head);
        const uint32_t shift = 8 * (head & 3), first = head >> 2;         // dst word i = source bytes head + 4 i .. + 3
        for (uint32_t i = lane; i < n_words; i += COPY_THREADS) {
            const uint32_t a = src_w[first + i], b = shift ? src_w[first + i + 1] : 0u; // (the slot has room past the row's bytes)
            dst_w[i] = __funnelshift_r(a, b, shift);
        }
        for (uint32_t i = head + 4 * n_words + lane; i < n; i += COPY_THREADS) dst[i] = src[i];
    }
}


// One thread per row: metadata columns and the byte lengths of the three binary columns.
__global__ void __launch_bounds__(256) k_compress_gather(const int64_t *__restrict__ ts, const uint64_t *__restrict__ unit_off,
                                                         const SegRecord *__restrict__ recs, const uint32_t *__restrict__ row_unit, uint64_t n_rows,
                                                         int8_t *model_type_id, int64_t *start_time, int64_t *end_time, float *min_value,
                                                         float *max_value, uint32_t *ts_len, uint32_t *val_len, uint32_t *res_len) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const SegRecord rec = recs[r];
    const int64_t *uts = ts + unit_off[row_unit[r]];
    model_type_id[r] = rec.model_type_id;
    start_time[r] = uts[rec.start_index];
    end_time[r] = uts[rec.res_end_index];
    min_value[r] = rec.min_value;
    max_value[r] = rec.max_value;
    ts_len[r] = rec.ts_len;
    val_len[r] = rec.val_len;
    res_len[r] = rec.res_len;
}

__global__ void __launch_bounds__(64) k_compress_emit(const int64_t *__restrict__ ts, const float *__restrict__ values, const uint64_t *__restrict__ unit_off,
                                                      const uint8_t *__restrict__ eb_kind, const float *__restrict__ eb_value,
                                                      const SegRecord *__restrict__ recs, const uint32_t *__restrict__ row_unit, uint64_t n_rows,
                                                      const uint64_t *__restrict__ ts_off, uint8_t *ts_data, const uint64_t *__restrict__ val_off,
                                                      uint8_t *val_data, const uint64_t *__restrict__ res_off, uint8_t *res_data) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    uint32_t u = row_unit[r];
    const SegRecord rec = recs[r];
    ErrorBound eb = make_error_bound(eb_kind[u], eb_value[u]);
    uint64_t a = unit_off[u];
    compress_emit_segment(eb, rec, ts + a, values + a, ts_data + ts_off[r], val_data + val_off[r], res_data + res_off[r]);
}

// Chunk length.  With one warp per chain, ~128 chains per SM already saturate the machine in round 0;
// longer chunks mean fewer fixpoint rounds for units whose chains do not re-synchronise (measured on
// 4e8 points: 4096 -> 140 ms, 16384 -> 74 ms, 65536 -> 81 ms).  The one-thread engine wants 32x more
// chains (one per lane).  Within [4096, 65536] points.
// With one lane per chain every chain is a thread, and the chains of a call should fill the resident lanes ONCE: a lane that
// finishes its chunk can only start another whole chunk, so 1.3 waves of chunks take as long as 2.  The chunk length is
// therefore (points / resident lanes), allowing for one partial chunk per unit; within [4096, 65536] points (a speculative
// lane spends a warm-up of a few thousand points before its chunk, LaneChain), in k equal waves when the call is that big.
// With units enough to occupy the lanes by themselves, chunks are as long as they get: no speculation, no warm-up.
static uint32_t choose_chunk_len(const mdbcu_context *ctx, uint64_t n_points, uint64_t n_units, uint64_t resident_lanes) {
    if (ctx->chunk_len_override) return ctx->chunk_len_override;
    if (resident_lanes) {
        if (n_units >= resident_lanes / 2) return 65536;
        for (uint64_t waves = 1;; waves++) {
            const uint64_t chains = waves * resident_lanes - n_units;
            const uint64_t len = ((n_points + chains - 1) / chains + 63) / 64 * 64;
            if (len <= 65536) return (uint32_t)std::max<uint64_t>(len, 4096);
        }
    }
    uint64_t target_chains = (uint64_t)ctx->sm_count * 128 * (ctx->fit_mode == 1 ? 32 : 1);
    uint64_t len = n_points / target_chains;
    uint32_t l = 4096;
    while (l < len && l < 65536) l <<= 1;
    return l;
}

extern "C" {

int mdbcu_context_set_chunk_len(mdbcu_context *ctx, uint32_t chunk_len) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (chunk_len != 0 && chunk_len < 8) return fail("chunk_len must be 0 (automatic) or >= 8");
    ctx->chunk_len_override = chunk_len;
    return MDBCU_SUCCESS;
}

uint32_t mdbcu_context_last_compress_rounds(const mdbcu_context *ctx) { return ctx ? ctx->last_rounds : 0; }

int mdbcu_context_set_lane_warmup(mdbcu_context *ctx, uint32_t points) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    ctx->lane_rounds_by_lanes = (points & 0x80000000u) != 0; // (tuning switch in the top bit: repair rounds by lanes instead of warps)
    ctx->lane_warmup = points & 0x7FFFFFFFu;
    return MDBCU_SUCCESS;
}

// Tuning knobs by name (tests and benchmarks; results never depend on them).
int mdbcu_context_set_option(mdbcu_context *ctx, const char *name, int64_t value) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (!name) return fail("set_option: name is null");
    const std::string n(name);
    if (n == "grid_tma_stores") ctx->grid_tma_stores = value != 0;
    else if (n == "grid_tile_scan") ctx->grid_tile_scan = value != 0;
    else if (n == "overlap_regular_check") ctx->overlap_regular_check = value != 0;
    else if (n == "lane_rows_min") ctx->lane_rows_min = (uint32_t)std::max<int64_t>(1, value);
    else if (n == "fit_wide") ctx->fit_wide = value != 0;
    else if (n == "block_row_warps") ctx->block_row_warps = (int)value;
    else if (n == "block_row_min") ctx->block_row_min = (uint32_t)std::max<int64_t>(512, value);
    else if (n == "lane_warmup") ctx->lane_warmup = (uint32_t)std::max<int64_t>(0, value);
    else if (n == "lane_rounds_by_lanes") ctx->lane_rounds_by_lanes = value != 0;
    else if (n == "chunk_len") return mdbcu_context_set_chunk_len(ctx, (uint32_t)value);
    else if (n == "fit_engine") return mdbcu_context_set_fit_engine(ctx, (int)value);
    else return fail("set_option: unknown option '" + n + "'");
    return MDBCU_SUCCESS;
}

int mdbcu_context_set_fit_engine(mdbcu_context *ctx, int engine) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (engine < 0 || engine > 5)
        return fail("fit engine must be 0 (automatic), 1 (one thread per chain, rounds), 2 (one warp per chain, rounds), 3 (one warp per chain, "
                    "asynchronous scheduling), 4 (one lane per chain, then 3 for the stitching) or 5 (3 with the screened fit)");
    ctx->fit_mode = engine;
    return MDBCU_SUCCESS;
}

int mdbcu_compress(mdbcu_context *ctx, mdbcu_space space, const int64_t *timestamps, const float *values, const uint64_t *unit_off,
                   uint64_t n_units, const uint8_t *eb_kind, const float *eb_value, mdbcu_segments **out) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (!out) return fail("compress: out is null");
    *out = nullptr;
    if (n_units > 0xFFFFFFF0ull) return fail("compress: more than 2^32 units");
    if (n_units && (!unit_off || !eb_kind || !eb_value)) return fail("compress: unit_off / eb_kind / eb_value is null");
    cudaStream_t s = ctx->stream;

    // total number of points: unit_off[0] .. unit_off[n_units]
    uint64_t first = 0, last = 0;
    if (n_units) {
        if (space == MDBCU_HOST) {
            first = unit_off[0];
            last = unit_off[n_units];
        } else {
            CUDA_TRY(post(ctx, 0, unit_off, 1));
            CUDA_TRY(post(ctx, 1, unit_off + n_units, 1));
            CUDA_TRY(sync_stream(ctx));
            first = ctx->mailbox[0];
            last = ctx->mailbox[1];
        }
        if (last < first) return fail("compress: unit_off is not monotone");
    }
    uint64_t n_points = last; // arrays are indexed by absolute unit_off values
    if (n_points && (!timestamps || !values)) return fail("compress: timestamps / values is null");

    DBuf<int64_t> ts_buf;
    DBuf<float> val_buf;
    DBuf<uint64_t> off_buf;
    DBuf<uint8_t> kind_buf;
    DBuf<float> ebv_buf;
    const int64_t *d_ts = timestamps;
    const float *d_val = values;
    const uint64_t *d_off = unit_off;
    const uint8_t *d_kind = eb_kind;
    const float *d_ebv = eb_value;
    if (space == MDBCU_HOST && n_units) {
        CUDA_TRY(upload(ctx, off_buf, unit_off, n_units + 1));
        CUDA_TRY(upload(ctx, kind_buf, eb_kind, n_units));
        CUDA_TRY(upload(ctx, ebv_buf, eb_value, n_units));
        CUDA_TRY(upload(ctx, ts_buf, timestamps, n_points));
        CUDA_TRY(upload(ctx, val_buf, values, n_points));
        d_ts = ts_buf.p; d_val = val_buf.p; d_off = off_buf.p; d_kind = kind_buf.p; d_ebv = ebv_buf.p;
    }

    mdbcu_segments *sg = new mdbcu_segments();
    sg->ctx = ctx;
    sg->n_units = n_units;
    auto bail = [&](int rc) {
        if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream); // (a check still running there reads buffers this scope frees)
        mdbcu_segments_free(sg);
        return rc;
    };
#define TRY_SG(expr)                                                                                \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fail(std::string(#expr) + ": " + cudaGetErrorName(e_) + ": " + cudaGetErrorString(e_)); \
            return bail(MDBCU_FAILURE);                                                             \
        }                                                                                           \
    } while (0)

    TRY_SG(cudaMallocAsync((void **)&sg->unit_seg_off, (n_units + 1) * sizeof(uint64_t), s));
    uint64_t S = 0;
    DBuf<SegRecord> recs;
    DBuf<uint32_t> row_unit, wide_rows; // wide_rows: long MacaqueV rows, sized and written by a whole warp
    DBuf<unsigned int> n_wide; // (two words: posted to the host as one 64-bit word)
    DBuf<uint8_t> wide_tmp;    // staging area of the long MacaqueV rows' bytes
    uint32_t n_wide_host = 0;
    TRY_SG(n_wide.alloc(2, s));
    TRY_SG(cudaMemsetAsync(n_wide.p, 0, 2 * sizeof(unsigned int), s));
    ctx->last_rounds = 0;

    if (n_units) {
        // ---- chunks
        // Engines (mdbcu_context_set_fit_engine).  Automatic = the screened cooperative fit (mdb_fit_screen.cuh) whenever a unit
        // has a lossy bound, else the exact cooperative fit; both under the asynchronous scheduler.  Measured on B200 per 10^9
        // points, chain kernel only (round 2): 1000 series x 10^6 points, 1 %: exact 17.8 ms, screened 10.7 ms; 100 000 series x
        // 10^4 points, 1 %: one lane per chain (mdb_fit_lanes.cuh, engine 4) 13.1 ms, screened 8.7 ms; 5 % bound (models of
        // thousands of points): exact 20.0 ms, screened 23.3 ms -- the one case the automatic choice loses.
        // The lanes (engine 4) are decided first, because their chunks are shorter.
        bool use_lanes = ctx->fit_mode == 4;
        // The screened engine needs the same per-unit facts as the lanes: regular timestamps, bound kind.
        bool use_screen = ctx->fit_mode == 5 || ctx->fit_mode == 0;
        DBuf<LaneUnit> lane_units;
        DBuf<unsigned int> lane_words; // [0..2] qualifying units per bound kind, [3] the chunk counter
        unsigned int kind_units[4] = {0, 0, 0, 0};
        if (use_lanes || use_screen) {
            TRY_SG(lane_units.alloc(n_units, s));
            TRY_SG(lane_words.alloc(4, s));
            TRY_SG(cudaMemsetAsync(lane_words.p, 0, 4 * sizeof(unsigned int), s));
            LAUNCH(ctx, k_lanes_units, div_up(n_units, 128), 128, 0, d_ts, d_off, n_units, d_kind, d_ebv, lane_units.p, lane_words.p);
            TRY_SG(post(ctx, 0, lane_words.p, 2));
            TRY_SG(sync_stream(ctx));
            std::memcpy(kind_units, ctx->mailbox, sizeof(kind_units));
            use_lanes = use_lanes && kind_units[0] + kind_units[1] + kind_units[2] > 0;
            use_screen = use_screen && kind_units[KIND_ABSOLUTE] + kind_units[KIND_RELATIVE] > 0; // (lossless fits are all ties: the exact engine's)
        }
        uint64_t resident_lanes = 0; // of the lane kernel (the smallest over the bound kinds present)
        int lane_blocks_per_sm[3] = {0, 0, 0};
        if (use_lanes) {
            const void *fns[3] = {(const void *)k_spec_lanes<KIND_LOSSLESS>, (const void *)k_spec_lanes<KIND_ABSOLUTE>, (const void *)k_spec_lanes<KIND_RELATIVE>};
            for (int kind = 0; kind < 3; kind++) {
                if (!kind_units[kind]) continue;
                TRY_SG(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lane_blocks_per_sm[kind], fns[kind], LANES_WARPS * 32, 0));
                if (lane_blocks_per_sm[kind] < 1) return bail(fail("compress: the lane kernel does not fit on this device"));
                const uint64_t r = (uint64_t)ctx->sm_count * lane_blocks_per_sm[kind] * LANES_WARPS * 32;
                resident_lanes = resident_lanes ? std::min(resident_lanes, r) : r;
            }
        }
        const uint32_t chunk_len = choose_chunk_len(ctx, n_points - first, n_units, resident_lanes);
        DBuf<Status> status;
        if (new_status(ctx, status)) return bail(MDBCU_FAILURE);
        DBuf<uint32_t> unit_chunks;
        DBuf<uint64_t> chunk_base;
        TRY_SG(unit_chunks.alloc(n_units, s));
        TRY_SG(chunk_base.alloc(n_units + 1, s));
        LAUNCH(ctx, k_unit_chunks, div_up(n_units, 256), 256, 0, d_off, n_units, d_kind, d_ebv, chunk_len, unit_chunks.p, status.p);
        if (exclusive_scan<uint32_t>(ctx, unit_chunks.p, n_units, chunk_base.p)) return bail(MDBCU_FAILURE);
        TRY_SG(post(ctx, 0, chunk_base.p + n_units, 1));
        Status h;
        if (read_status(ctx, status.p, h, "unit (bad unit_off or error bound)")) return bail(MDBCU_FAILURE);
        const uint64_t G = ctx->mailbox[0];
        if (G > 0xFFFFFFF0ull) return bail(fail("compress: too many chunks"));

        DBuf<ChunkState> st;
        DBuf<uint32_t> chunk_unit, list_cap, rows;
        DBuf<uint64_t> list_base, row_base;
        DBuf<uint8_t> unit_irregular;
        DBuf<CompressCounters> counters;
        TRY_SG(st.alloc(G, s));
        TRY_SG(chunk_unit.alloc(G, s));
        TRY_SG(list_cap.alloc(G, s));
        TRY_SG(list_base.alloc(G + 1, s));
        TRY_SG(rows.alloc(G, s));
        TRY_SG(row_base.alloc(G + 1, s));
        TRY_SG(unit_irregular.alloc(n_units, s));
        TRY_SG(counters.alloc(1, s));
        LAUNCH(ctx, k_spec_init, div_up(n_units, 128), 128, 0, d_off, n_units, chunk_base.p, chunk_len, st.p, chunk_unit.p, list_cap.p);
        if (exclusive_scan<uint32_t>(ctx, list_cap.p, G, list_base.p)) return bail(MDBCU_FAILURE);
        TRY_SG(post(ctx, 0, list_base.p + G, 1));
        TRY_SG(sync_stream(ctx));
        const uint64_t n_models_cap = ctx->mailbox[0];
        DBuf<FittedModel> lists;
        TRY_SG(lists.alloc(n_models_cap, s));

        const bool async_sched = ctx->fit_mode == 0 || ctx->fit_mode >= 3;
        // Every interval of every qualifying unit (coalesced, bandwidth bound: 1.1 ms per 10^9 points).  The screened chain
        // kernel does not read a timestamp and is nowhere near the memory bandwidth, so the check runs BESIDE it on a second
        // stream: the chains assume what lane_unit_init saw (first interval, last timestamp), and if the check then finds an
        // irregular interval inside such a unit -- which costs it its `ok` -- the chains are simply run again, with the exact
        // engine.  (With per-kernel profiling on, everything stays on one stream so that the kernels can be timed.)
        const bool overlap_check = use_screen && !use_lanes && G && !ctx->profiling && ctx->overlap_regular_check;
        if (overlap_check) {
            if (!ctx->aux_stream) {
                TRY_SG(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
                TRY_SG(cudaEventCreateWithFlags(&ctx->aux_ready, cudaEventDisableTiming));
                TRY_SG(cudaEventCreateWithFlags(&ctx->aux_done, cudaEventDisableTiming));
            }
            TRY_SG(cudaEventRecord(ctx->aux_ready, s));
            TRY_SG(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ready, 0));
            k_lanes_regular<<<(unsigned int)G, REGULAR_THREADS, 0, ctx->aux_stream>>>(d_ts, d_off, chunk_base.p, chunk_unit.p, chunk_len, lane_units.p);
            ctx->launches++;
            TRY_SG(cudaEventRecord(ctx->aux_done, ctx->aux_stream));
        } else if ((use_lanes || use_screen) && G) {
            LAUNCH(ctx, k_lanes_regular, (unsigned int)G, REGULAR_THREADS, 0, d_ts, d_off, chunk_base.p, chunk_unit.p, chunk_len, lane_units.p);
        }
        if (use_lanes && G) {
            // ---- one lane per chunk: the bulk of the chains; what they leave open is stitched below
            DBuf<uint32_t> lane_worklist;
            DBuf<uint2> lane_resume;
            DBuf<CompressCounters> lane_counters;
            TRY_SG(lane_worklist.alloc(G, s));
            TRY_SG(lane_resume.alloc(n_units, s));
            TRY_SG(lane_counters.alloc(1, s));
            TRY_SG(cudaMemsetAsync(lane_resume.p, 0, n_units * sizeof(uint2), s));
            const uint32_t *d_list = nullptr; // the first pass runs every chunk
            uint64_t n_list = G;
            // Pass 0, then rounds: k_spec_propagate gives every chunk whose chain did not start where the chain before it ends
            // that end as its new entry (optimistically: the chunk before may itself be re-run in the same round), and the lanes
            // re-run those chunks until they meet their old chains.  Chains of one unit are thus repaired in parallel, not one
            // after the other; a joint is still wrong afterwards only if the chunk before it changed its exit.  What is left
            // after a few rounds (and everything the lanes do not handle) is the stitching below, which alone decides what is final.
            for (uint32_t pass = 0; pass <= LANE_ROUNDS && n_list; pass++) {
                if (pass >= 1 && !ctx->lane_rounds_by_lanes) {
                    // a round's duration is its longest re-run (up to two chunk lengths): the cooperative engine walks a single
                    // chain ~15x faster than a lane does (measured: ~20 ns against ~300 ns per point when little else runs)
                    LAUNCH(ctx, k_spec_chain_warp, div_up(n_list, CHAIN_WARPS), CHAIN_WARPS * 32, 0, d_ts, d_val, d_off, d_kind, d_ebv, chunk_base.p,
                           chunk_unit.p, d_list, n_list, chunk_len, st.p, lists.p, list_base.p, list_cap.p);
                } else
                for (int kind = 0; kind < 3; kind++) {
                    if (!kind_units[kind]) continue;
                    const unsigned int n_blocks =
                        (unsigned int)std::min<uint64_t>((uint64_t)ctx->sm_count * lane_blocks_per_sm[kind], div_up(n_list, LANES_WARPS * 32));
                    TRY_SG(cudaMemsetAsync(lane_words.p + 3, 0, sizeof(unsigned int), s));
                    if (kind == KIND_LOSSLESS)
                        LAUNCH(ctx, k_spec_lanes<KIND_LOSSLESS>, n_blocks, LANES_WARPS * 32, 0, d_val, n_points, d_off, lane_units.p, chunk_base.p,
                               chunk_unit.p, chunk_len, ctx->lane_warmup, d_list, (uint32_t)n_list, st.p, lists.p, list_base.p, list_cap.p, lane_words.p + 3);
                    else if (kind == KIND_ABSOLUTE)
                        LAUNCH(ctx, k_spec_lanes<KIND_ABSOLUTE>, n_blocks, LANES_WARPS * 32, 0, d_val, n_points, d_off, lane_units.p, chunk_base.p,
                               chunk_unit.p, chunk_len, ctx->lane_warmup, d_list, (uint32_t)n_list, st.p, lists.p, list_base.p, list_cap.p, lane_words.p + 3);
                    else
                        LAUNCH(ctx, k_spec_lanes<KIND_RELATIVE>, n_blocks, LANES_WARPS * 32, 0, d_val, n_points, d_off, lane_units.p, chunk_base.p,
                               chunk_unit.p, chunk_len, ctx->lane_warmup, d_list, (uint32_t)n_list, st.p, lists.p, list_base.p, list_cap.p, lane_words.p + 3);
                }
                if (pass == LANE_ROUNDS) break;
                TRY_SG(cudaMemsetAsync(lane_counters.p, 0, sizeof(CompressCounters), s));
                LAUNCH(ctx, k_spec_propagate, div_up(n_units, 128), 128, 0, d_off, n_units, chunk_base.p, chunk_len, st.p, 1, lane_resume.p,
                       lane_worklist.p, lane_counters.p);
                TRY_SG(post(ctx, 0, lane_counters.p, 1));
                TRY_SG(sync_stream(ctx));
                CompressCounters hc;
                std::memcpy(&hc, ctx->mailbox, sizeof(hc));
                const uint64_t n_next = hc.dirty;
                if (pass >= 1 && n_next * 10 > n_list * 9) break; // the rounds no longer help (chunks the lanes do not handle)
                n_list = n_next;
                d_list = lane_worklist.p;
            }
            TRY_SG(cudaGetLastError());
        }
        for (int attempt = 0; async_sched && G && attempt < 2; attempt++) {
            if (attempt == 1) { // the overlapped check found an irregular unit among the screened ones: again, exactly
                use_screen = false;
                LAUNCH(ctx, k_spec_init, div_up(n_units, 128), 128, 0, d_off, n_units, chunk_base.p, chunk_len, st.p, chunk_unit.p, list_cap.p);
            }
            // ---- one persistent kernel: work queue of chunks, per-unit frontiers (sched_advance)
            int blocks_per_sm = 0;
            const bool wide_fit = !use_screen && (use_lanes || ctx->fit_wide);
            TRY_SG(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &blocks_per_sm,
                use_screen ? k_spec_async<WarpFitScreen, MDB_SCREEN_MIN_BLOCKS>
                           : (wide_fit ? k_spec_async<WarpFitWide, MDB_CHAIN_MIN_BLOCKS> : k_spec_async<WarpFit, MDB_CHAIN_MIN_BLOCKS>),
                CHAIN_WARPS * 32, 0));
            if (blocks_per_sm < 1) return bail(fail("compress: the chain kernel does not fit on this device"));
            const uint64_t n_blocks = std::min<uint64_t>((uint64_t)ctx->sm_count * blocks_per_sm, div_up(G, CHAIN_WARPS));
            const uint64_t capacity = 3 * G + n_blocks * CHAIN_WARPS + 8;
            if (capacity > 0xFFFFFFF0ull) return bail(fail("compress: too many chunks"));
            DBuf<UnitSched> units;
            DBuf<SchedQueue> queue;
            DBuf<uint32_t> items, per_index, cursor;
            DBuf<uint64_t> index_base;
            TRY_SG(units.alloc(n_units, s));
            TRY_SG(queue.alloc(1, s));
            TRY_SG(items.alloc(capacity, s));
            TRY_SG(per_index.alloc(G, s));
            TRY_SG(cursor.alloc(G, s));
            TRY_SG(index_base.alloc(G + 1, s));
            TRY_SG(cudaMemsetAsync(queue.p, 0, sizeof(SchedQueue), s));
            TRY_SG(cudaMemsetAsync(items.p, 0, capacity * sizeof(uint32_t), s));
            TRY_SG(cudaMemsetAsync(per_index.p, 0, G * sizeof(uint32_t), s));
            TRY_SG(cudaMemsetAsync(cursor.p, 0, G * sizeof(uint32_t), s));
            LAUNCH(ctx, k_sched_count, div_up(G, 256), 256, 0, chunk_base.p, chunk_unit.p, G, st.p, per_index.p);
            if (exclusive_scan<uint32_t>(ctx, per_index.p, G, index_base.p)) return bail(MDBCU_FAILURE);
            LAUNCH(ctx, k_sched_fill, div_up(G, 256), 256, 0, chunk_base.p, chunk_unit.p, G, st.p, index_base.p, cursor.p, items.p);
            // slots handed out without the queue (after the lanes only the chunks they left alone are pre-filled: their
            // number is known on the device alone, so every worker takes a ticket)
            const uint32_t n_initial = use_lanes ? 0u : (uint32_t)std::min<uint64_t>(n_blocks * CHAIN_WARPS, G);
            LAUNCH(ctx, k_sched_units, div_up(n_units, 128), 128, 0, chunk_base.p, n_units, G, (uint32_t)capacity, n_initial, units.p, queue.p);
            if (use_lanes) LAUNCH(ctx, k_sched_kick, div_up(n_units, 128), 128, 0, d_off, n_units, chunk_base.p, chunk_len, st.p, units.p, queue.p, items.p);
            const LaneUnit *no_info = nullptr;
            if (use_screen) // Swing's decisions screened in f32, the doubtful ones in the reference's f64 (mdb_fit_screen.cuh)
                LAUNCH(ctx, (k_spec_async<WarpFitScreen, MDB_SCREEN_MIN_BLOCKS>), (unsigned int)n_blocks, CHAIN_WARPS * 32, 0, d_ts, d_val, d_off, d_kind,
                       d_ebv, chunk_base.p, chunk_unit.p, chunk_len, st.p, lists.p, list_base.p, list_cap.p, units.p, queue.p, items.p, n_initial,
                       (uint32_t)G, lane_units.p);
            else if (wide_fit) // what the lanes left: stitching, and the long models they cut -- the engine with the wide steps
                LAUNCH(ctx, (k_spec_async<WarpFitWide, MDB_CHAIN_MIN_BLOCKS>), (unsigned int)n_blocks, CHAIN_WARPS * 32, 0, d_ts, d_val, d_off, d_kind,
                       d_ebv, chunk_base.p, chunk_unit.p, chunk_len, st.p, lists.p, list_base.p, list_cap.p, units.p, queue.p, items.p, n_initial,
                       (uint32_t)G, no_info);
            else
                LAUNCH(ctx, (k_spec_async<WarpFit, MDB_CHAIN_MIN_BLOCKS>), (unsigned int)n_blocks, CHAIN_WARPS * 32, 0, d_ts, d_val, d_off, d_kind,
                       d_ebv, chunk_base.p, chunk_unit.p, chunk_len, st.p, lists.p, list_base.p, list_cap.p, units.p, queue.p, items.p, n_initial,
                       (uint32_t)G, no_info);
            static_assert(sizeof(SchedQueue) == 32, "SchedQueue is posted as four words");
            TRY_SG(post(ctx, 0, queue.p, 4));
            const bool check_late = overlap_check && attempt == 0;
            if (check_late) { // the regularity check has finished by now, or is waited for here
                TRY_SG(cudaStreamWaitEvent(s, ctx->aux_done, 0));
                TRY_SG(cudaMemsetAsync(lane_words.p + 3, 0, sizeof(unsigned int), s));
                LAUNCH(ctx, k_lanes_late, div_up(n_units, 256), 256, 0, lane_units.p, n_units, lane_words.p + 3);
                TRY_SG(post(ctx, 4, lane_words.p + 2, 1)); // (words 2 and 3 of lane_words as one 64-bit word: the count is the high half)
            }
            TRY_SG(sync_stream(ctx));
            TRY_SG(cudaGetLastError());
            SchedQueue hq;
            std::memcpy(&hq, ctx->mailbox, sizeof(hq));
            if (hq.finished != 1 || hq.units_done != hq.live_units)
                return bail(fail("compress: chain scheduler stopped early (internal error)"));
            ctx->last_rounds = 1;
            if (!check_late || (ctx->mailbox[4] >> 32) == 0) break;
        }

        // ---- rounds: the chains of the chunks in the worklist, then the per-unit walk that builds the next worklist
        DBuf<uint32_t> worklist;
        DBuf<uint2> unit_resume;
        TRY_SG(worklist.alloc(G, s));
        TRY_SG(unit_resume.alloc(n_units, s));
        TRY_SG(cudaMemsetAsync(unit_resume.p, 0, n_units * sizeof(uint2), s));
        uint32_t round = 0;
        uint64_t n_work = async_sched ? 0 : G;
        const uint32_t *d_work = nullptr; // round 0 runs every chunk
        const bool trace = std::getenv("MDBCU_TRACE_ROUNDS") != nullptr; // diagnostics: chunks and wall time per round
        while (n_work) {
            const auto t_round = std::chrono::steady_clock::now();
            const uint64_t n_this = n_work;
            const uint32_t lanes = (uint32_t)std::min<uint64_t>(32, std::max<uint64_t>(1, div_up(n_work, (uint64_t)ctx->sm_count * 32)));
            if (ctx->fit_mode == 1)
                LAUNCH(ctx, k_spec_chain, div_up(n_work, lanes), 32, 0, d_ts, d_val, d_off, d_kind, d_ebv, chunk_base.p, chunk_unit.p, d_work,
                       n_work, lanes, chunk_len, st.p, lists.p, list_base.p, list_cap.p);
            else
                LAUNCH(ctx, k_spec_chain_warp, div_up(n_work, CHAIN_WARPS), CHAIN_WARPS * 32, 0, d_ts, d_val, d_off, d_kind, d_ebv, chunk_base.p,
                       chunk_unit.p, d_work, n_work, chunk_len, st.p, lists.p, list_base.p, list_cap.p);
            round++;
            TRY_SG(cudaMemsetAsync(counters.p, 0, sizeof(CompressCounters), s));
            LAUNCH(ctx, k_spec_propagate, div_up(n_units, 128), 128, 0, d_off, n_units, chunk_base.p, chunk_len, st.p, round == 1 ? 1 : 0,
                   unit_resume.p, worklist.p, counters.p);
            static_assert(sizeof(CompressCounters) == 8, "CompressCounters is posted as one word");
            TRY_SG(post(ctx, 0, counters.p, 1));
            TRY_SG(sync_stream(ctx));
            TRY_SG(cudaGetLastError());
            CompressCounters hc;
            std::memcpy(&hc, ctx->mailbox, sizeof(hc));
            n_work = hc.dirty;
            d_work = worklist.p;
            if (trace)
                fprintf(stderr, "[mdbcu] round %u: %llu of %llu chunks (chunk_len %u), %.3f ms\n", round, (unsigned long long)n_this,
                        (unsigned long long)G, chunk_len,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_round).count());
            if (round > 4 * G + 8) return bail(fail("compress: chunk fixpoint did not converge (internal error)"));
        }
        if (!async_sched) ctx->last_rounds = round;

        // ---- rows
        LAUNCH(ctx, k_spec_finalize, div_up(n_units, 128), 128, 0, d_off, n_units, chunk_base.p, chunk_len, st.p, unit_irregular.p);
        if (G && ctx->fit_mode != 1) {
            // pending Swing models: short ones one lane each; long ones are listed and get a block each (their number is only
            // known on the device: as many blocks as there could be, all but the listed ones leave at once)
            DBuf<uint2> long_models;
            DBuf<unsigned int> n_long;
            const uint64_t long_cap = (n_points - first) / SWING_FINISH_LONG + 1;
            TRY_SG(long_models.alloc(long_cap, s));
            TRY_SG(n_long.alloc(1, s));
            TRY_SG(cudaMemsetAsync(n_long.p, 0, sizeof(unsigned int), s));
            LAUNCH(ctx, k_swing_finish, div_up(G, 4), 128, 0, d_ts, d_val, d_off, chunk_unit.p, G, st.p, lists.p, list_base.p, list_cap.p, unit_irregular.p,
                   long_models.p, n_long.p);
            LAUNCH(ctx, k_swing_finish_long_par, (unsigned int)std::min<uint64_t>(long_cap, (uint64_t)ctx->sm_count * 8), 256, 0, d_ts, d_val, d_off,
                   chunk_unit.p, st.p, lists.p, list_base.p, list_cap.p, unit_irregular.p, long_models.p, n_long.p);
            LAUNCH(ctx, k_swing_finish_long, (unsigned int)std::min<uint64_t>(long_cap, 1u << 20), 64, 0, d_ts, d_val, d_off, chunk_unit.p, st.p, lists.p,
                   list_base.p, list_cap.p, long_models.p, n_long.p);
        }
        if (G) LAUNCH(ctx, k_spec_count_rows, div_up(G, 128), 128, 0, st.p, G, lists.p, list_base.p, list_cap.p, rows.p);
        if (exclusive_scan<uint32_t>(ctx, rows.p, G, row_base.p)) return bail(MDBCU_FAILURE);
        LAUNCH(ctx, k_unit_seg_off, div_up(n_units + 1, 256), 256, 0, chunk_base.p, n_units, row_base.p, sg->unit_seg_off);
        TRY_SG(post(ctx, 0, row_base.p + G, 1));
        TRY_SG(sync_stream(ctx));
        TRY_SG(cudaGetLastError());
        S = ctx->mailbox[0];
        TRY_SG(recs.alloc(S, s));
        TRY_SG(row_unit.alloc(S, s));
        const uint32_t lanes = (uint32_t)std::min<uint64_t>(32, std::max<uint64_t>(1, div_up(G, (uint64_t)ctx->sm_count * 32)));
        TRY_SG(wide_rows.alloc(S, s));
        TRY_SG(cudaMemsetAsync(n_wide.p, 0, 2 * sizeof(unsigned int), s));
        if (G && S) {
            LAUNCH(ctx, k_spec_records, div_up(G, 4), 128, 0, d_ts, d_val, d_off, d_kind, d_ebv, chunk_unit.p, G, st.p, lists.p,
                   list_base.p, list_cap.p, unit_irregular.p, row_base.p, recs.p, row_unit.p, wide_rows.p, n_wide.p);
            TRY_SG(post(ctx, 0, n_wide.p, 1));
            TRY_SG(sync_stream(ctx));
            n_wide_host = (uint32_t)ctx->mailbox[0];
            if (n_wide_host) { // long MacaqueV rows: encoded once, into a staging area (see k_records_macaque_warp)
                TRY_SG(wide_tmp.alloc(6 * (size_t)n_points + 64, s));
                LAUNCH(ctx, k_records_macaque_warp, std::min<unsigned int>(div_up(n_wide_host, WIDE_WARPS), (unsigned int)ctx->sm_count * 8), WIDE_WARPS * 32,
                       0, d_val, d_off, d_kind, d_ebv, row_unit.p, wide_rows.p, n_wide.p, recs.p, wide_tmp.p);
            }
        }
        // the scratch above is released (stream-ordered) when this scope ends
    } else {
        TRY_SG(cudaMemsetAsync(sg->unit_seg_off, 0, sizeof(uint64_t), s));
    }
    sg->n_segments = S;

    // ---- row metadata in final order + byte offsets of the three binary columns
    TRY_SG(cudaMallocAsync((void **)&sg->model_type_id, (S ? S : 1) * sizeof(int8_t), s));
    TRY_SG(cudaMallocAsync((void **)&sg->start_time, (S ? S : 1) * sizeof(int64_t), s));
    TRY_SG(cudaMallocAsync((void **)&sg->end_time, (S ? S : 1) * sizeof(int64_t), s));
    TRY_SG(cudaMallocAsync((void **)&sg->min_value, (S ? S : 1) * sizeof(float), s));
    TRY_SG(cudaMallocAsync((void **)&sg->max_value, (S ? S : 1) * sizeof(float), s));
    TRY_SG(cudaMallocAsync((void **)&sg->ts_off, (S + 1) * sizeof(uint64_t), s));
    TRY_SG(cudaMallocAsync((void **)&sg->val_off, (S + 1) * sizeof(uint64_t), s));
    TRY_SG(cudaMallocAsync((void **)&sg->res_off, (S + 1) * sizeof(uint64_t), s));
    DBuf<uint32_t> ts_len, val_len, res_len;
    TRY_SG(ts_len.alloc(S, s));
    TRY_SG(val_len.alloc(S, s));
    TRY_SG(res_len.alloc(S, s));
    if (S)
        LAUNCH(ctx, k_compress_gather, div_up(S, 256), 256, 0, d_ts, d_off, recs.p, row_unit.p, S, sg->model_type_id, sg->start_time, sg->end_time,
               sg->min_value, sg->max_value, ts_len.p, val_len.p, res_len.p);
    if (exclusive_scan<uint32_t>(ctx, ts_len.p, S, sg->ts_off)) return bail(MDBCU_FAILURE);
    if (exclusive_scan<uint32_t>(ctx, val_len.p, S, sg->val_off)) return bail(MDBCU_FAILURE);
    if (exclusive_scan<uint32_t>(ctx, res_len.p, S, sg->res_off)) return bail(MDBCU_FAILURE);
    TRY_SG(post(ctx, 0, sg->ts_off + S, 1));
    TRY_SG(post(ctx, 1, sg->val_off + S, 1));
    TRY_SG(post(ctx, 2, sg->res_off + S, 1));
    TRY_SG(sync_stream(ctx));
    sg->ts_bytes = ctx->mailbox[0];
    sg->val_bytes = ctx->mailbox[1];
    sg->res_bytes = ctx->mailbox[2];

    // ---- byte columns
    TRY_SG(cudaMallocAsync((void **)&sg->ts_data, sg->ts_bytes ? sg->ts_bytes : 1, s));
    TRY_SG(cudaMallocAsync((void **)&sg->val_data, sg->val_bytes ? sg->val_bytes : 1, s));
    TRY_SG(cudaMallocAsync((void **)&sg->res_data, sg->res_bytes ? sg->res_bytes : 1, s));
    if (S) {
        LAUNCH(ctx, k_compress_emit, div_up(S, 64), 64, 0, d_ts, d_val, d_off, d_kind, d_ebv, recs.p, row_unit.p, S, sg->ts_off, sg->ts_data,
               sg->val_off, sg->val_data, sg->res_off, sg->res_data);
        if (n_wide_host)
            LAUNCH(ctx, k_copy_macaque_warp, std::min<unsigned int>(n_wide_host, (unsigned int)ctx->sm_count * 32), COPY_THREADS, 0,
                   d_off, row_unit.p, wide_rows.p, n_wide.p, recs.p, wide_tmp.p, sg->val_off, sg->val_data);
    }
    TRY_SG(cudaGetLastError());
    TRY_SG(sync_stream(ctx));
#undef TRY_SG
    *out = sg;
    return MDBCU_SUCCESS;
}

} // extern "C"

// ---- diagnostics -----------------------------------------------------------------------------------
// fit_next_model for a list of start indices with either engine, so that tests can compare the
// warp-cooperative fit against the one-thread fit model by model.
struct DebugFit {
    uint32_t start_index, end_index;
    float min_value, max_value, model_last_value, bytes_per_value;
    int32_t model_type_id, values_len, aborted, irregular;
};

__global__ void __launch_bounds__(32) k_debug_fit(const int64_t *ts, const float *values, uint32_t n, int kind, float value, int engine,
                                                  const uint32_t *starts, const uint32_t *budget_ends, uint32_t n_starts, DebugFit *out) {
    __shared__ double smem[WarpFit::SMEM_DOUBLES];
    uint32_t k = blockIdx.x;
    if (k >= n_starts) return;
    ErrorBound eb = make_error_bound(kind, value);
    bool aborted = false, irregular = false;
    FittedModel m;
    if (engine == 5) { // the screened fit; the pre-pass over the unit (k_lanes_units, k_lanes_regular) is done here by the warp
        __shared__ LaneUnit lu;
        if (threadIdx.x == 0) lu = lane_unit_init(ts, n, eb);
        __syncwarp();
        bool bad = false;
        if (lu.ok)
            for (uint32_t i = threadIdx.x; i < n; i += 32) bad |= ts[i] != ts[0] + (int64_t)i * (ts[1] - ts[0]);
        bad = __any_sync(FULL_MASK, bad);
        if (threadIdx.x == 0 && bad) lu.irregular = 1;
        __syncwarp();
        WarpFitScreen f(eb, ts, values, n, smem, &lu);
        f.begin(starts[k]);
        m = f.fit(starts[k], budget_ends[k], aborted);
        irregular = f.irregular();
        if (!aborted && m.pending) swing_finish(m, ts, values);
    } else if (engine == 2) {
        WarpFit f(eb, ts, values, n, smem);
        f.begin(starts[k]);
        m = f.fit(starts[k], budget_ends[k], aborted);
        irregular = f.irregular();
        if (!aborted && m.pending) swing_finish(m, ts, values); // every lane, redundantly
    } else {
        ScalarFit f(eb, ts, values, n);
        f.begin(starts[k]);
        m = f.fit(starts[k], budget_ends[k], aborted);
        irregular = f.irregular();
    }
    if (threadIdx.x == 0) {
        DebugFit d;
        d.start_index = m.start_index; d.end_index = m.end_index;
        d.min_value = m.min_value; d.max_value = m.max_value; d.model_last_value = m.model_last_value;
        d.bytes_per_value = m.bytes_per_value;
        d.model_type_id = m.model_type_id; d.values_len = m.values_len; d.aborted = aborted; d.irregular = irregular;
        out[k] = d;
    }
}

extern "C" int mdbcu_debug_fit_models(mdbcu_context *ctx, const int64_t *timestamps, const float *values, uint32_t n, int eb_kind,
                                      float eb_value, int engine, const uint32_t *starts, const uint32_t *budget_ends, uint32_t n_starts,
                                      void *out /* n_starts x 40 bytes */) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    cudaStream_t s = ctx->stream;
    DBuf<int64_t> d_ts;
    DBuf<float> d_val;
    DBuf<uint32_t> d_starts, d_budget;
    DBuf<DebugFit> d_out;
    CUDA_TRY(upload(ctx, d_ts, timestamps, n));
    CUDA_TRY(upload(ctx, d_val, values, n));
    CUDA_TRY(upload(ctx, d_starts, starts, n_starts));
    CUDA_TRY(upload(ctx, d_budget, budget_ends, n_starts));
    CUDA_TRY(d_out.alloc(n_starts, s));
    if (n_starts) LAUNCH(ctx, k_debug_fit, n_starts, 32, 0, d_ts.p, d_val.p, n, eb_kind, eb_value, engine, d_starts.p, d_budget.p, n_starts, d_out.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(d2h_bytes(ctx, out, d_out.p, n_starts * sizeof(DebugFit)));
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}

// Diagnostics: rewrite_position (mdb_device.cuh) as a step function of the f32 bit pattern, computed ON THE DEVICE for every
// pattern in [first_bits, last_bits]: the patterns at which the position differs from the one before (first_bits is always
// listed).  The tests compare the list with the oracle's libm form (oracle/mdb_oracle.cc: mdbo_rewrite_position_steps).
__global__ void __launch_bounds__(256) k_debug_rewrite_steps(uint32_t first_bits, uint64_t count, uint32_t *bits_out, int32_t *pos_out, uint32_t cap,
                                                             unsigned int *n_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const uint32_t b = first_bits + (uint32_t)k;
        const int32_t p = rewrite_position(__uint_as_float(b));
        if (k == 0 || rewrite_position(__uint_as_float(b - 1u)) != p) {
            const unsigned int slot = atomicAdd(n_out, 1u);
            if (slot < cap) {
                bits_out[slot] = b;
                pos_out[slot] = p;
            }
        }
    }
}

extern "C" int mdbcu_debug_rewrite_position_steps(mdbcu_context *ctx, uint32_t first_bits, uint32_t last_bits, uint32_t *bits_out, int32_t *pos_out,
                                                  uint32_t cap, uint32_t *n_steps) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (last_bits < first_bits || !bits_out || !pos_out || !n_steps) return fail("rewrite_position_steps: bad arguments");
    cudaStream_t s = ctx->stream;
    DBuf<uint32_t> d_bits;
    DBuf<int32_t> d_pos;
    DBuf<unsigned int> d_n;
    CUDA_TRY(d_bits.alloc(cap, s));
    CUDA_TRY(d_pos.alloc(cap, s));
    CUDA_TRY(d_n.alloc(1, s));
    CUDA_TRY(cudaMemsetAsync(d_n.p, 0, sizeof(unsigned int), s));
    const uint64_t count = (uint64_t)last_bits - first_bits + 1;
    LAUNCH(ctx, k_debug_rewrite_steps, (unsigned int)ctx->sm_count * 16, 256, 0, first_bits, count, d_bits.p, d_pos.p, cap, d_n.p);
    CUDA_TRY(cudaGetLastError());
    unsigned int n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, d_n.p, sizeof(n), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(sync_stream(ctx));
    *n_steps = n;
    const uint32_t m = n < cap ? n : cap;
    CUDA_TRY(d2h_bytes(ctx, bits_out, d_bits.p, m * sizeof(uint32_t)));
    CUDA_TRY(d2h_bytes(ctx, pos_out, d_pos.p, m * sizeof(int32_t)));
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS; // (the entries are in completion order: the caller sorts them by bit pattern)
}

// Diagnostics: reads (and clears) the fit event counters; all zero unless built with -DMDB_FIT_COUNTERS.
extern "C" int mdbcu_debug_counters(mdbcu_context *ctx, uint64_t *out8 /* 16 entries */) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[16], z[16] = {0};
    CUDA_TRY(cudaMemcpyFromSymbol(h, g_fit_counters, sizeof(h)));
    CUDA_TRY(cudaMemcpyToSymbol(g_fit_counters, z, sizeof(z)));
    for (int i = 0; i < 16; i++) out8[i] = h[i];
    return MDBCU_SUCCESS;
}
