// mdb_sort.inl -- the front end of try_compress_multivariate_time_series on the device (included at the end of mdb_cuda.cu).
//
// The reference sorts the rows of an ingested batch by all tag columns and then by time (`lexsort_to_indices` + `take`,
// compression.rs:110-141) and cuts the sorted rows into time series where a tag changes (:64-93).  The tag columns are strings
// and stay with the host, which hands every row the dense CODE of its tag tuple (codes ordered like the tuples); the sort
// itself -- rows by (code, timestamp) -- and the gather of the timestamp and field columns run here:
//   k_sort_range     min / max timestamp and the largest code: only the bytes that differ are sorted on
//   k_sort_hist / k_sort_scatter   one stable LSD radix pass over 8 bits of the key: per-block digit histograms, one exclusive
//                    scan over (digit, block), and a scatter that ranks equal digits in row order (per round of 256 rows:
//                    __match_any_sync inside a warp, a prefix over the warps' counts, a running count per digit)
//   k_take_*         out[i] = in[order[i]]
// Passes: the significant bytes of (timestamp - min timestamp), least significant first, then those of the code -- a stable
// sort by the minor key followed by a stable sort by the major key is the lexicographic sort, and rows with equal (tags,
// timestamp) keep their input order (the reference leaves their order unspecified).

constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS = 8;                         // rows per thread
constexpr int SORT_BLOCK = SORT_THREADS * SORT_ROUNDS; // rows per block

struct SortRange {
    unsigned long long ts_min_biased, ts_max_biased; // timestamps with the sign bit flipped (unsigned order = signed order)
    unsigned int code_max, pad;
};

__device__ __forceinline__ unsigned long long sort_bias(int64_t t) { return (unsigned long long)t ^ 0x8000000000000000ull; }

__global__ void __launch_bounds__(256) k_sort_range(const uint32_t *__restrict__ code, const int64_t *__restrict__ ts, uint64_t n, SortRange *range) {
    unsigned long long lo = ~0ull, hi = 0ull;
    unsigned int cmax = 0u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long b = sort_bias(ts[i]);
        lo = min(lo, b);
        hi = max(hi, b);
        cmax = max(cmax, code[i]);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&range->ts_min_biased, lo);
        atomicMax(&range->ts_max_biased, hi);
        atomicMax(&range->code_max, cmax);
    }
}

// Digit `pass` of row r's key: passes 0 .. ts_passes - 1 are bytes of (timestamp - min), the following ones bytes of the code.
__device__ __forceinline__ unsigned int sort_digit(const uint32_t *__restrict__ code, const int64_t *__restrict__ ts, unsigned long long ts_min_biased,
                                                   int pass, int ts_passes, uint32_t r) {
    if (pass < ts_passes) return (unsigned int)(((sort_bias(ts[r]) - ts_min_biased) >> (8 * pass)) & 255ull);
    return (code[r] >> (8 * (pass - ts_passes))) & 255u;
}

// hist[digit * n_blocks + block] = rows of the block with that digit (rows taken in the order of order_in; nullptr: identity)
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint32_t *__restrict__ code, const int64_t *__restrict__ ts, const SortRange *range, int pass,
                                                            int ts_passes, const uint32_t *__restrict__ order_in, uint64_t n, uint32_t *hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long ts_min = range->ts_min_biased;
    const uint64_t base = (uint64_t)blockIdx.x * SORT_BLOCK;
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint64_t i = base + (uint64_t)r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[sort_digit(code, ts, ts_min, pass, ts_passes, order_in ? order_in[i] : (uint32_t)i)], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// offset[digit * n_blocks + block]: where the block's first row with that digit goes (the exclusive scan of hist)
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint32_t *__restrict__ code, const int64_t *__restrict__ ts, const SortRange *range,
                                                               int pass, int ts_passes, const uint32_t *__restrict__ order_in, uint64_t n,
                                                               const uint64_t *__restrict__ offset, uint32_t *__restrict__ order_out) {
    __shared__ unsigned int cnt[SORT_THREADS / 32][256]; // this round's rows per (warp, digit)
    __shared__ unsigned int running[256];                // rows of earlier rounds per digit
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    running[threadIdx.x] = 0;
    for (int w = 0; w < SORT_THREADS / 32; w++) cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long ts_min = range->ts_min_biased;
    const uint64_t base = (uint64_t)blockIdx.x * SORT_BLOCK;
    const uint64_t my_offset = offset[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x]; // thread d keeps digit d's base
    __shared__ uint64_t digit_base[256];
    digit_base[threadIdx.x] = my_offset;
    __syncthreads();
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint64_t i = base + (uint64_t)r * SORT_THREADS + threadIdx.x; // rows of a round are consecutive: (round, thread) is row order
        const bool valid = i < n;
        uint32_t row = 0;
        unsigned int d = 256u + (unsigned int)lane; // (invalid rows: a digit nobody shares)
        if (valid) {
            row = order_in ? order_in[i] : (uint32_t)i;
            d = sort_digit(code, ts, ts_min, pass, ts_passes, row);
        }
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const unsigned int before = __popc(peers & ((1u << lane) - 1u)); // equal digits in earlier lanes of this warp
        if (valid && before == 0) cnt[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned int earlier_warps = 0;
            for (int w = 0; w < warp; w++) earlier_warps += cnt[w][d];
            order_out[digit_base[d] + running[d] + earlier_warps + before] = row;
        }
        __syncthreads();
        {
            unsigned int total = 0;
            for (int w = 0; w < SORT_THREADS / 32; w++) {
                total += cnt[w][threadIdx.x];
                cnt[w][threadIdx.x] = 0;
            }
            running[threadIdx.x] += total;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_take_i64(const uint32_t *__restrict__ order, uint64_t n, const int64_t *__restrict__ in, int64_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[order[i]];
}
__global__ void __launch_bounds__(256) k_take_f32(const uint32_t *__restrict__ order, uint64_t n, const float *__restrict__ in, float *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[order[i]];
}
__global__ void __launch_bounds__(256) k_take_u32(const uint32_t *__restrict__ order, uint64_t n, const uint32_t *__restrict__ in, uint32_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[order[i]];
}

extern "C" {

int mdbcu_sort_rows(mdbcu_context *ctx, mdbcu_space space, const uint32_t *series_code, const int64_t *timestamps, uint64_t n, uint32_t *order_out) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (n == 0) return MDBCU_SUCCESS;
    if (n > 0xFFFFFFF0ull) return fail("sort_rows: more than 2^32 rows in one batch");
    if (!series_code || !timestamps || !order_out) return fail("sort_rows: null pointer");
    cudaStream_t s = ctx->stream;
    DBuf<uint32_t> code_buf, order_buf[2], hist;
    DBuf<int64_t> ts_buf;
    DBuf<uint64_t> offset;
    DBuf<SortRange> range;
    const uint32_t *d_code = series_code;
    const int64_t *d_ts = timestamps;
    if (space == MDBCU_HOST) {
        CUDA_TRY(upload(ctx, code_buf, series_code, n));
        CUDA_TRY(upload(ctx, ts_buf, timestamps, n));
        d_code = code_buf.p;
        d_ts = ts_buf.p;
    }
    CUDA_TRY(range.alloc(1, s));
    const SortRange init = {~0ull, 0ull, 0u, 0u};
    CUDA_TRY(cudaMemcpyAsync(range.p, &init, sizeof(init), cudaMemcpyHostToDevice, s));
    LAUNCH(ctx, k_sort_range, (unsigned int)std::min<uint64_t>(div_up(n, 256), (uint64_t)ctx->sm_count * 8), 256, 0, d_code, d_ts, n, range.p);
    static_assert(sizeof(SortRange) == 24, "SortRange is posted as three words");
    CUDA_TRY(post(ctx, 0, range.p, 3));
    CUDA_TRY(sync_stream(ctx));
    SortRange h;
    std::memcpy(&h, ctx->mailbox, sizeof(h));
    auto bytes_of = [](unsigned long long x) { int b = 0; while (x) { b++; x >>= 8; } return b; };
    const int ts_passes = bytes_of(h.ts_max_biased - h.ts_min_biased), code_passes = bytes_of(h.code_max);
    const int passes = ts_passes + code_passes;
    const unsigned int n_blocks = div_up(n, SORT_BLOCK);
    CUDA_TRY(order_buf[0].alloc(n, s));
    CUDA_TRY(order_buf[1].alloc(n, s));
    CUDA_TRY(hist.alloc((size_t)256 * n_blocks, s));
    CUDA_TRY(offset.alloc((size_t)256 * n_blocks + 1, s));
    // the last pass writes straight into the caller's array when that is device memory
    const uint32_t *in = nullptr; // identity
    for (int p = 0; p < passes; p++) {
        uint32_t *out = (p == passes - 1 && space == MDBCU_DEVICE) ? order_out : order_buf[p & 1].p;
        LAUNCH(ctx, k_sort_hist, n_blocks, SORT_THREADS, 0, d_code, d_ts, range.p, p, ts_passes, in, n, hist.p);
        if (exclusive_scan<uint32_t>(ctx, hist.p, (uint64_t)256 * n_blocks, offset.p)) return MDBCU_FAILURE;
        LAUNCH(ctx, k_sort_scatter, n_blocks, SORT_THREADS, 0, d_code, d_ts, range.p, p, ts_passes, in, n, offset.p, out);
        in = out;
    }
    CUDA_TRY(cudaGetLastError());
    if (passes == 0) { // every key equal: the identity
        std::vector<uint32_t> ident(n);
        for (uint64_t i = 0; i < n; i++) ident[i] = (uint32_t)i;
        if (space == MDBCU_HOST) std::memcpy(order_out, ident.data(), n * sizeof(uint32_t));
        else {
            CUDA_TRY(cudaMemcpyAsync(order_out, ident.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            CUDA_TRY(sync_stream(ctx)); // (`ident` goes out of scope)
        }
    } else if (space == MDBCU_HOST) {
        CUDA_TRY(d2h_bytes(ctx, order_out, in, n * sizeof(uint32_t)));
    }
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}

int mdbcu_take_rows(mdbcu_context *ctx, mdbcu_space space, const uint32_t *order, uint64_t n, const int64_t *timestamps_in, int64_t *timestamps_out,
                    const float *const *fields_in, float *const *fields_out, uint32_t n_fields) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (n == 0) return MDBCU_SUCCESS;
    if (!order || (n_fields && (!fields_in || !fields_out))) return fail("take_rows: null pointer");
    cudaStream_t s = ctx->stream;
    DBuf<uint32_t> order_buf;
    const uint32_t *d_order = order;
    if (space == MDBCU_HOST) {
        CUDA_TRY(upload(ctx, order_buf, order, n));
        d_order = order_buf.p;
    }
    const unsigned int blocks = div_up(n, 256);
    if (timestamps_in) {
        if (!timestamps_out) return fail("take_rows: timestamps_out is null");
        if (space == MDBCU_HOST) {
            DBuf<int64_t> a, b;
            CUDA_TRY(upload(ctx, a, timestamps_in, n));
            CUDA_TRY(b.alloc(n, s));
            LAUNCH(ctx, k_take_i64, blocks, 256, 0, d_order, n, a.p, b.p);
            CUDA_TRY(d2h_bytes(ctx, timestamps_out, b.p, n * sizeof(int64_t)));
            CUDA_TRY(sync_stream(ctx));
        } else {
            LAUNCH(ctx, k_take_i64, blocks, 256, 0, d_order, n, timestamps_in, timestamps_out);
        }
    }
    for (uint32_t f = 0; f < n_fields; f++) {
        if (!fields_in[f] || !fields_out[f]) return fail("take_rows: field pointer is null");
        if (space == MDBCU_HOST) {
            DBuf<float> a, b;
            CUDA_TRY(upload(ctx, a, fields_in[f], n));
            CUDA_TRY(b.alloc(n, s));
            LAUNCH(ctx, k_take_f32, blocks, 256, 0, d_order, n, a.p, b.p);
            CUDA_TRY(d2h_bytes(ctx, fields_out[f], b.p, n * sizeof(float)));
            CUDA_TRY(sync_stream(ctx));
        } else {
            LAUNCH(ctx, k_take_f32, blocks, 256, 0, d_order, n, fields_in[f], fields_out[f]);
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}

} // extern "C"
