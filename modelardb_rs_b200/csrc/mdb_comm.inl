// mdb_comm.inl -- multi-GPU behind the C-ABI (included at the end of mdb_cuda.cu).
//
// The hot path shards by unit (time series) with no data-path exchange: a rank compresses, grids and aggregates the units
// it owns.  The one thing that travels is the result of an aggregate query: per-group (COUNT, MIN, MAX, SUM) records of
// 24 bytes.  For GROUP BY series the ranks own disjoint, contiguous ranges of groups (mdbcu_shard_units), so ONE
// ncclAllGather of the packed records yields the table's result in unit order on every rank; for an ungrouped aggregate
// every rank contributes one record and the records are folded in rank order (= row order), so the f64 SUM does not
// depend on a reduction tree (model_simple_aggregates.rs:481-511 adds row by row).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library has no link-time dependency on it, a host that never
// creates a communicator never loads it, and inside a process that already has NCCL loaded (PyTorch) the same copy is used.
// One communicator per context: one process (or host thread) per GPU creates its context and joins with
// mdbcu_comm_create; mdbcu_comm_create_all is the single-process form (ncclCommInitAll over the contexts' devices).

#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.error = std::string("NCCL is not available: ") + dlerror();
        return &api;
    }
    auto sym = [&](const char *name) {
        void *p = dlsym(api.handle, name);
        if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name;
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return &api;
}

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (expr);                                                                   \
        if (r_ != ncclSuccess) return fail(std::string(#expr) + ": " + nccl_api()->GetErrorString(r_)); \
    } while (0)

struct mdbcu_comm {
    mdbcu_context *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
};

struct AggRecord { // 24 bytes on the wire
    int64_t count;
    float min, max;
    double sum;
};
static_assert(sizeof(AggRecord) == 24, "AggRecord layout");

__global__ void __launch_bounds__(256) k_agg_pack(const int64_t *count, const float *mn, const float *mx, const double *sum, uint64_t n, uint64_t padded,
                                                  AggRecord *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    AggRecord r;
    r.count = i < n ? count[i] : 0;
    r.min = i < n ? mn[i] : 0.0f;
    r.max = i < n ? mx[i] : 0.0f;
    r.sum = i < n ? sum[i] : 0.0;
    out[i] = r;
}

// gathered: world x widest records; group g of the table is record (g - lo_r) of rank r's block.
__global__ void __launch_bounds__(256) k_agg_unpack(const AggRecord *gathered, uint64_t n_total, int world, uint64_t widest, int64_t *count, float *mn,
                                                    float *mx, double *sum) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_total) return;
    const uint64_t base = n_total / (uint64_t)world, extra = n_total % (uint64_t)world; // mdbcu_shard_units
    const uint64_t big = extra * (base + 1);
    const uint64_t r = g < big ? g / (base + 1) : extra + (g - big) / (base ? base : 1);
    const uint64_t lo = r * base + (r < extra ? r : extra);
    const AggRecord rec = gathered[r * widest + (g - lo)];
    count[g] = rec.count;
    mn[g] = rec.min;
    mx[g] = rec.max;
    sum[g] = rec.sum;
}

// One record per rank, folded in rank order with the accumulators' own operations.
__global__ void k_agg_fold_ranks(const AggRecord *gathered, int world, int64_t *count, float *mn, float *mx, double *sum) {
    if (threadIdx.x || blockIdx.x) return;
    GroupAgg a = group_agg_identity();
    for (int r = 0; r < world; r++) {
        GroupAgg b;
        b.count = gathered[r].count;
        b.min = gathered[r].min;
        b.max = gathered[r].max;
        b.sum = gathered[r].sum;
        a = group_agg_combine(a, b);
    }
    *count = a.count;
    *mn = a.min;
    *mx = a.max;
    *sum = a.sum;
}

extern "C" {

int mdbcu_shard_units(uint64_t n_units, int world, int rank, uint64_t *lo, uint64_t *hi) {
    if (world < 1 || rank < 0 || rank >= world || !lo || !hi) return fail("shard_units: rank / world out of range");
    const uint64_t base = n_units / (uint64_t)world, extra = n_units % (uint64_t)world, r = (uint64_t)rank;
    *lo = r * base + (r < extra ? r : extra);
    *hi = *lo + base + (r < extra ? 1 : 0);
    return MDBCU_SUCCESS;
}

int mdbcu_comm_unique_id(uint8_t *id128) {
    if (!id128) return fail("comm_unique_id: id is null");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(api->error);
    ncclUniqueId id;
    NCCL_TRY(api->GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, sizeof(id));
    return MDBCU_SUCCESS;
}

int mdbcu_comm_create(mdbcu_context *ctx, int world, int rank, const uint8_t *id128, mdbcu_comm **out) {
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (!out || !id128) return fail("comm_create: out / id is null");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail("comm_create: rank / world out of range");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(api->error);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    NCCL_TRY(api->CommInitRank(&c, world, id, rank));
    mdbcu_comm *m = new mdbcu_comm();
    m->ctx = ctx;
    m->comm = c;
    m->world = world;
    m->rank = rank;
    *out = m;
    return MDBCU_SUCCESS;
}

int mdbcu_comm_create_all(mdbcu_context *const *ctxs, int n, mdbcu_comm **out) {
    if (!ctxs || !out || n < 1) return fail("comm_create_all: bad arguments");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(api->error);
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return fail("comm_create_all: context is null");
        devs[(size_t)i] = ctxs[i]->device;
    }
    std::vector<ncclComm_t> comms((size_t)n, nullptr);
    NCCL_TRY(api->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) {
        mdbcu_comm *m = new mdbcu_comm();
        m->ctx = ctxs[i];
        m->comm = comms[(size_t)i];
        m->world = n;
        m->rank = i;
        out[i] = m;
    }
    return MDBCU_SUCCESS;
}

void mdbcu_comm_destroy(mdbcu_comm *comm) {
    if (!comm) return;
    if (comm->comm && nccl_api()->CommDestroy) {
        cudaSetDevice(comm->ctx->device);
        cudaStreamSynchronize(comm->ctx->stream);
        nccl_api()->CommDestroy(comm->comm);
    }
    delete comm;
}

int mdbcu_comm_world(const mdbcu_comm *comm) { return comm ? comm->world : 0; }
int mdbcu_comm_rank(const mdbcu_comm *comm) { return comm ? comm->rank : -1; }

// GROUP BY over a table whose groups (units) are sharded over the ranks in contiguous balanced ranges (mdbcu_shard_units):
// `segments` / `group_off` describe THIS rank's n_local groups; count / min / max / sum receive all n_total groups of the table
// in unit order, on every rank.  One packed ncclAllGather; everything else is mdbcu_aggregate.
int mdbcu_aggregate_sharded(mdbcu_comm *comm, mdbcu_space space, const mdbcu_segments_view *segments, const uint64_t *group_off, uint64_t n_local,
                            uint64_t n_total, int64_t *count, float *min, float *max, double *sum) {
    if (!comm) return fail("aggregate_sharded: communicator is null");
    mdbcu_context *ctx = comm->ctx;
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    uint64_t lo, hi;
    if (mdbcu_shard_units(n_total, comm->world, comm->rank, &lo, &hi)) return MDBCU_FAILURE;
    if (hi - lo != n_local) return fail("aggregate_sharded: this rank owns " + std::to_string(hi - lo) + " of the " + std::to_string(n_total) +
                                        " groups, " + std::to_string(n_local) + " were passed");
    if (n_total == 0) return MDBCU_SUCCESS;
    if (!count || !min || !max || !sum) return fail("aggregate_sharded: output pointer is null");
    if (n_local && !group_off) return fail("aggregate_sharded: group_off is null");
    cudaStream_t s = ctx->stream;
    const uint64_t widest = (n_total + (uint64_t)comm->world - 1) / (uint64_t)comm->world;
    DBuf<int64_t> l_count, g_count;
    DBuf<float> l_min, l_max, g_min, g_max;
    DBuf<double> l_sum, g_sum;
    DBuf<AggRecord> packed, gathered;
    CUDA_TRY(l_count.alloc(n_local, s));
    CUDA_TRY(l_min.alloc(n_local, s));
    CUDA_TRY(l_max.alloc(n_local, s));
    CUDA_TRY(l_sum.alloc(n_local, s));
    CUDA_TRY(packed.alloc(widest, s));
    CUDA_TRY(gathered.alloc(widest * (uint64_t)comm->world, s));
    if (n_local) {
        // local groups on the device (the device-space call does not wait for the stream when the outputs are device memory)
        if (space == MDBCU_DEVICE) {
            if (mdbcu_aggregate(ctx, space, segments, group_off, n_local, l_count.p, l_min.p, l_max.p, l_sum.p)) return MDBCU_FAILURE;
        } else {
            std::vector<int64_t> hc(n_local);
            std::vector<float> hmn(n_local), hmx(n_local);
            std::vector<double> hs(n_local);
            if (mdbcu_aggregate(ctx, space, segments, group_off, n_local, hc.data(), hmn.data(), hmx.data(), hs.data())) return MDBCU_FAILURE;
            CUDA_TRY(h2d_bytes(ctx, l_count.p, hc.data(), n_local * sizeof(int64_t)));
            CUDA_TRY(h2d_bytes(ctx, l_min.p, hmn.data(), n_local * sizeof(float)));
            CUDA_TRY(h2d_bytes(ctx, l_max.p, hmx.data(), n_local * sizeof(float)));
            CUDA_TRY(h2d_bytes(ctx, l_sum.p, hs.data(), n_local * sizeof(double)));
            CUDA_TRY(sync_stream(ctx)); // (the staging vectors go out of scope)
        }
    }
    LAUNCH(ctx, k_agg_pack, div_up(widest, 256), 256, 0, l_count.p, l_min.p, l_max.p, l_sum.p, n_local, widest, packed.p);
    NCCL_TRY(nccl_api()->AllGather(packed.p, gathered.p, widest * sizeof(AggRecord), ncclUint8, comm->comm, s));
    ctx->launches++; // (the collective's kernel)
    int64_t *d_count = count;
    float *d_min = min, *d_max = max;
    double *d_sum = sum;
    if (space == MDBCU_HOST) {
        CUDA_TRY(g_count.alloc(n_total, s));
        CUDA_TRY(g_min.alloc(n_total, s));
        CUDA_TRY(g_max.alloc(n_total, s));
        CUDA_TRY(g_sum.alloc(n_total, s));
        d_count = g_count.p; d_min = g_min.p; d_max = g_max.p; d_sum = g_sum.p;
    }
    LAUNCH(ctx, k_agg_unpack, div_up(n_total, 256), 256, 0, gathered.p, n_total, comm->world, widest, d_count, d_min, d_max, d_sum);
    CUDA_TRY(cudaGetLastError());
    if (space == MDBCU_HOST) {
        CUDA_TRY(d2h_bytes(ctx, count, d_count, n_total * sizeof(int64_t)));
        CUDA_TRY(d2h_bytes(ctx, min, d_min, n_total * sizeof(float)));
        CUDA_TRY(d2h_bytes(ctx, max, d_max, n_total * sizeof(float)));
        CUDA_TRY(d2h_bytes(ctx, sum, d_sum, n_total * sizeof(double)));
    }
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}

// The ungrouped aggregate (what the reference's rule rewrites) over rows sharded across the ranks: one record per rank,
// gathered and folded in rank order.  Outputs are single values in `space`, identical on every rank.
int mdbcu_aggregate_all_sharded(mdbcu_comm *comm, mdbcu_space space, const mdbcu_segments_view *segments, int64_t *count, float *min, float *max,
                                double *sum) {
    if (!comm) return fail("aggregate_all_sharded: communicator is null");
    mdbcu_context *ctx = comm->ctx;
    if (check_ctx(ctx)) return MDBCU_FAILURE;
    if (!count || !min || !max || !sum) return fail("aggregate_all_sharded: output pointer is null");
    cudaStream_t s = ctx->stream;
    DBuf<int64_t> l_count;
    DBuf<float> l_min, l_max;
    DBuf<double> l_sum;
    DBuf<AggRecord> packed, gathered;
    CUDA_TRY(l_count.alloc(1, s));
    CUDA_TRY(l_min.alloc(1, s));
    CUDA_TRY(l_max.alloc(1, s));
    CUDA_TRY(l_sum.alloc(1, s));
    CUDA_TRY(packed.alloc(1, s));
    CUDA_TRY(gathered.alloc((uint64_t)comm->world, s));
    if (space == MDBCU_DEVICE) {
        if (mdbcu_aggregate(ctx, space, segments, nullptr, 1, l_count.p, l_min.p, l_max.p, l_sum.p)) return MDBCU_FAILURE;
    } else {
        int64_t hc;
        float hmn, hmx;
        double hs;
        if (mdbcu_aggregate(ctx, space, segments, nullptr, 1, &hc, &hmn, &hmx, &hs)) return MDBCU_FAILURE;
        CUDA_TRY(h2d_bytes(ctx, l_count.p, &hc, sizeof(hc)));
        CUDA_TRY(h2d_bytes(ctx, l_min.p, &hmn, sizeof(hmn)));
        CUDA_TRY(h2d_bytes(ctx, l_max.p, &hmx, sizeof(hmx)));
        CUDA_TRY(h2d_bytes(ctx, l_sum.p, &hs, sizeof(hs)));
        CUDA_TRY(sync_stream(ctx));
    }
    LAUNCH(ctx, k_agg_pack, 1, 256, 0, l_count.p, l_min.p, l_max.p, l_sum.p, (uint64_t)1, (uint64_t)1, packed.p);
    NCCL_TRY(nccl_api()->AllGather(packed.p, gathered.p, sizeof(AggRecord), ncclUint8, comm->comm, s));
    ctx->launches++;
    LAUNCH(ctx, k_agg_fold_ranks, 1, 32, 0, gathered.p, comm->world, l_count.p, l_min.p, l_max.p, l_sum.p);
    CUDA_TRY(cudaGetLastError());
    if (space == MDBCU_HOST) {
        CUDA_TRY(d2h_bytes(ctx, count, l_count.p, sizeof(int64_t)));
        CUDA_TRY(d2h_bytes(ctx, min, l_min.p, sizeof(float)));
        CUDA_TRY(d2h_bytes(ctx, max, l_max.p, sizeof(float)));
        CUDA_TRY(d2h_bytes(ctx, sum, l_sum.p, sizeof(double)));
    } else {
        CUDA_TRY(cudaMemcpyAsync(count, l_count.p, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(min, l_min.p, sizeof(float), cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(max, l_max.p, sizeof(float), cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(sum, l_sum.p, sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    CUDA_TRY(sync_stream(ctx));
    return MDBCU_SUCCESS;
}

} // extern "C"
