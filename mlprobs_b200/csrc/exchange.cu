// Multi-GPU exchange of the sparse posteriors: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Every rank computes the pairs of its shard (mlp_set_shard).  Because the row-pointer pool and the per-pair tables
// have the SAME fixed layout on every rank and slots of foreign pairs are kept zero, "gather" is a sum:
//   distances, row pointers, cell counts, cell offsets -> ncclAllReduce(sum) in place (x + 0 == x, exact);
//   cells (variable size, bump-allocated per rank)       -> one grouped ncclBroadcast per rank into a common pool, rank
//                                                          q's cells landing at base_q = sum of the lower ranks' usage
//                                                          (the cell offsets of owned pairs are shifted by base_q first).
// NCCL is bound at run time (dlopen libnccl.so.2) so the library loads on machines without it; inside a torch process
// the already-loaded torch-bundled NCCL is the one that resolves.
#include "ctx.h"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.h = h;
#define BIND(field, sym) *(void**)(&g_nccl.field) = dlsym(h, sym); if (!g_nccl.field) return false;
    BIND(GetUniqueId, "ncclGetUniqueId") BIND(CommInitRank, "ncclCommInitRank") BIND(CommDestroy, "ncclCommDestroy")
    BIND(AllReduce, "ncclAllReduce") BIND(AllGather, "ncclAllGather") BIND(Broadcast, "ncclBroadcast")
    BIND(GroupStart, "ncclGroupStart") BIND(GroupEnd, "ncclGroupEnd") BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    g_nccl.ok = true;
    return true;
}

__global__ void k_shift_offsets(const PairTask* tasks, int ntasks, int n, long long* nz_off, long long base) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntasks) return;
    const PairTask p = tasks[t];
    nz_off[(long long)p.a * n + p.b] += base;
    nz_off[(long long)p.b * n + p.a] += base;
}
}  // namespace

#define NK(call)                                                                              \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + g_nccl.GetErrorString(r__);                \
            return MLP_E_NCCL;                                                                \
        }                                                                                     \
    } while (0)

extern "C" int mlp_nccl_unique_id(uint8_t id128[128]) {
    if (!id128) return MLP_E_ARG;
    if (!load_nccl()) return MLP_E_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MLP_E_NCCL;
    memcpy(id128, &id, 128);
    return MLP_OK;
}

extern "C" int mlp_comm_init(mlp_ctx* ctx, const uint8_t id128[128], int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return MLP_E_ARG;
    if (!load_nccl()) { ctx->err = "libnccl.so.2 could not be loaded"; return MLP_E_NCCL; }
    cudaSetDevice(ctx->device);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NK(g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
    return mlp_set_shard(ctx, rank, world);
}

// QuickProbs cells are (column < 65536, code / 65535.0f with a 16-bit code): they cross NVLink as 4 bytes instead of 8.
__global__ void k_xpack(const int2* __restrict__ cells, unsigned* __restrict__ q, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int2 c = cells[i];
        q[i] = ((unsigned)c.x & 0xffffu) | (__float2uint_rn(__fmul_rn(__int_as_float(c.y), 65535.0f)) << 16);
    }
}
__global__ void k_xunpack(const unsigned* __restrict__ q, int2* __restrict__ cells, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned v = q[i];
        cells[i] = make_int2((int)(v & 0xffffu), __float_as_int(__fdiv_rn((float)(v >> 16), 65535.0f)));   // SparseEntry.h:31
    }
}

// Second half of an exchange started with mlp_exchange_begin: wait for the copies and fill the statistics.
extern "C" int mlp_exchange_end(mlp_ctx* ctx) {
    if (!ctx) return MLP_E_ARG;
    if (!ctx->exch_pending) return MLP_OK;
    cudaSetDevice(ctx->device);
    ctx->exch_pending = false;
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats = mlp_stage_stats{};
    ctx->stats.ms_total = ms;
    ctx->stats.nnz = (int64_t)(ctx->exch_total / 2);
    ctx->stats.launches = 1;
    return MLP_OK;
}

extern "C" int mlp_exchange(mlp_ctx* ctx) {
    const int rc = mlp_exchange_begin(ctx);
    return rc != MLP_OK ? rc : mlp_exchange_end(ctx);
}

// Enqueue the whole exchange and return without waiting for the cell broadcasts.  The distances are complete as soon as
// their all-reduce has run (event ev_dist): mlp_get_distances waits for that event only, so the host guide tree can be built
// while the sparse cells are still travelling over NVLink.
extern "C" int mlp_exchange_begin(mlp_ctx* ctx) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->comm_world <= 1) return MLP_OK;
    if (ctx->exch_pending) { const int rc0 = mlp_exchange_end(ctx); if (rc0 != MLP_OK) return rc0; }
    if (!ctx->nccl_comm || !ctx->have_sets) { ctx->err = "mlp_comm_init and a posterior/relax stage must come first"; return MLP_E_STATE; }
    // The merge is a SUM over ranks of tables whose foreign slots are zero: it is only correct on a set a sharded stage has
    // just produced.  A set that is already complete (second call without a stage in between) is left alone.
    if (!ctx->set_partial) return MLP_OK;
    cudaSetDevice(ctx->device);
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int W = ctx->comm_world, R = ctx->comm_rank, n = ctx->n;
    const int cur = ctx->cur, oth = 1 - ctx->cur;
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], st));
    // 1. how many cells does every rank hold?
    if (!ctx->d_xused) CK(cudaMalloc(&ctx->d_xused, 64 * sizeof(unsigned long long)));   // kept for the life of the context
    if (W > 64) { ctx->err = "more than 64 ranks"; return MLP_E_ARG; }
    unsigned long long* d_used = ctx->d_xused;
    NK(g_nccl.AllGather(ctx->set[cur].cursor, d_used, 1, ncclUint64, comm, st));
    std::vector<unsigned long long> used(W);
    CK(cudaMemcpyAsync(used.data(), d_used, (size_t)W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));   // the slab sizes decide the broadcast arguments: one host round trip per exchange
    std::vector<long long> base(W + 1, 0);
    for (int r = 0; r < W; ++r) base[r + 1] = base[r] + (long long)used[r];
    const long long total = base[W];
    if (total + 1024 > ctx->set[oth].cap) { int rc = grow_cells(ctx, oth, total + 1024, 0); if (rc != MLP_OK) return rc; }
    // 2. shift the offsets of my pairs to their place in the common pool
    if (!ctx->owned.empty()) {
        int rc = MLP_OK;
        if (ctx->owned.size() > ctx->tasks_cap) {
            free_dev(ctx->d_tasks); free_dev(ctx->d_pout); ctx->d_tasks = nullptr; ctx->d_pout = nullptr;
            const size_t cap = ctx->owned.size() + 64;
            CK(cudaMalloc(&ctx->d_tasks, cap * sizeof(PairTask)));
            CK(cudaMalloc(&ctx->d_pout, cap * sizeof(PairOut)));
            ctx->tasks_cap = cap;
        }
        (void)rc;
        CK(cudaMemcpyAsync(ctx->d_tasks, ctx->owned.data(), ctx->owned.size() * sizeof(PairTask), cudaMemcpyHostToDevice, st));
        const int nt = (int)ctx->owned.size();
        k_shift_offsets<<<(nt + 255) / 256, 256, 0, st>>>(ctx->d_tasks, nt, n, ctx->set[cur].nz_off, base[R]);
        CK(cudaGetLastError());
    }
    // 3. fixed-layout tables: sum == union
    // distances: only the posterior stage writes them (owned pairs, zeros elsewhere); after a relaxation every rank still holds
    // the complete matrix from the first exchange and summing it again would multiply it by the world size
    if (ctx->dist_partial) NK(g_nccl.AllReduce(ctx->d_dist, ctx->d_dist, (size_t)n * n, ncclFloat32, ncclSum, comm, st));
    ctx->dist_partial = false;
    ctx->set_partial = false;
    CK(cudaEventRecord(ctx->ev_dist, st));
    NK(g_nccl.AllReduce(ctx->set[cur].rp_pool, ctx->set[cur].rp_pool, (size_t)ctx->rp_total, ncclInt32, ncclSum, comm, st));
    NK(g_nccl.AllReduce(ctx->set[cur].nz_cnt, ctx->set[cur].nz_cnt, (size_t)n * n, ncclInt32, ncclSum, comm, st));
    NK(g_nccl.AllReduce(ctx->set[cur].nz_off, ctx->set[cur].nz_off, (size_t)n * n, ncclInt64, ncclSum, comm, st));
    // 4. cells: every rank broadcasts its slab into the common pool
    // Packed cells halve the wire bytes but add a pass over the whole pool (k_xunpack).  Measured per step at 1000 x 300: 2 GPUs
    // 52 -> 40 ms, 8 GPUs (NVSwitch, 1/8 of the set per sender) 55 -> 72 ms.  So: packed on two GPUs, raw otherwise;
    // MLP_XCHG_PACKED=0/1 overrides.
    const char* xenv = getenv("MLP_XCHG_PACKED");
    const bool packed = ctx->flavour_of_set == MLP_QP && (xenv ? atoi(xenv) != 0 : W == 2);
    if (packed) {
        // 4 bytes per cell on the wire (pack own slab -> broadcast in place -> unpack the whole pool)
        if ((size_t)total + 8 > ctx->xq_cap) {
            free_dev(ctx->d_xq); ctx->d_xq = nullptr; ctx->xq_cap = 0;
            CK(cudaMalloc(&ctx->d_xq, ((size_t)total + total / 8 + 64) * sizeof(unsigned)));
            ctx->xq_cap = (size_t)total + total / 8 + 56;
        }
        unsigned* q = ctx->d_xq;
        const int grid = ctx->num_sms * 8;
        if (used[R]) { k_xpack<<<grid, 256, 0, st>>>(ctx->set[cur].cells, q + base[R], (long long)used[R]); CK(cudaGetLastError()); }
        NK(g_nccl.GroupStart());
        for (int r = 0; r < W; ++r)
            if (used[r]) NK(g_nccl.Broadcast(q + base[r], q + base[r], (size_t)used[r], ncclUint32, r, comm, st));
        NK(g_nccl.GroupEnd());
        if (total) { k_xunpack<<<grid, 256, 0, st>>>(q, ctx->set[oth].cells, total); CK(cudaGetLastError()); }
    } else {
        // 8-byte cells sent as uint64 (cpnp: full floats)
        NK(g_nccl.GroupStart());
        for (int r = 0; r < W; ++r)
            if (used[r])
                NK(g_nccl.Broadcast(ctx->set[cur].cells, ctx->set[oth].cells + base[r], (size_t)used[r], ncclUint64, r, comm, st));
        NK(g_nccl.GroupEnd());
    }
    ctx->exch_total = (unsigned long long)total;     // lives in the context: the copy below may read it after this call returns
    CK(cudaMemcpyAsync(ctx->set[cur].cursor, &ctx->exch_total, sizeof(ctx->exch_total), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[1], st));
    std::swap(ctx->set[cur].cells, ctx->set[oth].cells);   // everything enqueued later on this stream sees the common pool
    std::swap(ctx->set[cur].cap, ctx->set[oth].cap);
    ctx->exch_pending = true;
    return MLP_OK;
}
