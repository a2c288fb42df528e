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
    if (ctx->rb_set >= 0) { const int rcr = mlp_get_csr_packed_end(ctx); if (rcr != MLP_OK) return rcr; }   // the gather lands in the other set's cell pool
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

// ------------------------------------------------------------------------------------------------ selective exchange (QuickProbs flavour)
// QuickProbs' consistency only ever reads S_xz with subtree distance d[x][z] <= selectivity (ConsistencyStage.cpp:181-216): about a
// fifth of the matrices at N = 1000.  Once the guide tree is known, every rank ships just those of its matrices to the others:
//   1. mlp_exchange_distances: all-reduce of the distance matrix (the host tree needs all of it);
//   2. mlp_exchange_needed(seldist, selectivity): the per-matrix cell counts are gathered (sum over ranks of a zero-padded table),
//      every rank derives the SAME import list on the host (owner-major, pair order of the cost-sorted list, (a,b) then (b,a)),
//      packs its own needed matrices into its slice of an import region behind its local cells, and one grouped in-place
//      ncclBroadcast per rank fills the region everywhere; row pointers travel the same way and are scattered into the fixed
//      row-pointer layout.  Matrices nobody can read stay where they were computed.
// The relaxation then runs on the owned pairs; its output is again sharded (mlp_exchange gathers it when a tail needs the whole set).
namespace {
struct XImport { long long slot; long long cell_off; long long rp_off; int cnt; int rows; };   // one needed ordered matrix

__global__ void k_xgather(const XImport* __restrict__ imp, int n_imp, const long long* __restrict__ nz_off, const long long* __restrict__ rp_off,
                          const int2* __restrict__ cells_src, int2* __restrict__ region, const int* __restrict__ rp_pool, int* __restrict__ rp_region) {
    for (int m = blockIdx.x; m < n_imp; m += gridDim.x) {           // one CTA per matrix I own
        const XImport e = imp[m];
        const int2* src = cells_src + nz_off[e.slot];
        for (int k = threadIdx.x; k < e.cnt; k += blockDim.x) region[e.cell_off + k] = src[k];
        const int* rsrc = rp_pool + rp_off[e.slot];
        for (int k = threadIdx.x; k < e.rows; k += blockDim.x) rp_region[e.rp_off + k] = rsrc[k];
    }
}
__global__ void k_ximport(const XImport* __restrict__ imp, int n_imp, long long* __restrict__ nz_off, int* __restrict__ nz_cnt,
                          const long long* __restrict__ rp_off, long long region_base, int* __restrict__ rp_pool, const int* __restrict__ rp_region) {
    for (int m = blockIdx.x; m < n_imp; m += gridDim.x) {           // one CTA per matrix another rank owns
        const XImport e = imp[m];
        if (threadIdx.x == 0) { nz_off[e.slot] = region_base + e.cell_off; nz_cnt[e.slot] = e.cnt; }
        int* rdst = rp_pool + rp_off[e.slot];
        for (int k = threadIdx.x; k < e.rows; k += blockDim.x) rdst[k] = rp_region[e.rp_off + k];
    }
}
}  // namespace

extern "C" int mlp_exchange_distances(mlp_ctx* ctx) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->comm_world <= 1) return MLP_OK;
    if (ctx->exch_pending) { const int rc0 = mlp_exchange_end(ctx); if (rc0 != MLP_OK) return rc0; }
    if (!ctx->nccl_comm || !ctx->have_sets) { ctx->err = "mlp_comm_init and the posterior stage must come first"; return MLP_E_STATE; }
    if (!ctx->dist_partial) return MLP_OK;
    cudaSetDevice(ctx->device);
    NK(g_nccl.AllReduce(ctx->d_dist, ctx->d_dist, (size_t)ctx->n * ctx->n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dist_partial = false;
    return MLP_OK;
}

extern "C" int mlp_exchange_needed(mlp_ctx* ctx, const float* seldist_nxn, float selectivity) {
    if (!ctx || !seldist_nxn) return MLP_E_ARG;
    if (ctx->comm_world <= 1) return MLP_OK;
    if (ctx->exch_pending) { const int rc0 = mlp_exchange_end(ctx); if (rc0 != MLP_OK) return rc0; }
    if (!ctx->nccl_comm || !ctx->have_sets) { ctx->err = "mlp_comm_init and the posterior stage must come first"; return MLP_E_STATE; }
    if (!ctx->set_partial) return MLP_OK;
    { const int rcd = mlp_exchange_distances(ctx); if (rcd != MLP_OK) return rcd; }
    cudaSetDevice(ctx->device);
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int W = ctx->comm_world, R = ctx->comm_rank, n = ctx->n;
    const size_t nn = (size_t)n * n;
    const int cur = ctx->cur;
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], st));
    // 1. every rank's cell counts (foreign slots are zero: sum == union), into a scratch table
    if (nn > ctx->xcnt_cap) {
        free_dev(ctx->d_xcnt); ctx->d_xcnt = nullptr;
        CK(cudaMalloc(&ctx->d_xcnt, (nn + 64) * sizeof(int)));
        ctx->xcnt_cap = nn + 64;
    }
    NK(g_nccl.AllReduce(ctx->set[cur].nz_cnt, ctx->d_xcnt, nn, ncclInt32, ncclSum, comm, st));
    std::vector<int> cnt(nn);
    unsigned long long used_local = 0;
    CK(cudaMemcpyAsync(cnt.data(), ctx->d_xcnt, nn * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&used_local, ctx->set[cur].cursor, sizeof(used_local), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // 2. the import list, identical on every rank: owner-major, cost-sorted pair order, (a,b) before (b,a)
    std::vector<XImport> mine, theirs;
    std::vector<long long> cbase(W + 1, 0), rbase(W + 1, 0);
    long long coff = 0, roff = 0;
    for (int q = 0; q < W; ++q) {
        cbase[q] = coff; rbase[q] = roff;
        for (size_t k = q; k < ctx->all_pairs.size(); k += W) {
            const PairTask& t = ctx->all_pairs[k];
            for (int o = 0; o < 2; ++o) {
                const int a = o ? t.b : t.a, b = o ? t.a : t.b;
                const size_t slot = (size_t)a * n + b;
                if (!(seldist_nxn[slot] <= selectivity)) continue;
                XImport e; e.slot = (long long)slot; e.cell_off = coff; e.rp_off = roff; e.cnt = cnt[slot]; e.rows = ctx->len[a] + 2;
                (q == R ? mine : theirs).push_back(e);
                coff += (e.cnt + 1) & ~1;          // matrices stay 16-byte aligned inside the region (bulk copies of the relaxation)
                roff += (e.rows + 3) & ~3;
            }
        }
    }
    cbase[W] = coff; rbase[W] = roff;
    // 3. room: import region behind the local cells (even base), row-pointer region in the wire buffer
    const long long region_base = (long long)((used_local + 1) & ~1ull);
    if (region_base + coff + 1024 > ctx->set[cur].cap) { int rc = grow_cells(ctx, cur, region_base + coff + coff / 16 + 1024, used_local); if (rc != MLP_OK) return rc; }
    if ((size_t)roff + 64 > ctx->xq_cap) {
        free_dev(ctx->d_xq); ctx->d_xq = nullptr; ctx->xq_cap = 0;
        CK(cudaMalloc(&ctx->d_xq, ((size_t)roff + roff / 8 + 64) * sizeof(unsigned)));
        ctx->xq_cap = (size_t)roff + roff / 8 + 56;
    }
    const size_t n_list = mine.size() + theirs.size();
    if (n_list > ctx->ximp_cap) {
        free_dev(ctx->d_ximp); ctx->d_ximp = nullptr;
        CK(cudaMalloc(&ctx->d_ximp, (n_list + n_list / 4 + 64) * sizeof(XImport)));
        ctx->ximp_cap = n_list + n_list / 4 + 64;
    }
    XImport* d_mine = reinterpret_cast<XImport*>(ctx->d_ximp);
    XImport* d_theirs = d_mine + mine.size();
    if (!mine.empty()) CK(cudaMemcpyAsync(d_mine, mine.data(), mine.size() * sizeof(XImport), cudaMemcpyHostToDevice, st));
    if (!theirs.empty()) CK(cudaMemcpyAsync(d_theirs, theirs.data(), theirs.size() * sizeof(XImport), cudaMemcpyHostToDevice, st));
    int2* region = ctx->set[cur].cells + region_base;
    int* rp_region = reinterpret_cast<int*>(ctx->d_xq);
    if (!mine.empty()) {
        k_xgather<<<std::min<int>((int)mine.size(), ctx->num_sms * 8), 128, 0, st>>>(d_mine, (int)mine.size(), ctx->set[cur].nz_off, ctx->d_rp_off,
                                                                                   ctx->set[cur].cells, region, ctx->set[cur].rp_pool, rp_region);
        CK(cudaGetLastError());
    }
    // 4. one grouped in-place broadcast per rank: cells (as uint64) and row pointers
    NK(g_nccl.GroupStart());
    for (int q = 0; q < W; ++q) {
        const size_t nc = (size_t)(cbase[q + 1] - cbase[q]), nr = (size_t)(rbase[q + 1] - rbase[q]);
        if (nc) NK(g_nccl.Broadcast(region + cbase[q], region + cbase[q], nc, ncclUint64, q, comm, st));
        if (nr) NK(g_nccl.Broadcast(rp_region + rbase[q], rp_region + rbase[q], nr, ncclInt32, q, comm, st));
    }
    NK(g_nccl.GroupEnd());
    // 5. point the foreign slots at the region, scatter their row pointers into the fixed layout
    if (!theirs.empty()) {
        k_ximport<<<std::min<int>((int)theirs.size(), ctx->num_sms * 8), 128, 0, st>>>(d_theirs, (int)theirs.size(), ctx->set[cur].nz_off, ctx->set[cur].nz_cnt,
                                                                                     ctx->d_rp_off, region_base, ctx->set[cur].rp_pool, rp_region);
        CK(cudaGetLastError());
    }
    ctx->exch_total = (unsigned long long)(region_base + coff);
    CK(cudaMemcpyAsync(ctx->set[cur].cursor, &ctx->exch_total, sizeof(ctx->exch_total), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats = mlp_stage_stats{};
    ctx->stats.ms_total = ms;
    ctx->stats.nnz = coff;                         // cells in the import region (all ranks' needed matrices)
    ctx->stats.pairs = (int64_t)n_list;
    ctx->stats.launches = 2;
    ctx->set_partial = false;                      // every matrix a relaxation of the owned pairs can read is now present
    ctx->imported = true;
    ctx->own_cells = (long long)used_local;
    return MLP_OK;
}
