// C ABI implementation (include/mlprobs_b200.h): host driver of the posterior and consistency stages.
// One context = one GPU + one stream.  Pairs are cost-sorted, sharded (rank/world) and processed in batches
// whose dense DP layers fit the scratch budget; every batch runs the sweep kernels back to back on the stream.
#include "../../include/mlprobs_b200.h"
#include "posterior.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.h"

void free_dev(void* p) { if (p) cudaFree(p); }

static void release_sets(mlp_ctx* ctx) {
    for (int s = 0; s < 2; ++s) {
        free_dev(ctx->set[s].rp_pool); free_dev(ctx->set[s].nz_off); free_dev(ctx->set[s].nz_cnt);
        free_dev(ctx->set[s].cells); free_dev(ctx->set[s].cursor);
        ctx->set[s] = CsrSetDev{};
    }
    free_dev(ctx->d_rp_off); ctx->d_rp_off = nullptr;
    free_dev(ctx->d_dist); ctx->d_dist = nullptr;
    ctx->rp_cap = ctx->nn_cap = 0;
    ctx->have_sets = false;
}

static void release_launch_scratch(mlp_ctx* ctx) {
    free_dev(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    free_dev(ctx->d_tasks); ctx->d_tasks = nullptr; free_dev(ctx->d_pout); ctx->d_pout = nullptr; ctx->tasks_cap = 0;
    free_dev(ctx->d_stage); ctx->d_stage = nullptr; ctx->stage_warps = 0; ctx->stage_cap = 0;
    free_dev(ctx->d_tfill); ctx->d_tfill = nullptr; ctx->tfill_warps = 0;
    free_dev(ctx->d_rowexp); ctx->d_rowexp = nullptr; ctx->rowexp_cap = 0;
    free_dev(ctx->d_rowaux); ctx->d_rowaux = nullptr; ctx->rowaux_cap = 0;
    free_dev(ctx->d_edge); ctx->d_edge = nullptr; ctx->edge_warps = 0;
    free_dev(ctx->d_wk); ctx->d_wk = nullptr; ctx->wk_warps = 0;
}

// a split read-back must have finished before the set it reads (or any pool) is touched again
#define END_READBACK(ctx) do { if ((ctx)->rb_set >= 0) { const int rcr__ = mlp_get_csr_packed_end(ctx); if (rcr__ != MLP_OK) return rcr__; } } while (0)

extern "C" int mlp_create(int device, mlp_ctx** out) {
    if (!out) return MLP_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return MLP_E_NO_DEVICE;
    if (device < 0 || device >= count) return MLP_E_ARG;
    mlp_ctx* ctx = new mlp_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return MLP_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return MLP_E_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MLP_E_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MLP_E_CUDA; }
    cudaEventCreate(&ctx->ev[0]); cudaEventCreate(&ctx->ev[1]);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_dist, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->stream_rb, cudaStreamNonBlocking); cudaEventCreateWithFlags(&ctx->ev_rb, cudaEventDisableTiming);
    if (const char* e = getenv("MLP_OVERLAP")) ctx->overlap = atoi(e);
    if (const char* e = getenv("MLP_BPS_PART")) ctx->bps_part = atoi(e);
    if (const char* e = getenv("MLP_BPS_HMM")) ctx->bps_hmm = atoi(e);
    if (cudaMalloc(&ctx->d_counter, 16 * sizeof(int)) != cudaSuccess || cudaMalloc(&ctx->d_err, sizeof(int)) != cudaSuccess) {
        delete ctx; return MLP_E_CUDA;
    }
    cudaMemset(ctx->d_err, 0, sizeof(int));
    if (cudaHostAlloc((void**)&ctx->h_flags, 4 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->d_hflags, ctx->h_flags, 0) != cudaSuccess) { delete ctx; return MLP_E_CUDA; }
    *out = ctx;
    return MLP_OK;
}

extern "C" void mlp_destroy(mlp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->stream_rb) { cudaStreamSynchronize(ctx->stream_rb); cudaStreamDestroy(ctx->stream_rb); }
    if (ctx->ev_rb) cudaEventDestroy(ctx->ev_rb);
    free_dev(ctx->d_pack); free_dev(ctx->d_rs_off); free_dev(ctx->d_rs_tasks);
    if (ctx->h_rs_stage) cudaFreeHost(ctx->h_rs_stage);
    release_sets(ctx);
    release_launch_scratch(ctx);
    if (ctx->tail_prov && ctx->tail_prov_free) ctx->tail_prov_free(ctx->tail_prov);
    free_dev(ctx->d_res); free_dev(ctx->d_seq_off);
    free_dev(ctx->d_match); free_dev(ctx->d_ins); free_dev(ctx->d_sub);
    free_dev(ctx->d_counter); free_dev(ctx->d_err);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    free_dev(ctx->d_sdigest); free_dev(ctx->d_len); free_dev(ctx->d_tree); free_dev(ctx->d_weights); free_dev(ctx->d_seldist); free_dev(ctx->d_xq); free_dev(ctx->d_xused); free_dev(ctx->d_xcnt); free_dev(ctx->d_ximp); free_dev(ctx->d_relax_tasks);
    if (ctx->ev[0]) cudaEventDestroy(ctx->ev[0]);
    if (ctx->ev[1]) cudaEventDestroy(ctx->ev[1]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_dist) cudaEventDestroy(ctx->ev_dist);
    delete ctx;
}

extern "C" const char* mlp_last_error(const mlp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int mlp_configure(mlp_ctx* ctx, int64_t scratch_bytes, int64_t cell_capacity) {
    if (!ctx || scratch_bytes < 0 || cell_capacity < 0) return MLP_E_ARG;
    ctx->scratch_budget = scratch_bytes;
    ctx->cell_capacity_req = cell_capacity;
    return MLP_OK;
}

// Streamed posterior stage for families whose sparse set does not fit HBM (BASELINE config #5: 4,000 x 500, 16 M matrices).
// Between _begin and _end mlp_posterior_all_pairs(MLP_QP) finishes every batch on the spot -- the re-quantisation a consistency
// repetition applies to a matrix whose pair accepts no third sequence (`reps` of them), then the per-matrix digest -- and reuses the
// cell pool for the next batch.  Distances are complete afterwards; the matrices of the pairs that DO accept third sequences (same
// <= selectivity-leaf subtree) are recomputed by an ordinary stage restricted to them (mlp_restrict_pairs) and relaxed as usual.
extern "C" int mlp_stream_begin(mlp_ctx* ctx, int reps) {
    if (!ctx || reps < 1 || reps > 8) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    const int n = ctx->n;
    if (n < 2) { ctx->err = "set sequences first"; return MLP_E_STATE; }
    // Only the last repetition runs with the 1e-5 cutoff (ConsistencyStage.cpp:108-115); an earlier one uses 0.01 and would drop the
    // cells whose quantised value fell just below it.  QuickProbs runs ONE repetition above 50 sequences (Configuration.cpp:100-103),
    // and a family small enough for two has no pair outside a 200-leaf subtree, i.e. nothing to stream.
    if (reps != 1) { ctx->err = "the streamed stage implements the single consistency repetition QuickProbs runs above 50 sequences"; return MLP_E_UNSUPPORTED; }
    const long long nn = (long long)n * n;
    if (nn > ctx->sdigest_cap) {
        free_dev(ctx->d_sdigest); ctx->d_sdigest = nullptr; ctx->sdigest_cap = 0;
        CK(cudaMalloc(&ctx->d_sdigest, (size_t)nn * sizeof(unsigned long long)));
        ctx->sdigest_cap = nn;
    }
    if (n > ctx->len_cap) {
        free_dev(ctx->d_len); ctx->d_len = nullptr; ctx->len_cap = 0;
        CK(cudaMalloc(&ctx->d_len, (size_t)n * sizeof(int)));
        ctx->len_cap = n;
    }
    CK(cudaMemsetAsync(ctx->d_sdigest, 0, (size_t)nn * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_len, ctx->len.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream_mode = true; ctx->stream_reps = reps;
    return MLP_OK;
}

// Ends streamed mode; per_matrix_nn (n*n, may be NULL) receives the digests of the streamed matrices in mlp_set_digest's format.
extern "C" int mlp_stream_end(mlp_ctx* ctx, uint64_t* per_matrix_nn) {
    if (!ctx) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    ctx->stream_mode = false;
    if (per_matrix_nn) {
        if (!ctx->d_sdigest) return MLP_E_STATE;
        CK(cudaMemcpyAsync(per_matrix_nn, ctx->d_sdigest, (size_t)ctx->n * ctx->n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MLP_OK;
}

// Restricts this rank's shard to the pairs whose subtree-size distance is within the selectivity -- the only pairs that accept third
// sequences (ConsistencyStage.cpp:181-216) and the only matrices a consistency repetition reads.  mlp_set_shard restores the shard.
extern "C" int mlp_restrict_pairs(mlp_ctx* ctx, const float* seldist_nxn, float selectivity) {
    if (!ctx || !seldist_nxn) return MLP_E_ARG;
    const int n = ctx->n;
    std::vector<PairTask> keep;
    for (const PairTask& t : ctx->owned)
        if (seldist_nxn[(size_t)t.a * n + t.b] <= selectivity) keep.push_back(t);
    ctx->owned.swap(keep);
    ctx->restricted = true;
    ctx->relax_tasks.clear();
    ctx->relax_tasks_on_device = false;
    return MLP_OK;
}

extern "C" int mlp_set_tables(mlp_ctx* ctx, const mlp_hmm_tables* hmm, const mlp_part_tables* part) {
    if (!ctx || !hmm || !part) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    if (!(part->tgo == 1.0 && part->tge == 1.0)) {
        ctx->err = "terminal gap terms other than exp(0)=1 are not supported (both references hard-code them)";
        return MLP_E_UNSUPPORTED;
    }
    ctx->hmm = *hmm; ctx->part = *part;
    if (!ctx->d_match) {
        CK(cudaMalloc(&ctx->d_match, 676 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_ins, 32 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_sub, 676 * sizeof(double)));
    }
    CK(cudaMemcpyAsync(ctx->d_match, &hmm->match[0][0], 676 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ins, hmm->ins, 26 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_sub, &part->sub[0][0], 676 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DevScalars s;
    for (int q = 0; q < 5; ++q) {
        s.init[q] = hmm->init[q];
        s.t0q[q] = hmm->trans[0][q];
        s.tqq[q] = hmm->trans[q][q];
        s.tq0[q] = hmm->trans[q][0];
    }
    s.lt00 = hmm->ltrans[0][0]; s.lt01 = hmm->ltrans[0][1]; s.lt02 = hmm->ltrans[0][2];
    s.lt10 = hmm->ltrans[1][0]; s.lt11 = hmm->ltrans[1][1];
    s.lt20 = hmm->ltrans[2][0]; s.lt22 = hmm->ltrans[2][2];
    s.r = hmm->rtrans[1];
    s.r2 = 2 * hmm->rtrans[1];
    s.go = part->go; s.ge = part->ge;
    CK(posterior_set_scalars(s, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_tables = true;
    ctx->stats.h2d_bytes += 676 * 4 + 26 * 4 + 676 * 8 + (int64_t)sizeof(DevScalars);
    return MLP_OK;
}

static void pair_geometry(int L2, int& C, int& nb) {
    const int cols = L2 + 2;   // columns 0..L2 plus the virtual column L2+1 of the reverse sweeps
    // Columns per lane are capped at 8 (kCmaxLimit, the unroll width of the sweep kernels): two column blocks of 5 columns per lane
    // instead of one of 10 halve the per-warp band in shared memory, which lifts the FP64 reverse sweep from 12 to 24 resident
    // warps per SM (measured at 1000 x 300: part_rev 284 -> 228 ms, final 348 -> 330, hmm sweeps +25 ms).  MLP_CMAX overrides.
    static const int lim = []() { const char* e = getenv("MLP_CMAX"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= kCmaxLimit) ? v : 8; }();
    nb = (cols + 32 * lim - 1) / (32 * lim);
    C = (cols + 32 * nb - 1) / (32 * nb);
}

static void build_sorted_pairs(int n, const int32_t* len, std::vector<PairTask>& out) {
    out.clear();
    out.reserve((size_t)n * (n - 1) / 2);
    int pidx = 0;
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) {
            PairTask t;
            t.a = a; t.b = b; t.L1 = len[a]; t.L2 = len[b];
            pair_geometry(t.L2, t.C, t.nb);
            t.pidx = pidx++; t.flags = 0; t.off = 0;
            out.push_back(t);
        }
    // grouped by columns per lane (the register-band kernels are compiled per C: one launch per group), cost-sorted (largest
    // first) inside a group for load balance; stable, so ties keep row-major pair order
    std::stable_sort(out.begin(), out.end(), [](const PairTask& x, const PairTask& y) {
        if (x.C != y.C) return x.C > y.C;
        const long long cx = (long long)x.nb * (x.L1 + 32) * x.C, cy = (long long)y.nb * (y.L1 + 32) * y.C;
        return cx > cy;
    });
}

extern "C" int mlp_shard_pairs(int n, const int32_t* len, int rank, int world, int32_t* pairs_out, int64_t* count) {
    if (n < 2 || !len || world < 1 || rank < 0 || rank >= world || !count) return MLP_E_ARG;
    std::vector<PairTask> all;
    build_sorted_pairs(n, len, all);
    int64_t c = 0;
    for (size_t k = 0; k < all.size(); ++k)
        if ((int)(k % world) == rank) {
            if (pairs_out) { pairs_out[2 * c] = all[k].a; pairs_out[2 * c + 1] = all[k].b; }
            ++c;
        }
    *count = c;
    return MLP_OK;
}

// Host utility (no GPU work): the pairs of mlp_shard_pairs(rank, world) that mlp_restrict_pairs keeps -- the multi-GPU contract of the
// streamed flow can be checked without a device (tests/test_sharding_gloo.py).
extern "C" int mlp_shard_pairs_within(int n, const int32_t* len, int rank, int world, const float* seldist_nxn, float selectivity,
                                      int32_t* pairs_out, int64_t* count) {
    if (n < 2 || !len || world < 1 || rank < 0 || rank >= world || !count || !seldist_nxn) return MLP_E_ARG;
    std::vector<PairTask> all;
    build_sorted_pairs(n, len, all);
    int64_t c = 0;
    for (size_t k = 0; k < all.size(); ++k)
        if ((int)(k % world) == rank && seldist_nxn[(size_t)all[k].a * n + all[k].b] <= selectivity) {
            if (pairs_out) { pairs_out[2 * c] = all[k].a; pairs_out[2 * c + 1] = all[k].b; }
            ++c;
        }
    *count = c;
    return MLP_OK;
}

extern "C" int mlp_set_sequences(mlp_ctx* ctx, int n, const int32_t* len, const uint8_t* residues) {
    if (!ctx || n < 2 || !len || !residues) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    // validate everything before touching the context: a rejected family must leave the previous one intact
    long long tot = 0;
    for (int i = 0; i < n; ++i) {
        if (len[i] < 1 || len[i] > 65535) { ctx->err = "sequence length must be in 1..65535"; return MLP_E_ARG; }
        tot += len[i];
    }
    for (long long k = 0; k < tot; ++k)
        if (residues[k] < 'A' || residues[k] > 'Z') { ctx->err = "residues must be upper-case letters"; return MLP_E_ARG; }
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }
    // same family shape as before (same n and lengths): the pooled layout is unchanged, keep the device pools
    const bool same_layout = ctx->have_sets && ctx->n == n && std::equal(len, len + n, ctx->len.begin());
    if (!same_layout) ctx->have_sets = false;        // the pools themselves stay: ensure_sets re-uses them when the new family fits
    const bool same_shape = ctx->n == n && (int)ctx->len.size() == n && std::equal(len, len + n, ctx->len.begin()) && !ctx->all_pairs.empty();
    if (!same_shape) END_READBACK(ctx);              // a different family may re-allocate the pools a split read-back is copying from
    ctx->flavour_of_set = -1;
    ctx->set_partial = ctx->dist_partial = false;
    ctx->stream_mode = false;                        // a streamed stage belongs to the family it was begun for
    ctx->n = n;
    ctx->len.assign(len, len + n);
    ctx->seq_off.resize(n);
    tot = 0;
    for (int i = 0; i < n; ++i) { ctx->seq_off[i] = tot; tot += len[i]; }
    ctx->total_res = tot;
    std::vector<uint8_t> codes(tot + 16, 0);
    for (long long k = 0; k < tot; ++k) codes[k] = (uint8_t)(residues[k] - 'A');
    ctx->codes_h = codes;
    if (tot + 16 > ctx->res_cap) {
        free_dev(ctx->d_res); ctx->d_res = nullptr; ctx->res_cap = 0;
        const long long cap = tot + tot / 2 + 4096;
        CK(cudaMalloc(&ctx->d_res, cap));
        ctx->res_cap = cap;
    }
    if (n > ctx->seqoff_cap) {
        free_dev(ctx->d_seq_off); ctx->d_seq_off = nullptr; ctx->seqoff_cap = 0;
        const int cap = n + n / 2 + 64;
        CK(cudaMalloc(&ctx->d_seq_off, (size_t)cap * sizeof(long long)));
        ctx->seqoff_cap = cap;
    }
    CK(cudaMemcpy(ctx->d_res, codes.data(), tot + 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_seq_off, ctx->seq_off.data(), n * sizeof(long long), cudaMemcpyHostToDevice));
    ctx->stats.h2d_bytes += tot + 16 + n * (int64_t)sizeof(long long);
    if (same_shape) {                 // same lengths as the family before (a re-submitted family): pair list, shard and layout stand
        if (ctx->restricted) return mlp_set_shard(ctx, ctx->rank, ctx->world);   // ... but not a shard cut down by mlp_restrict_pairs
        return MLP_OK;
    }
    build_sorted_pairs(n, len, ctx->all_pairs);
    // fixed row-pointer layout: ordered pair (a,b) owns len[a]+2 ints
    ctx->rp_off_h.assign((size_t)n * n, 0);
    long long rp = 0;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            ctx->rp_off_h[(size_t)a * n + b] = rp;
            if (a != b) rp += len[a] + 2;
        }
    ctx->rp_total = rp;
    // a context that belongs to a communicator (or was sharded by hand) keeps its shard for the next family
    if (ctx->nccl_comm) { ctx->rank = ctx->comm_rank; ctx->world = ctx->comm_world; }
    return mlp_set_shard(ctx, ctx->rank, ctx->world);
}

extern "C" int mlp_set_shard(mlp_ctx* ctx, int rank, int world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return MLP_E_ARG;
    ctx->rank = rank; ctx->world = world;
    ctx->restricted = false;
    ctx->owned.clear();
    for (size_t k = 0; k < ctx->all_pairs.size(); ++k)
        if ((int)(k % world) == rank) ctx->owned.push_back(ctx->all_pairs[k]);
    ctx->relax_tasks.clear();
    ctx->relax_tasks_on_device = false;
    return MLP_OK;
}

static int ensure_sets(mlp_ctx* ctx) {
    if (ctx->have_sets) return MLP_OK;
    const int n = ctx->n;
    long long cap = ctx->cell_capacity_req;
    if (cap <= 0) {
        long long s = 0;
        for (const PairTask& t : ctx->all_pairs) s += std::min(t.L1, t.L2);
        cap = 16 * s + (1 << 20);   // both orientations, ~8 cells per row of head-room
    }
    const long long nn = (long long)n * n;
    if (!ctx->d_rp_off || nn > ctx->nn_cap || ctx->rp_total + 8 > ctx->rp_cap) {
        // first family, or one that does not fit the pools kept from the previous families: re-allocate with head-room
        release_sets(ctx);
        const long long nn_cap = nn + std::min<long long>(nn / 4, 1 << 22) + 64, rp_cap = ctx->rp_total + std::min<long long>(ctx->rp_total / 4, 1LL << 28) + 64;   // rp_cap includes the 8 ints a 16-byte aligned bulk copy may read past the end
        CK(cudaMalloc(&ctx->d_rp_off, (size_t)nn_cap * sizeof(long long)));
        for (int s = 0; s < 2; ++s) {
            CK(cudaMalloc(&ctx->set[s].rp_pool, (size_t)rp_cap * sizeof(int)));
            CK(cudaMalloc(&ctx->set[s].nz_off, (size_t)nn_cap * sizeof(long long)));
            CK(cudaMalloc(&ctx->set[s].nz_cnt, (size_t)nn_cap * sizeof(int)));
            CK(cudaMalloc(&ctx->set[s].cursor, sizeof(unsigned long long)));
        }
        CK(cudaMalloc(&ctx->d_dist, (size_t)nn_cap * sizeof(float)));
        ctx->nn_cap = nn_cap;
        ctx->rp_cap = rp_cap;
    }
    CK(cudaMemcpy(ctx->d_rp_off, ctx->rp_off_h.data(), (size_t)n * n * sizeof(long long), cudaMemcpyHostToDevice));
    ctx->stats.h2d_bytes += (int64_t)n * n * 8;
    for (int s = 0; s < 2; ++s) {
        const long long this_cap = (s == 0) ? cap : 1024;   // the relax output pool is sized when mlp_relax runs
        if (!ctx->set[s].cells || ctx->set[s].cap < this_cap) {
            free_dev(ctx->set[s].cells); ctx->set[s].cells = nullptr; ctx->set[s].cap = 0;
            CK(cudaMalloc(&ctx->set[s].cells, ((size_t)this_cap + 4) * sizeof(int2)));
            ctx->set[s].cap = this_cap;
        }
        CK(cudaMemsetAsync(ctx->set[s].rp_pool, 0, (size_t)ctx->rp_total * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->set[s].nz_off, 0, (size_t)n * n * sizeof(long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->set[s].nz_cnt, 0, (size_t)n * n * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->set[s].cursor, 0, sizeof(unsigned long long), ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->d_dist, 0, (size_t)n * n * sizeof(float), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_sets = true;
    ctx->cur = 0;
    return MLP_OK;
}

// The context's stream is cudaStreamNonBlocking: a legacy-stream cudaMemcpy is NOT ordered against work queued on it.
// Every host read of a set's cursor goes through the stream itself.
__global__ void k_publish(const unsigned long long* cursor, const int* err, volatile unsigned long long* host_words) {
    if (cursor) host_words[0] = *cursor;
    if (err) host_words[1] = (unsigned long long)(unsigned)*err;
    __threadfence_system();
}
// cursor of a set and / or the error word, through the host-mapped words (no DMA copy, see ctx.h)
int read_words(mlp_ctx* ctx, int which, unsigned long long* cursor_out, int* err_out) {
    k_publish<<<1, 1, 0, ctx->stream>>>(cursor_out ? ctx->set[which].cursor : nullptr, err_out ? ctx->d_err : nullptr, ctx->d_hflags);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    if (cursor_out) *cursor_out = ctx->h_flags[0];
    if (err_out) *err_out = (int)ctx->h_flags[1];
    return MLP_OK;
}
int read_cursor(mlp_ctx* ctx, int which, unsigned long long* out) { return read_words(ctx, which, out, nullptr); }
#define END_EXCHANGE(ctx) do { if ((ctx)->exch_pending) { const int rce__ = mlp_exchange_end(ctx); if (rce__ != MLP_OK) return rce__; } } while (0)

// Re-allocate the cell pool of one set to new_cap cells, keeping the first `keep` cells.
int grow_cells(mlp_ctx* ctx, int which, long long new_cap, unsigned long long keep) {
    int2* fresh = nullptr;
    CK(cudaMalloc(&fresh, ((size_t)new_cap + 4) * sizeof(int2)));
    if (keep) CK(cudaMemcpy(fresh, ctx->set[which].cells, (size_t)keep * sizeof(int2), cudaMemcpyDeviceToDevice));
    cudaFree(ctx->set[which].cells);
    ctx->set[which].cells = fresh;
    ctx->set[which].cap = new_cap;
    return MLP_OK;
}

static int ensure_tasks(mlp_ctx* ctx, size_t ntasks) {
    if (ntasks <= ctx->tasks_cap) return MLP_OK;
    free_dev(ctx->d_tasks); free_dev(ctx->d_pout);
    ctx->d_tasks = nullptr; ctx->d_pout = nullptr;
    const size_t cap = ntasks + ntasks / 4 + 64;
    CK(cudaMalloc(&ctx->d_tasks, cap * sizeof(PairTask)));
    CK(cudaMalloc(&ctx->d_pout, cap * sizeof(PairOut)));
    ctx->tasks_cap = cap;
    return MLP_OK;
}

// per-warp buffers for `warps` resident warps
static int ensure_warp_buffers(mlp_ctx* ctx, long long warps, int maxL1, int maxL2, bool need_edge, int stage_mult) {
    const int want_stage = stage_mult * (maxL1 + 1);
    if (warps > ctx->stage_warps || want_stage > ctx->stage_cap) {
        free_dev(ctx->d_stage); ctx->d_stage = nullptr;
        ctx->stage_cap = std::max(want_stage, ctx->stage_cap);
        ctx->stage_warps = std::max(warps, ctx->stage_warps);
        CK(cudaMalloc(&ctx->d_stage, (size_t)ctx->stage_warps * ctx->stage_cap * sizeof(int4)));
    }
    const long long want_fill = ((maxL2 + 2 + 31) / 32) * 32;
    if (warps > ctx->tfill_warps || want_fill > ctx->tfill_stride) {
        free_dev(ctx->d_tfill); ctx->d_tfill = nullptr;
        ctx->tfill_stride = std::max(want_fill, ctx->tfill_stride);
        ctx->tfill_warps = std::max(warps, ctx->tfill_warps);
        CK(cudaMalloc(&ctx->d_tfill, (size_t)ctx->tfill_warps * ctx->tfill_stride * sizeof(int)));
    }
    if (need_edge) {
        const long long want_edge = (long long)(maxL1 + 1) * 5;   // elements (of 8 bytes: doubles or padded floats)
        if (warps > ctx->edge_warps || want_edge > ctx->edge_stride) {
            free_dev(ctx->d_edge); ctx->d_edge = nullptr;
            ctx->edge_stride = std::max(want_edge, ctx->edge_stride);
            ctx->edge_warps = std::max(warps, ctx->edge_warps);
            CK(cudaMalloc(&ctx->d_edge, (size_t)2 * ctx->edge_warps * ctx->edge_stride * sizeof(double)));   // second half: the partition sweeps when they run beside the HMM sweeps (MLP_OVERLAP)
        }
    }
    return MLP_OK;
}

struct KernelTimer {
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
    void begin(int k, cudaStream_t st) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st); spans.push_back({k, {a, b}});
    }
    void end(cudaStream_t st) { cudaEventRecord(spans.back().second.second, st); }
    void collect(mlp_stage_stats& s) {
        for (auto& sp : spans) {
            float ms = 0; cudaEventElapsedTime(&ms, sp.second.first, sp.second.second);
            if (sp.first >= 0 && sp.first < 8) s.ms_kernel[sp.first] += ms;
            cudaEventDestroy(sp.second.first); cudaEventDestroy(sp.second.second);
        }
        spans.clear();
    }
};

static int launch_one(mlp_ctx* ctx, int kernel, KArgs a, int ntasks, KernelTimer& kt, int stat_slot,
                      cudaStream_t st = nullptr, int bps_cap = 0, const std::vector<std::pair<int, int>>* cgroups = nullptr) {
    if (!st) st = ctx->stream;
    const int warps_per_cta = MLP_BLOCK / 32;
    a.counter = ctx->d_counter + (kernel & 15);      // every kernel id has its own work-queue head (kernels may overlap)
    if (const char* e = getenv("MLP_BPS_CAP")) { const int v = atoi(e); if (v > 0) bps_cap = bps_cap > 0 ? std::min(bps_cap, v) : v; }   // developer knob: resident CTAs per SM
    if (cgroups && posterior_c_available(kernel, a)) {
        // register-band kernels: one launch per run of tasks with the same columns-per-lane value (the batch is sorted by C)
        kt.begin(stat_slot, st);
        for (const auto& g : *cgroups) {
            const int begin = g.second;
            const int end = (&g == &cgroups->back()) ? ntasks : (&g)[1].second;
            const int C = g.first;
            int bps = std::min(posterior_c_max_blocks_per_sm(kernel, C, a), 16);
            if (bps_cap > 0) bps = std::min(bps, bps_cap);
            int grid = std::max(1, std::min(ctx->num_sms * bps, (end - begin + warps_per_cta - 1) / warps_per_cta));
            CK(cudaMemsetAsync(a.counter, 0, sizeof(int), st));
            a.task_begin = begin; a.ntasks = end;
            CK(posterior_c_launch(kernel, C, a, grid, st));
            ctx->stats.launches += 1;
        }
        kt.end(st);
        return MLP_OK;
    }
    const size_t smem = posterior_smem_bytes(kernel, a.Cmax, warps_per_cta);
    int bps = posterior_max_blocks_per_sm(kernel, smem);
    bps = std::min(bps, 16);
    if (bps_cap > 0) bps = std::min(bps, bps_cap);
    int grid = ctx->num_sms * bps;
    grid = std::min(grid, (ntasks + warps_per_cta - 1) / warps_per_cta);
    grid = std::max(grid, 1);
    CK(cudaMemsetAsync(a.counter, 0, sizeof(int), st));
    kt.begin(stat_slot, st);
    CK(posterior_launch(kernel, a, grid, smem, st));
    kt.end(st);
    ctx->stats.launches += 1;
    return MLP_OK;
}

__global__ void k_set_digest(const PairTask* __restrict__ tasks, int ntasks, int n, const int* __restrict__ len, const long long* __restrict__ rp_off,
                             const int* __restrict__ rp_pool, const long long* __restrict__ nz_off, const int* __restrict__ nz_cnt,
                             const int2* __restrict__ cells, unsigned long long* __restrict__ out);

// Streamed posterior stage: what ConsistencyStage::doRelaxation (ConsistencyStage.cpp:133-266) does to a matrix whose pair accepts no
// third sequence -- P / sumW with sumW = 1, cutoff, store as uint16 fixed point -- i.e. one more quantisation of every cell per
// repetition.  A cell below the cutoff would have to be dropped (and the row pointers rebuilt): flagged instead, it cannot happen
// with the reference's cutoffs (every kept posterior is >= 0.01 quantised down, the repetition's cutoff is <= 0.01 resp. 1e-5).
__global__ void k_stream_requant(int2* __restrict__ cells, long long count, int reps, float cutoff, int* __restrict__ err) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (long long)gridDim.x * blockDim.x) {
        float v = __int_as_float(cells[k].y);
        for (int r = 0; r < reps; ++r) {
            v = __fdiv_rn(v, 1.0f);
            if (!(v >= cutoff)) atomicOr(err, 32);
            v = dev_quantize_u16(v);
        }
        cells[k].y = __float_as_int(v);
    }
}

static int run_posterior_tasks(mlp_ctx* ctx, int flavour, uint32_t mask, float cutoff, const std::vector<PairTask>& tasks_in,
                               float* dense, float* dense5, float* denseP, float* denseL) {
    if (tasks_in.empty()) return MLP_OK;
    const bool useP = (mask & MLP_M_PART) != 0, use5 = (mask & MLP_M_HMM5) != 0, useL = (mask & MLP_M_LOCAL) != 0;
    // bytes per dense element: Z(8)+P(4) | S5(4) | SL(4)+VB(4, aliases Z when the partition model is also run)
    int bpe = 0;
    if (useP) bpe += 12;
    if (use5) bpe += 4;
    if (useL) bpe += useP ? 4 : 12;   // SL | Z terms + candidate lists of the Z chain (loc_c.cu), which alias Z when the partition model is also run
    // Memory plan: the first batch is small and measures the sparse density; the cell pool is then grown once to the
    // extrapolated size, the same amount is left free for the relaxation output set, and the rest goes to dense scratch.
    size_t budget = (size_t)1 << 31;
    bool planned = tasks_in.size() < 2048 || ctx->stream_mode;   // streamed stage: the pool holds one batch at a time, nothing to extrapolate
    if (planned) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        budget = ctx->scratch_budget > 0 ? (size_t)ctx->scratch_budget : (size_t)((free_b + ctx->scratch_bytes) * 0.5);
    }
    int maxL1 = 0, maxL2 = 0; bool need_edge = false; long long max_elems = 0;
    for (const PairTask& t : tasks_in) {
        maxL1 = std::max(maxL1, t.L1); maxL2 = std::max(maxL2, t.L2);
        need_edge |= (t.nb > 1);
        max_elems = std::max(max_elems, (long long)t.nb * (t.L1 + 32) * t.C * 32);
    }
    if ((size_t)max_elems * bpe > budget) budget = (size_t)max_elems * bpe;   // a single pair must fit
    // resident warps of the largest grid launch_one can choose (num_sms * 16 CTAs), never more than one warp per task
    const long long max_warps = std::min<long long>((long long)ctx->num_sms * 16 * (MLP_BLOCK / 32),
                                                    (((long long)tasks_in.size() + MLP_BLOCK / 32 - 1) / (MLP_BLOCK / 32)) * (MLP_BLOCK / 32));
    int stage_mult = 32;
    {
        int rc = ensure_warp_buffers(ctx, max_warps, maxL1, maxL2, need_edge, stage_mult);
        if (rc != MLP_OK) return rc;
    }
    KernelTimer kt;
    size_t pos = 0;
    std::vector<PairTask> batch;
    while (pos < tasks_in.size()) {
        batch.clear();
        long long elems = 0; int Cmax = 1;
        while (pos < tasks_in.size()) {
            PairTask t = tasks_in[pos];
            const long long e = (long long)t.nb * (t.L1 + 32) * t.C * 32;
            if (!batch.empty() && (size_t)(elems + e) * bpe > budget) break;
            t.off = elems; elems += e; Cmax = std::max(Cmax, t.C);
            batch.push_back(t); ++pos;
        }
        // the register-band kernels are compiled per columns-per-lane value: order the batch by C (stable, so the cost order
        // survives inside a group) and remember where every group starts
        if (!std::is_sorted(batch.begin(), batch.end(), [](const PairTask& x, const PairTask& y) { return x.C > y.C; }))   // the pair list already is
            std::stable_sort(batch.begin(), batch.end(), [](const PairTask& x, const PairTask& y) { return x.C > y.C; });
        std::vector<std::pair<int, int>> cgroups;   // (C, first task)
        {
            long long o = 0;
            for (size_t k = 0; k < batch.size(); ++k) {
                batch[k].off = o; o += (long long)batch[k].nb * (batch[k].L1 + 32) * batch[k].C * 32;
                if (cgroups.empty() || cgroups.back().first != batch[k].C) cgroups.push_back({batch[k].C, (int)k});
            }
        }
        const size_t need = (size_t)elems * bpe + 256;
        if (need > ctx->scratch_bytes) {
            free_dev(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
            CK(cudaMalloc(&ctx->d_scratch, need));
            ctx->scratch_bytes = need;
        }
        { int rc = ensure_tasks(ctx, batch.size()); if (rc != MLP_OK) return rc; }
        int maxnb = 1; for (const PairTask& t : batch) maxnb = std::max(maxnb, t.nb);
        const long long rowexp_stride = (long long)maxnb * (maxL1 + 4);   // [column block][row] scale exponents
        if (useP && flavour != MLP_QP && batch.size() * (size_t)rowexp_stride > ctx->rowexp_cap) {
            free_dev(ctx->d_rowexp); ctx->d_rowexp = nullptr;
            ctx->rowexp_cap = batch.size() * (size_t)rowexp_stride + 1024;
            CK(cudaMalloc(&ctx->d_rowexp, ctx->rowexp_cap * sizeof(int)));
        }
        const long long rowaux_stride = 3LL * (maxL1 + 2);
        if (useL && !ctx->loc_old && batch.size() * (size_t)rowaux_stride > ctx->rowaux_cap) {
            free_dev(ctx->d_rowaux); ctx->d_rowaux = nullptr;
            ctx->rowaux_cap = batch.size() * (size_t)rowaux_stride + 1024;
            CK(cudaMalloc(&ctx->d_rowaux, ctx->rowaux_cap * sizeof(float)));
        }
        CK(cudaMemcpyAsync(ctx->d_tasks, batch.data(), batch.size() * sizeof(PairTask), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (int64_t)(batch.size() * sizeof(PairTask));
        CK(cudaMemsetAsync(ctx->d_pout, 0, batch.size() * sizeof(PairOut), ctx->stream));

        unsigned long long cursor_before = 0;
        { int rcc = read_cursor(ctx, ctx->cur, &cursor_before); if (rcc != MLP_OK) return rcc; }
        for (int attempt = 0;; ++attempt) {
        KArgs a = {};
        a.tasks = ctx->d_tasks; a.ntasks = (int)batch.size(); a.counter = ctx->d_counter; a.pout = ctx->d_pout;
        a.residues = ctx->d_res; a.seq_off = ctx->d_seq_off; a.n = ctx->n;
        a.flavour = flavour; a.mask = mask; a.cutoff = cutoff; a.Cmax = Cmax;
        a.match = ctx->d_match; a.ins = ctx->d_ins; a.sub = ctx->d_sub;
        unsigned char* p = (unsigned char*)ctx->d_scratch;
        if (useP) { a.layerZ = (double*)p; p += (size_t)elems * 8; a.layerP = (float*)p; p += (size_t)elems * 4; }
        if (use5) { a.layerS5 = (float*)p; p += (size_t)elems * 4; }
        if (useL) {
            a.layerSL = (float*)p; p += (size_t)elems * 4;
            float* lc;
            if (useP) { a.layerVB = (float*)a.layerZ; lc = (float*)a.layerZ + elems; }
            else { a.layerVB = (float*)p; p += (size_t)elems * 4; lc = (float*)p; p += (size_t)elems * 4; }
            a.layerLC = ctx->loc_old ? nullptr : lc;
        }
        a.layerTB = (int*)(a.layerS5 ? a.layerS5 : (a.layerP ? a.layerP : a.layerSL));
        a.rowexp = ctx->d_rowexp; a.rowexp_stride = rowexp_stride;
        a.rowaux = ctx->d_rowaux; a.rowaux_stride = rowaux_stride;
        a.edge_f = need_edge ? (float*)ctx->d_edge : nullptr;
        a.edge_d = need_edge ? (double*)ctx->d_edge : nullptr;
        a.edge_stride = ctx->edge_stride;   // in elements of the kernel's own type; the buffer is sized for doubles
        a.rp_off = ctx->d_rp_off; a.out = ctx->set[ctx->cur]; a.in = CsrSetDev{};
        a.stage = ctx->d_stage; a.stage_cap = ctx->stage_cap;
        a.tfill = ctx->d_tfill; a.tfill_stride = ctx->tfill_stride;
        a.dist = ctx->d_dist; a.err = ctx->d_err;
        a.dense = dense; a.dense5 = dense5; a.denseP = denseP; a.denseL = denseL;

        const int nt = (int)batch.size();
        int rc;
        // The FP64 partition sweeps and the FP32 HMM sweeps are independent until the merge: run them on two streams
        // with capped residency so both kinds of warps share every SM (FP64 pipe + FP32/ALU pipes busy together).
        const bool fork = ctx->overlap && useP && (use5 || useL);
        cudaStream_t sp = fork ? ctx->stream2 : ctx->stream;
        if (fork) { CK(cudaEventRecord(ctx->ev_fork, ctx->stream)); CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0)); }
        const int capP = fork ? (ctx->bps_part > 0 ? ctx->bps_part : 2) : 0;
        const int capH = fork ? (ctx->bps_hmm > 0 ? ctx->bps_hmm : 4) : 0;
        if (useP) {
            KArgs ap = a;
            if (fork && need_edge) ap.edge_d = (double*)ctx->d_edge + (size_t)ctx->edge_warps * ctx->edge_stride;   // own hand-off buffer: the HMM sweeps run at the same time
            if ((rc = launch_one(ctx, MLP_K_PART_FWD, ap, nt, kt, MLP_K_PART_FWD, sp, capP, &cgroups)) != MLP_OK) return rc;
            if ((rc = launch_one(ctx, MLP_K_PART_REV, ap, nt, kt, MLP_K_PART_REV, sp, capP, &cgroups)) != MLP_OK) return rc;
        }
        if (useL) {
            // the local model's Z terms alias the partition layer: it must wait for the partition posterior
            if (fork) { CK(cudaEventRecord(ctx->ev_join, ctx->stream2)); CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); }
            if (posterior_c_available(MLP_K_LOCAL_FWD, a)) {
                // register-band sweeps; the sequential Z chain runs over row-major candidate lists, one thread per pair (loc_c.cu)
                static const bool split = getenv("MLP_LOC_SPLIT") != nullptr;   // developer knob (local model alone): candidate pass -> slots 0 / 2, chain -> slots 1 / 3
                for (int phase = 0; phase < 2; ++phase) {
                    const int kid = phase == 0 ? MLP_K_LOCAL_FWD : MLP_K_LOCAL_BWD;
                    const bool force_guard = getenv("MLP_LOC_FORCE_FALLBACK") != nullptr;   // test knob: the chain reports a failed bound check -> the batch is redone by the round-1 kernels
                    KArgs al = a; al.loc_phase = phase; al.loc_debug = (split ? 1 : 0) | (force_guard ? 2 : 0);
                    if ((rc = launch_one(ctx, kid, al, nt, kt, kid, nullptr, 0, &cgroups)) != MLP_OK) return rc;
                    if ((rc = launch_one(ctx, MLP_K_LOCAL_CAND, al, nt, kt, split ? 2 * phase : kid, nullptr, 0, &cgroups)) != MLP_OK) return rc;
                    kt.begin(split ? 2 * phase + 1 : kid, ctx->stream);
                    CK(loc_replay_launch(al, ctx->stream));
                    kt.end(ctx->stream);
                    ctx->stats.launches += 1;
                }
            } else {
                if ((rc = launch_one(ctx, MLP_K_LOCAL_FWD, a, nt, kt, MLP_K_LOCAL_FWD)) != MLP_OK) return rc;
                if ((rc = launch_one(ctx, MLP_K_LOCAL_BWD, a, nt, kt, MLP_K_LOCAL_BWD)) != MLP_OK) return rc;
            }
        }
        if (use5) {
            if ((rc = launch_one(ctx, MLP_K_HMM_FWD, a, nt, kt, MLP_K_HMM_FWD, nullptr, useL ? 0 : capH, &cgroups)) != MLP_OK) return rc;
            if ((rc = launch_one(ctx, MLP_K_HMM_BWD, a, nt, kt, MLP_K_HMM_BWD, nullptr, useL ? 0 : capH, &cgroups)) != MLP_OK) return rc;
        }
        if (fork && !useL) { CK(cudaEventRecord(ctx->ev_join, ctx->stream2)); CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); }
        if ((rc = launch_one(ctx, MLP_K_FINAL, a, nt, kt, MLP_K_FINAL, nullptr, 0, &cgroups)) != MLP_OK) return rc;
        if ((rc = launch_one(ctx, MLP_K_TRANSPOSE, a, nt, kt, MLP_K_FINAL)) != MLP_OK) return rc;
        int err = 0;
        { int rcw = read_words(ctx, ctx->cur, nullptr, &err); if (rcw != MLP_OK) return rcw; }   // syncs the stream
        if (!err) break;
        // capacity miss: grow what overflowed and redo this batch (batches are idempotent once the cursor is rewound)
        cudaMemset(ctx->d_err, 0, sizeof(int));
        if (err & 8) { ctx->err = "transpose met a column index outside its matrix (corrupt cell pool)"; return MLP_E_CUDA; }
        if (err & 16) ctx->loc_old = true;   // the filtered Z chain of the local model met a running sum below its checked bound: redo the batch with the unfiltered kernels
        if (attempt >= 6) { ctx->err = "sparse capacity still exhausted after 6 growth attempts"; return MLP_E_CAPACITY; }
        if (err & 1) {
            stage_mult *= 4;
            int rc2 = ensure_warp_buffers(ctx, max_warps, maxL1, maxL2, need_edge, stage_mult);
            if (rc2 != MLP_OK) return rc2;
        }
        if (err & 2) {
            unsigned long long used = 0;
            { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
            const double done_frac = (double)(pos) / (double)tasks_in.size();
            long long want = (long long)((double)used / std::max(done_frac, 1e-3) * 1.15) + (1 << 20);
            want = std::max(want, ctx->set[ctx->cur].cap * 2);
            if (ctx->stream_mode) want = ctx->set[ctx->cur].cap * 2;
            int rc2 = grow_cells(ctx, ctx->cur, want, cursor_before);
            if (rc2 != MLP_OK) return rc2;
        }
        CK(cudaMemcpyAsync(ctx->set[ctx->cur].cursor, &cursor_before, sizeof(cursor_before), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        }
        if (ctx->stream_mode) {
            // finish, digest and drop this batch's matrices: the next batch reuses the pool from the same cursor
            unsigned long long used = 0;
            { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
            const long long cnt = (long long)used - (long long)cursor_before;
            const CsrSetDev& S = ctx->set[ctx->cur];
            if (cnt > 0) {
                k_stream_requant<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(S.cells + cursor_before, cnt, ctx->stream_reps, 1e-5f, ctx->d_err);
                CK(cudaGetLastError());
            }
            k_set_digest<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->d_tasks, (int)batch.size(), ctx->n, ctx->d_len, ctx->d_rp_off, S.rp_pool, S.nz_off, S.nz_cnt, S.cells, ctx->d_sdigest);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(S.cursor, &cursor_before, sizeof(cursor_before), cudaMemcpyHostToDevice, ctx->stream));
            int err2 = 0;
            { int rcw = read_words(ctx, ctx->cur, nullptr, &err2); if (rcw != MLP_OK) return rcw; }
            if (err2) { ctx->err = "streamed stage: a cell fell below the consistency cutoff (not supported in streamed mode)"; return MLP_E_UNSUPPORTED; }
            ctx->stats.launches += 2;
            ctx->stats.nnz += cnt / 2;
        }
        if (!planned && pos < tasks_in.size()) {
            unsigned long long used = 0;
            { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
            double done_cost = 0, all_cost = 0;
            for (size_t k = 0; k < tasks_in.size(); ++k) { const double c = (double)std::min(tasks_in[k].L1, tasks_in[k].L2); all_cost += c; if (k < pos) done_cost += c; }   // kept cells scale with the shorter length
            const long long want = (long long)((double)(used - cursor_before) * (all_cost / std::max(done_cost, 1.0)) * 1.25) + (long long)cursor_before + (1 << 20);
            if (want > ctx->set[ctx->cur].cap) { int rc2 = grow_cells(ctx, ctx->cur, want, used); if (rc2 != MLP_OK) return rc2; }
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            const size_t reserve = (size_t)want * sizeof(int2) + ((size_t)2 << 30);   // relaxation output set + slack
            const size_t avail = free_b + ctx->scratch_bytes;
            size_t b2 = avail > reserve ? (size_t)((avail - reserve) * 0.9) : ((size_t)1 << 30);
            if (ctx->scratch_budget > 0) b2 = std::min(b2, (size_t)ctx->scratch_budget);
            budget = std::max(b2, (size_t)max_elems * bpe);
            planned = true;
        }
        if (useP && flavour != MLP_QP) {
            // cpnp runs the partition function in 80-bit long double; this FP64 kernel cannot represent Z beyond 1e308
            std::vector<PairOut> po(batch.size());
            CK(cudaMemcpy(po.data(), ctx->d_pout, batch.size() * sizeof(PairOut), cudaMemcpyDeviceToHost));
            for (const PairOut& o : po)
                if (!std::isfinite(o.Zpart) || o.Zpart == 0.0) {
                    ctx->err = "partition function left the FP64 range for at least one pair (cpnp uses long double)";
                    return MLP_E_OVERFLOW;
                }
        }
        for (const PairTask& t : batch) { ctx->stats.cells += (int64_t)(t.L1 + 1) * (t.L2 + 1); }
        ctx->stats.pairs += (int64_t)batch.size();
    }
    kt.collect(ctx->stats);
    return MLP_OK;
}

// developer hook (tools/loc_ab.py with MLP_LOC_SPLIT=1): candidates / firing cells of the local model's forward and backward Z chains since the last call
extern "C" int mlp_debug_loc_counters(mlp_ctx* ctx, unsigned long long* out4) {
    if (!ctx || !out4) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(loc_debug_counters(out4));
    return MLP_OK;
}

extern "C" int mlp_viterbi_all_pairs(mlp_ctx* ctx, int32_t* n_identical, int32_t* align_len) {
    return mlp_viterbi_all_pairs_ex(ctx, n_identical, align_len, nullptr, nullptr);
}

extern "C" int mlp_viterbi_all_pairs_ex(mlp_ctx* ctx, int32_t* n_identical, int32_t* align_len, char* aln, int64_t* aln_off) {
    if (!ctx || !n_identical || !align_len || ((aln == nullptr) != (aln_off == nullptr))) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    if (!ctx->have_tables || ctx->n < 2) { ctx->err = "set tables and sequences first"; return MLP_E_STATE; }
    const std::vector<PairTask>& tasks_in = ctx->owned;
    const size_t npairs_all = ctx->all_pairs.size();
    int* d_ident = nullptr; int* d_len = nullptr;
    CK(cudaMalloc(&d_ident, npairs_all * sizeof(int)));
    CK(cudaMalloc(&d_len, npairs_all * sizeof(int)));
    CK(cudaMemset(d_ident, 0, npairs_all * sizeof(int)));
    CK(cudaMemset(d_len, 0, npairs_all * sizeof(int)));
    char* d_aln = nullptr; long long* d_aln_off = nullptr;
    std::vector<long long> off_h;
    if (aln) {   // strings by pair index (row-major a<b), each with room for len[a]+len[b] columns
        off_h.assign(npairs_all + 1, 0);
        size_t p = 0;
        for (int a = 0; a < ctx->n; ++a) for (int b = a + 1; b < ctx->n; ++b, ++p) off_h[p + 1] = off_h[p] + ctx->len[a] + ctx->len[b];
        CK(cudaMalloc(&d_aln, (size_t)off_h[npairs_all] + 16));
        CK(cudaMalloc(&d_aln_off, (npairs_all + 1) * sizeof(long long)));
        CK(cudaMemcpy(d_aln_off, off_h.data(), (npairs_all + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    }
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = ctx->scratch_budget > 0 ? (size_t)ctx->scratch_budget : (size_t)((free_b + ctx->scratch_bytes) * 0.5);
    int maxL1 = 0, maxL2 = 0; bool need_edge = false;
    for (const PairTask& t : tasks_in) { maxL1 = std::max(maxL1, t.L1); maxL2 = std::max(maxL2, t.L2); need_edge |= (t.nb > 1); }
    const long long max_warps = (long long)ctx->num_sms * 16 * (MLP_BLOCK / 32);
    int rc = ensure_warp_buffers(ctx, max_warps, maxL1, maxL2, need_edge, 1);
    if (rc != MLP_OK) return rc;
    ctx->stats = mlp_stage_stats{};
    KernelTimer kt;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    size_t pos = 0;
    std::vector<PairTask> batch;
    while (pos < tasks_in.size()) {
        batch.clear();
        long long elems = 0; int Cmax = 1;
        while (pos < tasks_in.size()) {
            PairTask t = tasks_in[pos];
            const long long e = (long long)t.nb * (t.L1 + 32) * t.C * 32;
            if (!batch.empty() && (size_t)(elems + e) > budget) break;
            t.off = elems; elems += e; Cmax = std::max(Cmax, t.C);
            batch.push_back(t); ++pos;
        }
        const size_t need = (size_t)elems + 256;   // one traceback byte per cell
        if (need > ctx->scratch_bytes) {
            free_dev(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
            CK(cudaMalloc(&ctx->d_scratch, need));
            ctx->scratch_bytes = need;
        }
        rc = ensure_tasks(ctx, batch.size());
        if (rc != MLP_OK) return rc;
        CK(cudaMemcpyAsync(ctx->d_tasks, batch.data(), batch.size() * sizeof(PairTask), cudaMemcpyHostToDevice, ctx->stream));
        KArgs a = {};
        a.tasks = ctx->d_tasks; a.ntasks = (int)batch.size(); a.pout = ctx->d_pout;
        a.residues = ctx->d_res; a.seq_off = ctx->d_seq_off; a.n = ctx->n; a.Cmax = Cmax;
        a.match = ctx->d_match; a.ins = ctx->d_ins; a.sub = ctx->d_sub;
        a.layerTB8 = (unsigned char*)ctx->d_scratch; a.vit_ident = d_ident; a.vit_len = d_len;
        a.vit_init0 = logf((float)0.6080327034); a.vit_init1 = logf((float)0.1959836632);   // ProbabilisticModel.h:1070-1072
        a.vit_aln = d_aln; a.vit_aln_off = d_aln_off;
        a.edge_f = need_edge ? (float*)ctx->d_edge : nullptr; a.edge_stride = ctx->edge_stride; a.err = ctx->d_err;
        if ((rc = launch_one(ctx, MLP_K_VITERBI, a, (int)batch.size(), kt, MLP_K_LOCAL_FWD)) != MLP_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        for (const PairTask& t : batch) ctx->stats.cells += (int64_t)(t.L1 + 1) * (t.L2 + 1);
        ctx->stats.pairs += (int64_t)batch.size();
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_total = ms;
    kt.collect(ctx->stats);
    CK(cudaMemcpy(n_identical, d_ident, npairs_all * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(align_len, d_len, npairs_all * sizeof(int), cudaMemcpyDeviceToHost));
    ctx->stats.d2h_bytes += (int64_t)npairs_all * 8;
    if (aln) {   // compact the per-pair strings and turn them front to back (the traceback wrote them reversed)
        std::vector<char> raw((size_t)off_h[npairs_all] + 16);
        CK(cudaMemcpy(raw.data(), d_aln, (size_t)off_h[npairs_all], cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += off_h[npairs_all];
        int64_t w = 0;
        for (size_t p = 0; p < npairs_all; ++p) {
            aln_off[p] = w;
            const char* src = raw.data() + off_h[p];
            for (int k = align_len[p] - 1; k >= 0; --k) aln[w++] = src[k];
        }
        aln_off[npairs_all] = w;
        cudaFree(d_aln); cudaFree(d_aln_off);
    }
    cudaFree(d_ident); cudaFree(d_len);
    return MLP_OK;
}

extern "C" int mlp_posterior_all_pairs(mlp_ctx* ctx, int flavour, uint32_t model_mask, float cutoff) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }
    if (ctx->rb_set == 0) END_READBACK(ctx);         // this stage writes set 0; a read-back of set 1 (a relaxed set) may go on beside it
    cudaSetDevice(ctx->device);
    if (!ctx->have_tables || ctx->n < 2) { ctx->err = "set tables and sequences first"; return MLP_E_STATE; }
    if (flavour < MLP_QP || flavour > MLP_CPNP_P1 || (model_mask & 7u) == 0) return MLP_E_ARG;
    if (flavour == MLP_QP) model_mask = MLP_M_HMM5 | MLP_M_PART;
    if (ctx->stream_mode && flavour != MLP_QP) { ctx->err = "the streamed stage finishes matrices the way QuickProbs' consistency does; c_p_np_aln reads every matrix"; return MLP_E_UNSUPPORTED; }
    if (!ctx->restricted) ctx->tree_resident = false;   // a device guide tree belongs to the distance matrix it was built from (a stage over a restricted shard recomputes part of the same matrix)
    if ((model_mask & MLP_M_PART) && flavour != MLP_QP) {
        // letters J, O, U index sub_matrix[-1] in the reference: refuse instead of guessing
        for (long long k = 0; k < ctx->total_res; ++k) {
            const int c = ctx->codes_h[k];
            if (std::isnan(ctx->part.sub[c][c])) { ctx->err = "sequence contains a letter the reference partition table cannot score (J/O/U)"; return MLP_E_UNSUPPORTED; }
        }
    }
    int rc = ensure_sets(ctx);
    if (rc != MLP_OK) return rc;
    ctx->stats = mlp_stage_stats{};
    ctx->cur = 0;
    CK(cudaMemsetAsync(ctx->set[0].cursor, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->set[0].nz_cnt, 0, (size_t)ctx->n * ctx->n * sizeof(int), ctx->stream));
    if (ctx->world > 1) {   // slots of pairs other shards own must read as zero for the sum-exchange (mlp_exchange)
        CK(cudaMemsetAsync(ctx->set[0].rp_pool, 0, (size_t)ctx->rp_total * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->set[0].nz_off, 0, (size_t)ctx->n * ctx->n * sizeof(long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_dist, 0, (size_t)ctx->n * ctx->n * sizeof(float), ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = run_posterior_tasks(ctx, flavour, model_mask, cutoff, ctx->owned, nullptr, nullptr, nullptr, nullptr);
    if (rc != MLP_OK) return rc;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_total = ms;
    unsigned long long cur = 0;
    { int rcc = read_cursor(ctx, 0, &cur); if (rcc != MLP_OK) return rcc; }
    if (!ctx->stream_mode) ctx->stats.nnz = (int64_t)(cur / 2);
    ctx->flavour_of_set = ctx->stream_mode ? -1 : flavour;   // a streamed stage leaves distances and digests, not a set
    ctx->set_partial = ctx->dist_partial = (ctx->world > 1);
    ctx->imported = false;
    return MLP_OK;
}

extern "C" int mlp_debug_pair_dense(mlp_ctx* ctx, int flavour, uint32_t model_mask, int a, int b,
                                    float* merged, float* p_hmm5, float* p_part, float* p_local, float* distance) {
    if (!ctx || !merged) return MLP_E_ARG;
    END_READBACK(ctx);
    cudaSetDevice(ctx->device);
    if (!ctx->have_tables || ctx->n < 2) return MLP_E_STATE;
    if (a < 0 || b <= a || b >= ctx->n) return MLP_E_ARG;
    if (flavour == MLP_QP) model_mask = MLP_M_HMM5 | MLP_M_PART;
    int rc = ensure_sets(ctx);
    if (rc != MLP_OK) return rc;
    const int L1 = ctx->len[a], L2 = ctx->len[b];
    const size_t cells = (size_t)(L1 + 1) * (L2 + 1);
    float* d = nullptr;
    CK(cudaMalloc(&d, cells * 4 * sizeof(float)));
    CK(cudaMemset(d, 0, cells * 4 * sizeof(float)));
    std::vector<PairTask> one;
    for (const PairTask& t : ctx->all_pairs) if (t.a == a && t.b == b) one.push_back(t);
    // The run goes through the live set: save what it overwrites (cursor, the pair's table entries, row pointers and distance)
    // and put it back afterwards, so that a relaxed set survives a debug call.
    unsigned long long cursor_save = 0;
    { int rcc = read_cursor(ctx, ctx->cur, &cursor_save); if (rcc != MLP_OK) return rcc; }
    const int n = ctx->n;
    const size_t sAB = (size_t)a * n + b, sBA = (size_t)b * n + a;
    CsrSetDev& S = ctx->set[ctx->cur];
    long long off_save[2]; int cnt_save[2]; float dist_save[2];
    std::vector<int> rpA(L1 + 2), rpB(L2 + 2);
    CK(cudaMemcpy(&off_save[0], S.nz_off + sAB, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&off_save[1], S.nz_off + sBA, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cnt_save[0], S.nz_cnt + sAB, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cnt_save[1], S.nz_cnt + sBA, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&dist_save[0], ctx->d_dist + sAB, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&dist_save[1], ctx->d_dist + sBA, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rpA.data(), S.rp_pool + ctx->rp_off_h[sAB], rpA.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rpB.data(), S.rp_pool + ctx->rp_off_h[sBA], rpB.size() * 4, cudaMemcpyDeviceToHost));
    mlp_stage_stats keep = ctx->stats;
    rc = run_posterior_tasks(ctx, flavour == MLP_CPNP_P1 ? MLP_CPNP_P1 : flavour, model_mask, 0.01f, one, d, d + cells, d + 2 * cells, d + 3 * cells);
    ctx->stats = keep;
    float dist_new = 0.0f;
    if (rc == MLP_OK) cudaMemcpy(&dist_new, ctx->d_dist + sAB, sizeof(float), cudaMemcpyDeviceToHost);
    {
        CsrSetDev& R = ctx->set[ctx->cur];   // the cell pool may have been re-allocated; the tables are the same arrays
        cudaMemcpy(R.cursor, &cursor_save, 8, cudaMemcpyHostToDevice);
        cudaMemcpy(R.nz_off + sAB, &off_save[0], 8, cudaMemcpyHostToDevice); cudaMemcpy(R.nz_off + sBA, &off_save[1], 8, cudaMemcpyHostToDevice);
        cudaMemcpy(R.nz_cnt + sAB, &cnt_save[0], 4, cudaMemcpyHostToDevice); cudaMemcpy(R.nz_cnt + sBA, &cnt_save[1], 4, cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_dist + sAB, &dist_save[0], 4, cudaMemcpyHostToDevice); cudaMemcpy(ctx->d_dist + sBA, &dist_save[1], 4, cudaMemcpyHostToDevice);
        cudaMemcpy(R.rp_pool + ctx->rp_off_h[sAB], rpA.data(), rpA.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(R.rp_pool + ctx->rp_off_h[sBA], rpB.data(), rpB.size() * 4, cudaMemcpyHostToDevice);
    }
    if (rc == MLP_OK) {
        cudaMemcpy(merged, d, cells * sizeof(float), cudaMemcpyDeviceToHost);
        if (p_hmm5) cudaMemcpy(p_hmm5, d + cells, cells * sizeof(float), cudaMemcpyDeviceToHost);
        if (p_part) cudaMemcpy(p_part, d + 2 * cells, cells * sizeof(float), cudaMemcpyDeviceToHost);
        if (p_local) cudaMemcpy(p_local, d + 3 * cells, cells * sizeof(float), cudaMemcpyDeviceToHost);
        if (distance) *distance = dist_new;
    }
    cudaFree(d);
    return rc;
}

extern "C" int mlp_get_distances(mlp_ctx* ctx, float* nxn) {
    if (!ctx || !nxn) return MLP_E_ARG;
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    if (ctx->exch_pending) {   // split exchange in flight: the distances are ready once their all-reduce has run
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_dist, 0));
        CK(cudaMemcpyAsync(nxn, ctx->d_dist, (size_t)ctx->n * ctx->n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream2));
        CK(cudaStreamSynchronize(ctx->stream2));
        ctx->stats.d2h_bytes += (int64_t)ctx->n * ctx->n * 4;
        return MLP_OK;
    }
    CK(cudaMemcpy(nxn, ctx->d_dist, (size_t)ctx->n * ctx->n * sizeof(float), cudaMemcpyDeviceToHost));
    ctx->stats.d2h_bytes += (int64_t)ctx->n * ctx->n * 4;
    return MLP_OK;
}

extern "C" int mlp_get_csr(mlp_ctx* ctx, int a, int b, int32_t* row_ptr, int32_t* col, float* val, int64_t* nnz) {
    if (!ctx || a < 0 || b < 0 || a >= ctx->n || b >= ctx->n || a == b) return MLP_E_ARG;
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    const CsrSetDev& s = ctx->set[ctx->cur];
    const size_t slot = (size_t)a * ctx->n + b;
    int cnt = 0; long long off = 0;
    CK(cudaMemcpy(&cnt, s.nz_cnt + slot, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&off, s.nz_off + slot, sizeof(long long), cudaMemcpyDeviceToHost));
    if (nnz) *nnz = cnt;
    if (row_ptr) CK(cudaMemcpy(row_ptr, s.rp_pool + ctx->rp_off_h[slot], (size_t)(ctx->len[a] + 2) * sizeof(int), cudaMemcpyDeviceToHost));
    if ((col || val) && cnt > 0) {
        std::vector<int2> cells(cnt);
        CK(cudaMemcpy(cells.data(), s.cells + off, (size_t)cnt * sizeof(int2), cudaMemcpyDeviceToHost));
        for (int k = 0; k < cnt; ++k) {
            if (col) col[k] = cells[k].x;
            if (val) std::memcpy(&val[k], &cells[k].y, 4);
        }
    }
    return MLP_OK;
}

extern "C" int mlp_total_cells(mlp_ctx* ctx, int64_t* cells) {
    if (!ctx || !cells) return MLP_E_ARG;
    END_EXCHANGE(ctx);
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    std::vector<int> cnt((size_t)ctx->n * ctx->n);
    CK(cudaMemcpy(cnt.data(), ctx->set[ctx->cur].nz_cnt, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
    int64_t s = 0;
    for (int c : cnt) s += c;
    *cells = s;
    return MLP_OK;
}

extern "C" int mlp_get_csr_bulk(mlp_ctx* ctx, int64_t* nnz_per_pair, int32_t* row_ptr, int32_t* col, float* val) {
    if (!ctx) return MLP_E_ARG;
    END_EXCHANGE(ctx);
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    const int n = ctx->n;
    const CsrSetDev& s = ctx->set[ctx->cur];
    std::vector<int> cnt((size_t)n * n);
    std::vector<long long> off((size_t)n * n);
    CK(cudaMemcpy(cnt.data(), s.nz_cnt, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(off.data(), s.nz_off, off.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    ctx->stats.d2h_bytes += (int64_t)n * n * 12;
    std::vector<int> rp_all;
    if (row_ptr) {
        rp_all.resize(ctx->rp_total);
        CK(cudaMemcpy(rp_all.data(), s.rp_pool, (size_t)ctx->rp_total * sizeof(int), cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += ctx->rp_total * 4;
    }
    std::vector<int2> cells;
    unsigned long long used = 0;
    if (col || val) {
        { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
        used = std::min<unsigned long long>(used, (unsigned long long)s.cap);
        cells.resize(used);
        if (used) CK(cudaMemcpy(cells.data(), s.cells, used * sizeof(int2), cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += (int64_t)used * 8;
    }
    int64_t p = 0, rpos = 0, cpos = 0;
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b, ++p) {
            const size_t slot = (size_t)a * n + b;
            if (nnz_per_pair) nnz_per_pair[p] = cnt[slot];
            if (row_ptr) { std::memcpy(row_ptr + rpos, rp_all.data() + ctx->rp_off_h[slot], (size_t)(ctx->len[a] + 2) * sizeof(int)); rpos += ctx->len[a] + 2; }
            if (col || val)
                for (int k = 0; k < cnt[slot]; ++k) {
                    const int2 c = cells[off[slot] + k];
                    if (col) col[cpos] = c.x;
                    if (val) std::memcpy(&val[cpos], &c.y, 4);
                    ++cpos;
                }
        }
    return MLP_OK;
}

extern "C" int mlp_csr_layout(mlp_ctx* ctx, int64_t* rp_off, int64_t* rp_total, int64_t* cells_used) {
    if (!ctx) return MLP_E_ARG;
    END_EXCHANGE(ctx);
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    if (rp_off) for (size_t k = 0; k < ctx->rp_off_h.size(); ++k) rp_off[k] = ctx->rp_off_h[k];
    if (rp_total) *rp_total = ctx->rp_total;
    if (cells_used) {
        unsigned long long used = 0;
        { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
        *cells_used = (int64_t)std::min<unsigned long long>(used, (unsigned long long)ctx->set[ctx->cur].cap);
    }
    return MLP_OK;
}

extern "C" int mlp_get_csr_raw(mlp_ctx* ctx, int64_t* nz_off, int32_t* nz_cnt, int32_t* rp_pool, void* cells) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    const CsrSetDev& s = ctx->set[ctx->cur];
    const size_t nn = (size_t)ctx->n * ctx->n;
    unsigned long long used = 0;
    { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
    used = std::min<unsigned long long>(used, (unsigned long long)s.cap);
    if (nz_off) { CK(cudaMemcpyAsync(nz_off, s.nz_off, nn * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += (int64_t)nn * 8; }
    if (nz_cnt) { CK(cudaMemcpyAsync(nz_cnt, s.nz_cnt, nn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += (int64_t)nn * 4; }
    if (rp_pool) { CK(cudaMemcpyAsync(rp_pool, s.rp_pool, (size_t)ctx->rp_total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += ctx->rp_total * 4; }
    if (cells && used) { CK(cudaMemcpyAsync(cells, s.cells, (size_t)used * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += (int64_t)used * 8; }
    CK(cudaStreamSynchronize(ctx->stream));
    return MLP_OK;
}

// QuickProbs' own cell format (PackedSparseMatrix: uint16 column, uint16 fixed-point value): half the bytes of {int32, float}.
__global__ void k_pack_cells(const int2* __restrict__ cells, unsigned* __restrict__ out, long long ncell) {
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < ncell; i += stride) {
        const int4 two = *reinterpret_cast<const int4*>(cells + i);           // cells + i is 16-byte aligned (i even); the pool has slack
        const unsigned c0 = __float2uint_rn(__fmul_rn(__int_as_float(two.y), 65535.0f)), c1 = __float2uint_rn(__fmul_rn(__int_as_float(two.w), 65535.0f));
        *reinterpret_cast<uint2*>(out + i) = make_uint2(((unsigned)two.x & 0xffffu) | (c0 << 16), ((unsigned)two.z & 0xffffu) | (c1 << 16));   // {uint16 first = column, uint16 second = value}
    }
}
// cumulative row pointers -> uint16 row sizes, same pooled positions (entry r = cells in row r; the last entry of a pair is 0)
__global__ void k_row_sizes(const int* __restrict__ rp, unsigned short* __restrict__ out, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int d = (i + 1 < total) ? rp[i + 1] - rp[i] : 0;
        out[i] = (unsigned short)(d > 0 ? d : 0);                            // negative = boundary to the next pair's block
    }
}

extern "C" int mlp_get_csr_packed(mlp_ctx* ctx, int64_t* nz_off, int32_t* nz_cnt, uint16_t* row_sizes, uint32_t* cells) {
    if (!ctx) return MLP_E_ARG;
    END_READBACK(ctx);
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }
    if (!ctx->have_sets) return MLP_E_STATE;
    if (ctx->flavour_of_set != MLP_QP) { ctx->err = "packed cells are the QuickProbs format (uint16 fixed-point values)"; return MLP_E_UNSUPPORTED; }
    cudaSetDevice(ctx->device);
    const CsrSetDev& s = ctx->set[ctx->cur];
    const int other = 1 - ctx->cur;
    const size_t nn = (size_t)ctx->n * ctx->n;
    unsigned long long used = 0;
    { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
    used = std::min<unsigned long long>(used, (unsigned long long)s.cap);
    // scratch: the cell pool of the set that is not current (dead after a relaxation, unused before one)
    const unsigned long long used_even = (used + 1) & ~1ull;
    const size_t need_bytes = (size_t)used_even * 4 + (size_t)ctx->rp_total * 2 + 64;
    if ((size_t)ctx->set[other].cap * sizeof(int2) < need_bytes) {
        const int rc = grow_cells(ctx, other, (long long)(need_bytes / sizeof(int2) + 16), 0);
        if (rc != MLP_OK) return rc;
    }
    unsigned* d_cells = reinterpret_cast<unsigned*>(ctx->set[other].cells);
    unsigned short* d_sizes = reinterpret_cast<unsigned short*>(d_cells + used_even);
    const int grid = ctx->num_sms * 8;
    if (nz_off) { CK(cudaMemcpyAsync(nz_off, s.nz_off, nn * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += (int64_t)nn * 8; }
    if (nz_cnt) { CK(cudaMemcpyAsync(nz_cnt, s.nz_cnt, nn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)); ctx->stats.d2h_bytes += (int64_t)nn * 4; }
    if (row_sizes) {
        k_row_sizes<<<grid, 256, 0, ctx->stream>>>(s.rp_pool, d_sizes, ctx->rp_total);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(row_sizes, d_sizes, (size_t)ctx->rp_total * 2, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += ctx->rp_total * 2; ctx->stats.launches += 1;
    }
    if (cells && used) {
        k_pack_cells<<<grid, 256, 0, ctx->stream>>>(s.cells, d_cells, (long long)used);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(cells, d_cells, (size_t)used * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (int64_t)used * 4; ctx->stats.launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return MLP_OK;
}

// Per-matrix digest of the current set, computed on the device: one warp per owned pair and orientation, a polynomial hash
// mod 2^64 over the row pointers followed by the (column, value bits) of every cell.  Position-weighted, so it pins order as
// well as content, and independent of where the cells sit in the pool -- the same value on one GPU and on any sharding.
__global__ void k_set_digest(const PairTask* __restrict__ tasks, int ntasks, int n, const int* __restrict__ len, const long long* __restrict__ rp_off,
                             const int* __restrict__ rp_pool, const long long* __restrict__ nz_off, const int* __restrict__ nz_cnt,
                             const int2* __restrict__ cells, unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long B = 0x9E3779B97F4A7C15ull;
    unsigned long long B32 = 1, Bl = 1;
    for (int k = 0; k < 32; ++k) { B32 *= B; if (k < lane) Bl *= B; }
    for (long long w = gw; w < 2LL * ntasks; w += nw) {
        const PairTask t = tasks[w >> 1];
        const int a = (w & 1) ? t.b : t.a, b = (w & 1) ? t.a : t.b;
        const long long slot = (long long)a * n + b;
        const int rows = len[a] + 2, cnt = nz_cnt[slot];
        const int* rp = rp_pool + rp_off[slot];
        const int2* c = cells + nz_off[slot];
        unsigned long long h = 0, pw = Bl;
        const long long words = rows + 2LL * cnt;
        for (long long k = lane; k < words; k += 32) {
            unsigned v;
            if (k < rows) v = (unsigned)rp[k];
            else { const long long q = k - rows; const int2 e = c[q >> 1]; v = (q & 1) ? (unsigned)e.y : (unsigned)e.x; }
            h += ((unsigned long long)v + 1ull) * pw;
            pw *= B32;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
        if (lane == 0) out[slot] = h + (unsigned long long)cnt;
    }
}

extern "C" int mlp_set_digest(mlp_ctx* ctx, uint64_t* per_matrix_nn) {
    if (!ctx || !per_matrix_nn) return MLP_E_ARG;
    END_EXCHANGE(ctx);
    if (!ctx->have_sets) return MLP_E_STATE;
    cudaSetDevice(ctx->device);
    const int n = ctx->n;
    const size_t nn = (size_t)n * n;
    int rc = ensure_tasks(ctx, ctx->owned.size());
    if (rc != MLP_OK) return rc;
    unsigned long long* d_out = nullptr; int* d_len = nullptr;
    CK(cudaMalloc(&d_out, nn * sizeof(unsigned long long)));
    CK(cudaMalloc(&d_len, n * sizeof(int)));
    CK(cudaMemsetAsync(d_out, 0, nn * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemcpyAsync(d_len, ctx->len.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tasks, ctx->owned.data(), ctx->owned.size() * sizeof(PairTask), cudaMemcpyHostToDevice, ctx->stream));
    const CsrSetDev& s = ctx->set[ctx->cur];
    if (!ctx->owned.empty()) {
        k_set_digest<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->d_tasks, (int)ctx->owned.size(), n, d_len, ctx->d_rp_off, s.rp_pool, s.nz_off, s.nz_cnt, s.cells, d_out);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(per_matrix_nn, d_out, nn * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_out); cudaFree(d_len);
    return MLP_OK;
}

// Split read-back: _begin packs the current set into the context's own pack buffer and enqueues the copies on a separate stream,
// _end waits for them.  Between the two the caller may run the NEXT posterior stage (it writes set 0 and the dense scratch,
// the relaxed set being read is set 1): the PCIe transfer of one step's result then hides behind the compute of the next.
// Every call that would overwrite the set being read ends the read-back first.  The host buffers must stay untouched until _end.
// row sizes of the owned matrices only, concatenated in owned-list order ((a,b) then (b,a) of every task)
__global__ void k_row_sizes_owned(const PairTask* __restrict__ tasks, int ntasks, int n, const long long* __restrict__ rp_off, const int* __restrict__ rp_pool,
                                  const long long* __restrict__ rs_off, unsigned short* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = gw; w < 2LL * ntasks; w += nw) {
        const PairTask t = tasks[w >> 1];
        const int a = (w & 1) ? t.b : t.a, b = (w & 1) ? t.a : t.b, rows = ((w & 1) ? t.L2 : t.L1) + 2;
        const int* rp = rp_pool + rp_off[(long long)a * n + b];
        unsigned short* o = out + rs_off[w >> 1] + ((w & 1) ? t.L1 + 2 : 0);
        for (int i = lane; i < rows; i += 32) { const int d = (i + 1 < rows) ? rp[i + 1] - rp[i] : 0; o[i] = (unsigned short)(d > 0 ? d : 0); }
    }
}

extern "C" int mlp_get_csr_packed_begin(mlp_ctx* ctx, int64_t* nz_off, int32_t* nz_cnt, uint16_t* row_sizes, uint32_t* cells) {
    if (!ctx) return MLP_E_ARG;
    END_EXCHANGE(ctx);
    END_READBACK(ctx);
    if (!ctx->have_sets) return MLP_E_STATE;
    if (ctx->flavour_of_set != MLP_QP) { ctx->err = "packed cells are the QuickProbs format (uint16 fixed-point values)"; return MLP_E_UNSUPPORTED; }
    cudaSetDevice(ctx->device);
    const CsrSetDev& s = ctx->set[ctx->cur];
    const size_t nn = (size_t)ctx->n * ctx->n;
    unsigned long long used = 0;
    { int rcc = read_cursor(ctx, ctx->cur, &used); if (rcc != MLP_OK) return rcc; }
    used = std::min<unsigned long long>(used, (unsigned long long)s.cap);
    const unsigned long long used_even = (used + 1) & ~1ull;
    const size_t need_bytes = (size_t)used_even * 4 + (size_t)ctx->rp_total * 2 + 64;
    if (need_bytes > ctx->pack_bytes) {
        free_dev(ctx->d_pack); ctx->d_pack = nullptr; ctx->pack_bytes = 0;
        CK(cudaMalloc(&ctx->d_pack, need_bytes + need_bytes / 16));
        ctx->pack_bytes = need_bytes + need_bytes / 16;
    }
    unsigned* d_cells = reinterpret_cast<unsigned*>(ctx->d_pack);
    unsigned short* d_sizes = reinterpret_cast<unsigned short*>(d_cells + used_even);
    cudaStream_t st = ctx->stream_rb;
    CK(cudaEventRecord(ctx->ev_rb, ctx->stream));            // everything queued on the main stream so far (the stage that produced the set)
    CK(cudaStreamWaitEvent(st, ctx->ev_rb, 0));
    const int grid = ctx->num_sms * 2;                       // a thin grid: it shares the SMs with the next stage
    if (nz_off) { CK(cudaMemcpyAsync(nz_off, s.nz_off, nn * sizeof(long long), cudaMemcpyDeviceToHost, st)); ctx->stats.d2h_bytes += (int64_t)nn * 8; }
    if (nz_cnt) { CK(cudaMemcpyAsync(nz_cnt, s.nz_cnt, nn * sizeof(int), cudaMemcpyDeviceToHost, st)); ctx->stats.d2h_bytes += (int64_t)nn * 4; }
    ctx->rb_row_sizes = nullptr;
    if (row_sizes && ctx->world > 1) {
        // sharded set: this rank's matrices are an N-th of the fixed-layout table -- ship only those, scatter them on the host in _end
        const std::vector<PairTask>& own = ctx->owned;
        ctx->rs_off_h.resize(own.size() + 1);
        long long tot = 0;
        for (size_t k = 0; k < own.size(); ++k) { ctx->rs_off_h[k] = tot; tot += own[k].L1 + 2 + own[k].L2 + 2; }
        ctx->rs_off_h[own.size()] = tot;
        if ((size_t)tot + 64 > ctx->rs_stage_cap) {
            if (ctx->h_rs_stage) cudaFreeHost(ctx->h_rs_stage);
            ctx->h_rs_stage = nullptr; ctx->rs_stage_cap = 0;
            CK(cudaHostAlloc((void**)&ctx->h_rs_stage, ((size_t)tot + tot / 8 + 64) * 2, cudaHostAllocDefault));
            ctx->rs_stage_cap = (size_t)tot + tot / 8 + 64;
        }
        if (own.size() + 1 > ctx->rs_off_cap) {
            free_dev(ctx->d_rs_off); ctx->d_rs_off = nullptr;
            CK(cudaMalloc(&ctx->d_rs_off, (own.size() + own.size() / 4 + 64) * sizeof(long long)));
            ctx->rs_off_cap = own.size() + own.size() / 4 + 64;
        }
        if (own.size() > ctx->rs_tasks_cap) {        // own copy of the task list: d_tasks belongs to the stages running beside this read-back
            free_dev(ctx->d_rs_tasks); ctx->d_rs_tasks = nullptr;
            CK(cudaMalloc(&ctx->d_rs_tasks, (own.size() + own.size() / 4 + 64) * sizeof(PairTask)));
            ctx->rs_tasks_cap = own.size() + own.size() / 4 + 64;
        }
        CK(cudaMemcpyAsync(ctx->d_rs_tasks, own.data(), own.size() * sizeof(PairTask), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_rs_off, ctx->rs_off_h.data(), (own.size() + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        if (!own.empty()) {
            k_row_sizes_owned<<<grid, 256, 0, st>>>(ctx->d_rs_tasks, (int)own.size(), ctx->n, ctx->d_rp_off, s.rp_pool, ctx->d_rs_off, d_sizes);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(ctx->h_rs_stage, d_sizes, (size_t)tot * 2, cudaMemcpyDeviceToHost, st));
        }
        ctx->stats.d2h_bytes += tot * 2; ctx->stats.launches += 1;
        ctx->rb_row_sizes = row_sizes; ctx->rb_rs_total = tot;
    } else if (row_sizes) {
        k_row_sizes<<<grid, 256, 0, st>>>(s.rp_pool, d_sizes, ctx->rp_total);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(row_sizes, d_sizes, (size_t)ctx->rp_total * 2, cudaMemcpyDeviceToHost, st));
        ctx->stats.d2h_bytes += ctx->rp_total * 2; ctx->stats.launches += 1;
    }
    if (cells && used) {
        k_pack_cells<<<grid, 256, 0, st>>>(s.cells, d_cells, (long long)used);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(cells, d_cells, (size_t)used * 4, cudaMemcpyDeviceToHost, st));
        ctx->stats.d2h_bytes += (int64_t)used * 4; ctx->stats.launches += 1;
    }
    ctx->rb_set = ctx->cur;
    return MLP_OK;
}

extern "C" int mlp_get_csr_packed_end(mlp_ctx* ctx) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->rb_set < 0) return MLP_OK;
    cudaSetDevice(ctx->device);
    ctx->rb_set = -1;
    CK(cudaStreamSynchronize(ctx->stream_rb));
    if (ctx->rb_row_sizes) {          // sharded set: the owned matrices' row sizes go to their places in the caller's fixed-layout table
        const std::vector<PairTask>& own = ctx->owned;
        uint16_t* dst = ctx->rb_row_sizes;
        const int n = ctx->n;
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < (long long)own.size(); ++k) {
            const PairTask& t = own[k];
            const unsigned short* src = ctx->h_rs_stage + ctx->rs_off_h[k];
            std::memcpy(dst + ctx->rp_off_h[(size_t)t.a * n + t.b], src, (size_t)(t.L1 + 2) * 2);
            std::memcpy(dst + ctx->rp_off_h[(size_t)t.b * n + t.a], src + t.L1 + 2, (size_t)(t.L2 + 2) * 2);
        }
        ctx->rb_row_sizes = nullptr;
    }
    return MLP_OK;
}

extern "C" int mlp_alloc_pinned(int64_t bytes, void** out) {
    if (!out || bytes <= 0) return MLP_E_ARG;
    return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? MLP_OK : MLP_E_CUDA;
}

extern "C" void mlp_free_pinned(void* p) { if (p) cudaFreeHost(p); }

extern "C" int mlp_relax(mlp_ctx* ctx, int flavour, const float* weights, const float* seldist_nxn,
                         float selectivity, float selfweight, float cutoff) {
    if (!ctx) return MLP_E_ARG;
    END_READBACK(ctx);                               // the relaxation writes the set a split read-back may still be reading
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }
    if (!ctx->have_sets || ctx->flavour_of_set < 0) { ctx->err = "run mlp_posterior_all_pairs first"; return MLP_E_STATE; }
    cudaSetDevice(ctx->device);
    const int n = ctx->n;
    const bool resident_tree = (flavour == MLP_QP && !weights && !seldist_nxn && ctx->tree_resident);   // left on the device by mlp_qp_guide_tree_device
    if (flavour == MLP_QP && !resident_tree && (!weights || !seldist_nxn)) return MLP_E_ARG;
    if (flavour == MLP_QP && !(selectivity > 0)) return MLP_E_ARG;
    const int in = ctx->cur, out = 1 - ctx->cur;
    ctx->stats = mlp_stage_stats{};
    if (flavour == MLP_QP && !resident_tree) {
        ctx->tree_resident = false;
        if (n > ctx->weights_cap) {                   // sized for the current family (a context may see many families)
            free_dev(ctx->d_weights); free_dev(ctx->d_seldist);
            ctx->d_weights = nullptr; ctx->d_seldist = nullptr; ctx->weights_cap = 0;
            CK(cudaMalloc(&ctx->d_weights, n * sizeof(float)));
            CK(cudaMalloc(&ctx->d_seldist, (size_t)n * n * sizeof(float)));
            ctx->weights_cap = n;
        }
        CK(cudaMemcpyAsync(ctx->d_weights, weights, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_seldist, seldist_nxn, (size_t)n * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (int64_t)n * 4 + (int64_t)n * n * 4;
    }
    // Task order = 2-D tiles of the pair grid: the CTAs resident together then work on a TxT block of pairs and,
    // for each third sequence z, touch only 2T matrices (S_x z for T values of x, S_y z for T values of y) -> L2 reuse.
    // The order depends only on the owned pair list, so it is computed once per shard (cached in the context).
    if (ctx->relax_tasks.size() != ctx->owned.size() || ctx->relax_tasks_n != n) {
        ctx->relax_tasks = ctx->owned;
        const int T = 48;
        std::sort(ctx->relax_tasks.begin(), ctx->relax_tasks.end(), [T](const PairTask& x, const PairTask& y) {
            const int xa = x.a / T, xb = x.b / T, ya = y.a / T, yb = y.b / T;
            if (xa != ya) return xa < ya;
            if (xb != yb) return xb < yb;
            return x.pidx < y.pidx;
        });
        ctx->relax_tasks_n = n;
        ctx->relax_tasks_on_device = false;
    }
    const std::vector<PairTask>& tasks = ctx->relax_tasks;
    int rc = MLP_OK;
    if (tasks.size() > ctx->relax_tasks_cap) {       // the tile-ordered list has its own device copy: it only changes with the family or the shard
        free_dev(ctx->d_relax_tasks); ctx->d_relax_tasks = nullptr;
        ctx->relax_tasks_cap = tasks.size() + tasks.size() / 4 + 64;
        CK(cudaMalloc(&ctx->d_relax_tasks, ctx->relax_tasks_cap * sizeof(PairTask)));
        ctx->relax_tasks_on_device = false;
    }
    int maxL1 = 0, maxL2 = 0;
    for (const PairTask& t : tasks) { maxL1 = std::max(maxL1, t.L1); maxL2 = std::max(maxL2, t.L2); }
    const int bps = relax_blk_max_blocks_per_sm();
    const int grid = std::max(1, std::min(ctx->num_sms * bps, (int)tasks.size()));
    const long long warps = (long long)ctx->num_sms * 16 * (MLP_BLOCK / 32);
    rc = ensure_warp_buffers(ctx, warps, maxL1, maxL2, false, 1);
    if (rc != MLP_OK) return rc;
    // per-CTA scratch: weights and indices of the accepted third sequences, slice descriptors of the current band
    const long long wk_stride = relax_blk_scratch_words(n);
    const long long wk_units = (long long)ctx->num_sms * bps;
    if (wk_units * wk_stride > ctx->wk_warps) {
        free_dev(ctx->d_wk); ctx->d_wk = nullptr;
        CK(cudaMalloc(&ctx->d_wk, (size_t)(wk_units * wk_stride) * sizeof(float)));
        ctx->wk_warps = wk_units * wk_stride;
    }
    {   // the relaxed set can only shrink: size the output pool to what the input set holds
        unsigned long long used = 0;
        { int rcc = read_cursor(ctx, in, &used); if (rcc != MLP_OK) return rcc; }
        if (ctx->imported && ctx->own_cells > 0) used = (unsigned long long)ctx->own_cells;   // imported matrices are read, never written
        if ((long long)used + 1024 > ctx->set[out].cap) { rc = grow_cells(ctx, out, (long long)used + 1024, 0); if (rc != MLP_OK) return rc; }
    }
    if (!ctx->relax_tasks_on_device) {
        CK(cudaMemcpyAsync(ctx->d_relax_tasks, tasks.data(), tasks.size() * sizeof(PairTask), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (int64_t)(tasks.size() * sizeof(PairTask));
        ctx->relax_tasks_on_device = true;
    }
    CK(cudaMemsetAsync(ctx->set[out].cursor, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->set[out].nz_cnt, 0, (size_t)n * n * sizeof(int), ctx->stream));
    if (ctx->world > 1) {
        CK(cudaMemsetAsync(ctx->set[out].rp_pool, 0, (size_t)ctx->rp_total * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->set[out].nz_off, 0, (size_t)n * n * sizeof(long long), ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
    RelaxArgs ra = {};
    ra.tasks = ctx->d_relax_tasks; ra.ntasks = (int)tasks.size(); ra.counter = ctx->d_counter;
    ra.n = n; ra.flavour = flavour; ra.cutoff = cutoff; ra.rp_off = ctx->d_rp_off;
    ra.in = ctx->set[in]; ra.out = ctx->set[out];
    ra.weights = ctx->d_weights; ra.seldist = ctx->d_seldist; ra.selectivity = selectivity; ra.selfweight = selfweight;
    ra.wk_scratch = ctx->d_wk; ra.wk_stride = wk_stride; ra.err = ctx->d_err;
    ra.wide_span = getenv("MLP_RELAX_WIDE") ? atoi(getenv("MLP_RELAX_WIDE")) : (1 << 30);   // measured: merging wide rows is slower than giving them strips
    KernelTimer kt;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    kt.begin(MLP_K_RELAX_ID, ctx->stream);
    CK(relax_blk_launch(ra, grid, ctx->stream));
    kt.end(ctx->stream);
    ctx->stats.launches += 1;
    // second orientation of every new matrix
    KArgs a = {};
    a.tasks = ctx->d_relax_tasks; a.ntasks = (int)tasks.size(); a.counter = ctx->d_counter; a.n = n; a.flavour = flavour;
    a.rp_off = ctx->d_rp_off; a.out = ctx->set[out]; a.tfill = ctx->d_tfill; a.tfill_stride = ctx->tfill_stride; a.err = ctx->d_err;
    a.Cmax = 1;
    rc = launch_one(ctx, MLP_K_TRANSPOSE, a, (int)tasks.size(), kt, MLP_K_RELAX_ID);
    if (rc != MLP_OK) return rc;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_total = ms;
    kt.collect(ctx->stats);
    int err = 0;
    CK(cudaMemcpy(&err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
        cudaMemset(ctx->d_err, 0, sizeof(int));
        if (err & 4) { ctx->err = "relaxation: a staged copy never arrived"; return MLP_E_CUDA; }
        ctx->err = "sparse cell pool exhausted during relaxation";
        return MLP_E_CAPACITY;
    }
    ctx->cur = out;
    ctx->set_partial = (ctx->world > 1);
    ctx->imported = false;
    ctx->stats.pairs = (int64_t)tasks.size();
    unsigned long long cur = 0;
    { int rcc = read_cursor(ctx, out, &cur); if (rcc != MLP_OK) return rcc; }
    ctx->stats.nnz = (int64_t)(cur / 2);
    return MLP_OK;
}

extern "C" int mlp_last_stats(mlp_ctx* ctx, mlp_stage_stats* out) {
    if (!ctx || !out) return MLP_E_ARG;
    *out = ctx->stats;
    return MLP_OK;
}
