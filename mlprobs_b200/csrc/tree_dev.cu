// QuickProbs' guide tree on the device: UPGMA clustering (ClusterTree::build, ClusterTree.cpp:17-124), sequence weights
// (GuideTree::calculateSeqsWeights, GuideTree.cpp:114-154) and the subtree-size "selectivity" distances the consistency stage
// filters third sequences with (GuideTree::calculateSubtreeDistances, GuideTree.cpp:189-221), from the distance matrix that is
// already resident in HBM.  Same results as the host restatement (host_tree.cpp) bit for bit: lexicographically first minimal
// pair (strict '<' in a row-major scan), size-weighted average with separately rounded float multiply / add / divide, weights
// accumulated from the leaf upwards and normalised by a sequential float sum.
//
// The N-1 merges are inherently sequential, each is O(N) parallel work: ONE CTA of 1024 threads runs the whole clustering, the
// distance matrix stays in global memory (4 MB at N = 1000: L2-resident), the per-row minima of the lower triangle (value and
// first column) and the tree arrays live in shared memory.  Per merge: block-wide arg-min over the row minima, parallel update
// of row / column si fused with the repair of the row minima (rows whose minimum pointed at a merged slot are rescanned, one
// warp per row; the others only compare against their new entry in column si).  The depth-first leaf order behind the subtree distances is walked by one
// thread in shared memory at the end; the N x N subtree-distance matrix is then written by a second kernel, one CTA per inner node
// (every inner node is the lowest common ancestor of exactly |left leaves| x |right leaves| pairs).
#include "ctx.h"
#include <climits>

namespace {

#define UP_T 1024

struct UpgmaArgs {
    int n; float* D; float first_best;
    int* parent; int* lch; int* rch; float* branch; int* leaves;   // 2n-1 entries each (global)
    int* order; int* lo;                                          // n / 2n-1: leaf at depth-first position p, first position under a node
    float* weights; float min_weight;
    int* status;                                                  // 0 ok, 1 no joinable pair (first_best never undercut), 2 negative distance
};

__device__ __forceinline__ void warp_rescan(const float* row, const unsigned char* alive, int i, int lane, float& best, int& arg) {   // no __restrict__: the matrix changes while the kernel runs
    // smallest row[j] over alive j < i, first such j (host_tree.cpp rescan / the reference's strict '<' scan); 3.0 = nothing found
    float b = 3.0f; int a = -1;
    for (int j = lane; j < i; j += 32)
        if (alive[j]) { const float v = row[j]; if (v < b) { b = v; a = j; } }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float ob = __shfl_xor_sync(MLP_FULL, b, d);
        const int oa = __shfl_xor_sync(MLP_FULL, a, d);
        if (oa >= 0 && (ob < b || (ob == b && (a < 0 || oa < a)))) { b = ob; a = oa; }
    }
    best = b; arg = a;
}

__global__ void __launch_bounds__(UP_T) k_upgma(UpgmaArgs a) {
    extern __shared__ __align__(16) unsigned char up_raw[];
    const int n = a.n, total = 2 * n - 1;
    float* rowmin = reinterpret_cast<float*>(up_raw);            // [n]
    int* rowarg = reinterpret_cast<int*>(rowmin + n);            // [n]
    int* slot_node = rowarg + n;                                 // [n]
    int* list = slot_node + n;                                   // [n] rows to rescan
    int* s_l = list + n;                                         // [2n-1] children (also kept in global for the host)
    int* s_r = s_l + total;
    int* s_leaves = s_r + total;                                 // [2n-1]
    unsigned char* alive = reinterpret_cast<unsigned char*>(s_leaves + total);   // [n]
    __shared__ float red_v[32]; __shared__ int red_i[32];
    __shared__ int s_si, s_sj, s_nlist, s_bad; __shared__ float s_best; __shared__ unsigned s_isize, s_jsize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* D = a.D;

    if (tid == 0) { s_bad = 0; s_nlist = 0; }
    for (int i = tid; i < n; i += UP_T) { alive[i] = 1; slot_node[i] = i; }
    for (int v = tid; v < total; v += UP_T) { s_l[v] = -1; s_r[v] = -1; s_leaves[v] = v < n ? 1 : 0; a.parent[v] = -1; a.branch[v] = 0.0f; }
    __syncthreads();
    // negative distances are rejected (host_tree.cpp:33-34)
    for (long long k = tid; k < (long long)n * n; k += UP_T) { const int i = (int)(k / n), j = (int)(k % n); if (j < i && D[k] < 0.0f) s_bad = 2; }
    for (int i = warp; i < n; i += UP_T / 32) {
        float b; int g;
        warp_rescan(D + (size_t)i * n, alive, i, lane, b, g);
        if (lane == 0) { rowmin[i] = b; rowarg[i] = g; }
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) *a.status = s_bad; return; }

    for (int node = n; node < total; ++node) {
        // ---- 1. the lexicographically first minimal pair: smallest row minimum below first_best, smallest row among equals.
        // Row minima are repaired LAZILY: a row whose minimum pointed at a merged slot is only marked (rowarg = -2) and keeps its old
        // minimum, which stays a lower bound of the true one for ever (UPGMA replaces entries by means of two entries, never by
        // anything smaller).  Only marked rows whose bound could beat or tie the best exact minimum are rescanned before the
        // decision -- on star-like families one growing cluster is the nearest neighbour of half the rows, and rescanning all of
        // them at every merge was 11 of 17 us per merge.
        for (;;) {
            float bv = a.first_best; int bi = -1;
            for (int i = tid; i < n; i += UP_T)
                if (alive[i] && rowarg[i] >= 0 && rowmin[i] < bv) { bv = rowmin[i]; bi = i; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float ov = __shfl_xor_sync(MLP_FULL, bv, d);
                const int oi = __shfl_xor_sync(MLP_FULL, bi, d);
                if (oi >= 0 && (ov < bv || (ov == bv && (bi < 0 || oi < bi)))) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
            __syncthreads();
            if (warp == 0) {
                bv = red_v[lane]; bi = red_i[lane];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const float ov = __shfl_xor_sync(MLP_FULL, bv, d);
                    const int oi = __shfl_xor_sync(MLP_FULL, bi, d);
                    if (oi >= 0 && (ov < bv || (ov == bv && (bi < 0 || oi < bi)))) { bv = ov; bi = oi; }
                }
                if (lane == 0) { s_si = bi; s_best = bv; s_nlist = 0; }
            }
            __syncthreads();
            const float e0 = s_best;                 // best exact minimum (first_best when there is none)
            for (int i = tid; i < n; i += UP_T)
                if (alive[i] && rowarg[i] == -2 && rowmin[i] <= e0) list[atomicAdd(&s_nlist, 1)] = i;
            __syncthreads();
            const int nq = s_nlist;
            if (nq == 0) break;
            for (int k = warp; k < nq; k += UP_T / 32) {
                const int i = list[k];
                float b; int g;
                warp_rescan(D + (size_t)i * n, alive, i, lane, b, g);
                if (lane == 0) { rowmin[i] = b; rowarg[i] = g; }
            }
            __syncthreads();
        }
        if (tid == 0) {
            s_nlist = 0;
            const int bi = s_si;
            if (bi >= 0) {
                const int si = bi, sj = rowarg[si];
                const int ni = slot_node[si], nj = slot_node[sj];
                const float half = __fmul_rn(s_best, 0.5f);
                a.parent[ni] = node; a.parent[nj] = node; a.branch[ni] = half; a.branch[nj] = half;
                s_l[node] = ni; s_r[node] = nj;
                s_isize = (unsigned)s_leaves[ni]; s_jsize = (unsigned)s_leaves[nj];
                s_leaves[node] = s_leaves[ni] + s_leaves[nj];
                alive[sj] = 0; slot_node[si] = node; s_sj = sj;
            }
        }
        __syncthreads();
        const int si = s_si, sj = s_sj;
        if (si < 0) { if (tid == 0) *a.status = 1; return; }
        // ---- 2. row / column si <- size-weighted mean of rows si and sj (ClusterTree.cpp:96-103), in place; the thread that computes
        // the new d[idx][si] also repairs row idx's minimum (it owns that row's entry of the shared-memory arrays): rows whose
        // minimum pointed at a merged slot go to the rescan list, the others only compare against the new value
        {
            const float fi = (float)s_isize, fj = (float)s_jsize, fs = (float)(s_isize + s_jsize);
            for (int idx = tid; idx < n; idx += UP_T) {
                if (!alive[idx]) continue;
                if (idx == si) { list[atomicAdd(&s_nlist, 1)] = idx; continue; }
                const float idist = D[(size_t)si * n + idx], jdist = D[(size_t)sj * n + idx];
                const float v = __fdiv_rn(__fadd_rn(__fmul_rn(idist, fi), __fmul_rn(jdist, fj)), fs);
                D[(size_t)si * n + idx] = v; D[(size_t)idx * n + si] = v;
                if (idx <= sj) continue;
                const int g = rowarg[idx];
                if (g == -2) continue;                                           // already marked: its bound stays valid
                if (g == sj || g == si) { rowarg[idx] = -2; continue; }          // minimum pointed at a merged slot: lazy, see above
                if (idx > si && (v < rowmin[idx] || (v == rowmin[idx] && si < g))) { rowmin[idx] = v; rowarg[idx] = si; }
            }
        }
        __syncthreads();
        const int nl = s_nlist;
        for (int k = warp; k < nl; k += UP_T / 32) {
            const int i = list[k];
            float b; int g;
            warp_rescan(D + (size_t)i * n, alive, i, lane, b, g);
            if (lane == 0) { rowmin[i] = b; rowarg[i] = g; }
        }
        __syncthreads();
    }

    // ---- tree arrays for the host and the second kernel
    for (int v = tid; v < total; v += UP_T) { a.lch[v] = s_l[v]; a.rch[v] = s_r[v]; a.leaves[v] = s_leaves[v]; }
    // ---- weights: sum over the path to the root of branch / leaves-below, from the leaf upwards (GuideTree.cpp:114-154)
    for (int i = tid; i < n; i += UP_T) {
        float w = 0.0f;
        for (int c = i; a.parent[c] >= 0; c = a.parent[c]) w = __fadd_rn(w, __fdiv_rn(a.branch[c], (float)s_leaves[c]));
        a.weights[i] = w;
    }
    __syncthreads();
    if (tid == 0) {
        float wsum = 0.0f;
        for (int i = 0; i < n; ++i) wsum = __fadd_rn(wsum, a.weights[i]);       // sequential float sum, as the reference
        s_best = wsum;
        // depth-first leaf order, left before right (host_tree.cpp:95-107): position of every leaf, first position under every node
        int* stack = list;                                                        // depth never exceeds n
        int sp = 0, pos = 0;
        stack[sp++] = total - 1;
        while (sp > 0) {
            const int v = stack[--sp];
            a.lo[v] = pos;
            if (v < n) a.order[pos++] = v;
            else { stack[sp++] = s_r[v]; stack[sp++] = s_l[v]; }
        }
    }
    __syncthreads();
    {
        const float wsum = s_best;
        for (int i = tid; i < n; i += UP_T) {
            float w = (wsum == 0.0f) ? __fdiv_rn(1.0f, (float)n) : __fdiv_rn(a.weights[i], wsum);
            a.weights[i] = fmaxf(w, a.min_weight);
        }
    }
    if (tid == 0) *a.status = 0;
}

// distance(a, b) = number of leaves under the lowest common ancestor (GuideTree.cpp:189-221)
__global__ void k_subtree_dist(int n, const int* __restrict__ lch, const int* __restrict__ rch, const int* __restrict__ leaves,
                               const int* __restrict__ order, const int* __restrict__ lo, const int* __restrict__ status, float* __restrict__ out) {
    if (*status != 0) return;                       // the clustering failed: there is no tree
    const int v = n + blockIdx.x;
    const int l = lch[v], r = rch[v];
    const int nl = leaves[l], nr = leaves[r];
    const float d = (float)(nl + nr);
    const int* L = order + lo[l]; const int* R = order + lo[r];
    for (long long k = threadIdx.x; k < (long long)nl * nr; k += blockDim.x) {
        const int x = L[k / nr], y = R[k % nr];
        out[(size_t)x * n + y] = d; out[(size_t)y * n + x] = d;
    }
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < n; i += blockDim.x) out[(size_t)i * n + i] = 0.0f;
}

size_t upgma_smem(int n) { return (size_t)n * (4 + 4 + 4 + 4 + 1) + (size_t)(2 * n - 1) * 12 + 64; }

}  // namespace

// Builds QuickProbs' guide tree from the distance matrix resident on the device (which is left untouched: the clustering works on
// a scratch copy).  The saturated weights and the subtree distances stay resident for mlp_relax / mlp_exchange_needed
// (pass NULL there); the host gets the weights, the tree and -- on request -- the subtree distances.
extern "C" int mlp_qp_guide_tree_device(mlp_ctx* ctx, float min_weight, float* weights_out, int32_t* parent_out, int32_t* left_out,
                                        int32_t* right_out, float* seldist_out) {
    if (!ctx || !weights_out) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    const int n = ctx->n;
    if (n < 2 || !ctx->d_dist) { ctx->err = "run mlp_posterior_all_pairs first"; return MLP_E_STATE; }
    if (ctx->dist_partial) { ctx->err = "the distance matrix of a sharded stage must be exchanged first (mlp_exchange_distances)"; return MLP_E_STATE; }
    const size_t smem = upgma_smem(n);
    if (smem > 200 * 1024) { ctx->err = "family too large for the single-CTA device tree (use mlp_qp_guide_tree_ex)"; return MLP_E_UNSUPPORTED; }
    const int total = 2 * n - 1;
    if (n > ctx->weights_cap) {
        free_dev(ctx->d_weights); free_dev(ctx->d_seldist);
        ctx->d_weights = nullptr; ctx->d_seldist = nullptr; ctx->weights_cap = 0;
        CK(cudaMalloc(&ctx->d_weights, n * sizeof(float)));
        CK(cudaMalloc(&ctx->d_seldist, (size_t)n * n * sizeof(float)));
        ctx->weights_cap = n;
    }
    if (n > ctx->tree_cap) {
        free_dev(ctx->d_tree); ctx->d_tree = nullptr; ctx->tree_cap = 0;
        // scratch copy of the matrix | parent, lch, rch, leaves, lo (2n-1 ints each) | branch (2n-1 floats) | order (n) | status
        CK(cudaMalloc(&ctx->d_tree, (size_t)n * n * sizeof(float) + (size_t)(6 * total + n + 4) * sizeof(int)));
        ctx->tree_cap = n;
    }
    float* D = (float*)ctx->d_tree;
    int* ip = (int*)(D + (size_t)n * n);
    UpgmaArgs a;
    a.n = n; a.D = D; a.first_best = 2.0f;   // ClusterTree.cpp:41: the first candidate must undercut 2.0
    a.parent = ip; a.lch = ip + total; a.rch = ip + 2 * total; a.leaves = ip + 3 * total; a.lo = ip + 4 * total;
    a.branch = (float*)(ip + 5 * total); a.order = ip + 6 * total; a.status = ip + 6 * total + n;
    a.weights = ctx->d_weights; a.min_weight = min_weight;
    CK(cudaMemcpyAsync(D, ctx->d_dist, (size_t)n * n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemsetAsync(a.status, 0xff, sizeof(int), ctx->stream));
    CK(cudaFuncSetAttribute(k_upgma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_upgma<<<1, UP_T, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    k_subtree_dist<<<n - 1, 256, 0, ctx->stream>>>(n, a.lch, a.rch, a.leaves, a.order, a.lo, a.status, ctx->d_seldist);
    CK(cudaGetLastError());
    int status = -1;
    CK(cudaMemcpyAsync(weights_out, ctx->d_weights, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (parent_out) CK(cudaMemcpyAsync(parent_out, a.parent, total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (left_out) CK(cudaMemcpyAsync(left_out, a.lch, total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (right_out) CK(cudaMemcpyAsync(right_out, a.rch, total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (seldist_out) CK(cudaMemcpyAsync(seldist_out, ctx->d_seldist, (size_t)n * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&status, a.status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.launches += 2;
    ctx->stats.d2h_bytes += (int64_t)n * 4 + (seldist_out ? (int64_t)n * n * 4 : 0);
    if (status != 0) {
        ctx->tree_resident = false;
        ctx->err = status == 2 ? "negative distance" : "no pair of clusters closer than the reference's initial bound";
        return MLP_E_ARG;
    }
    ctx->tree_resident = true;
    return MLP_OK;
}

// Test hook: replaces the resident distance matrix (symmetric n*n floats), so that the device tree can be checked on matrices a
// posterior stage would not produce (masses of exact ties).
extern "C" int mlp_debug_set_distances(mlp_ctx* ctx, const float* nxn) {
    if (!ctx || !nxn) return MLP_E_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->n < 2 || !ctx->d_dist) { ctx->err = "set sequences first"; return MLP_E_STATE; }
    CK(cudaMemcpy(ctx->d_dist, nxn, (size_t)ctx->n * ctx->n * sizeof(float), cudaMemcpyHostToDevice));
    ctx->dist_partial = false; ctx->tree_resident = false;
    return MLP_OK;
}
