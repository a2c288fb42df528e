// QuickProbs flavour, device side of the tail: the profile-profile posterior (weighted sum of the pairwise sparse
// posteriors of two aligned groups, ParallelProbabilisticModel::buildPosterior ParallelProbabilisticModel.cpp:301-444)
// computed from the sparse set that is already resident in HBM, so the set never travels to the host.
//
// Exactness: every dense cell receives its contributions in the reference's order (group-1 sequence major, group-2
// sequence minor) with separate FP32 multiply and add.  A pair contributes at most once to a cell, so all cells of one
// pair are applied concurrently and successive pairs are applied one after the other.
//
// Kernel layout: one CTA per row of the dense matrix (= one column of group 1's alignment).  All 256 threads gather:
// thread t takes pair (i, j0+t), follows rp_off/nz_off -> row pointers -> cells -> column mapping and stages
// (dense column, w*v) records in shared memory; warp 0 then applies the staged pairs in order to the row accumulator
// kept in shared memory.  HBM traffic is one read of the sparse rows involved (8 B/cell) plus the dense row written once.
#include "ctx.h"
#include "qp_tail.h"
#include <algorithm>
#include <cstring>
#include <memory>

namespace {

constexpr int PP_THREADS = 256;
constexpr int PP_CAP = 8;            // staged cells per pair; longer sparse rows take the in-place path

struct PPArgs {
    const int* invA;                 // [nA][l1+1]    residue index of group-1 sequence i at alignment column r (0 = gap)
    const int* mapB;                 // [nB][ldB]     alignment column of residue k of group-2 sequence j
    const int* idsA; const int* idsB;
    const double* wA; const double* wB;
    double total;
    int wmode;                       // qptail::WeightSpec::Mode
    int nA, nB, l1, l2, ldB, n;
    const long long* rp_off; const long long* nz_off; const int* rp_pool; const int2* cells;
    float* dense;                    // (l1+1) x (l2+1)
};

__global__ void __launch_bounds__(PP_THREADS, 4) k_profile_posterior(PPArgs a) {
    extern __shared__ float sm[];
    float* acc = sm;                                              // l2+1 floats (rounded up to a multiple of 4)
    const int accn = (a.l2 + 1 + 3) & ~3;
    int* st_c = (int*)(sm + accn);                                // [PP_THREADS][PP_CAP]
    float* st_x = (float*)(st_c + PP_THREADS * PP_CAP);
    int* s_cnt = (int*)(st_x + PP_THREADS * PP_CAP);
    float* s_w = (float*)(s_cnt + PP_THREADS);
    long long* s_base = (long long*)(s_w + PP_THREADS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x;                                     // dense row, row 0 stays zero
    int* s_ii = (int*)(s_base + PP_THREADS);                      // residue index of this row in the k-th group-1 sequence that has one
    int* s_ik = s_ii + a.nA;                                      // ... and which sequence that is
    __shared__ int s_nk;
    for (int c = tid; c <= a.l2; c += PP_THREADS) acc[c] = 0.0f;
    // ordered compaction of the group-1 sequences with a residue in this column (warp 0, ballot scan)
    if (warp == 0) {
        int nk = 0;
        for (int i0 = 0; i0 < a.nA; i0 += 32) {
            const int i = i0 + lane;
            const int ii = (i < a.nA && r > 0) ? a.invA[(size_t)i * (a.l1 + 1) + r] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, ii != 0);
            if (ii != 0) { const int k = nk + __popc(m & ((1u << lane) - 1u)); s_ii[k] = ii; s_ik[k] = i; }
            nk += __popc(m);
        }
        if (lane == 0) s_nk = nk;
    }
    __syncthreads();
    {
        // pairs in the reference's order: p = k * nB + j (k over the compacted group-1 list, j over group 2), 256 per chunk
        const long long npairs = (long long)s_nk * a.nB;
        for (long long p0 = 0; p0 < npairs; p0 += PP_THREADS) {
            const long long p = p0 + tid;
            int cnt = 0;
            if (p < npairs) {
                const int k = (int)(p / a.nB), j = (int)(p - (long long)k * a.nB);
                const int ii = s_ii[k], i = s_ik[k];
                const long long slot = (long long)a.idsA[i] * a.n + a.idsB[j];
                const long long rpo = a.rp_off[slot];
                const int s = a.rp_pool[rpo + ii], e = a.rp_pool[rpo + ii + 1];
                const long long base = a.nz_off[slot] + s;
                const double wa = a.wA[i], wb = a.wB[j];
                const float w = a.wmode == 0 ? (float)((wa * wb) / a.total)
                              : (a.wmode == 1 ? __fdiv_rn((float)(wa * wb), (float)a.total) : 1.0f);
                cnt = e - s;
                const int* mb = a.mapB + (size_t)j * a.ldB;
                const int m = cnt < PP_CAP ? cnt : PP_CAP;
                for (int q = 0; q < m; ++q) {
                    const int2 cell = a.cells[base + q];
                    st_c[tid * PP_CAP + q] = mb[cell.x];
                    st_x[tid * PP_CAP + q] = __fmul_rn(w, __int_as_float(cell.y));
                }
                s_w[tid] = w;
                s_base[tid] = base;
            }
            s_cnt[tid] = cnt;
            __syncthreads();
            if (warp == 0) {
                const int np = (int)min((long long)PP_THREADS, npairs - p0);
                // software pipeline: the record of pair t+1 is fetched before the read-modify-write of pair t, so the only
                // serial chain left per pair is LDS(acc) -> FADD -> STS
                const int sl = lane < PP_CAP ? lane : 0;
                int ct = s_cnt[0], c = st_c[sl];
                float x = st_x[sl];
                for (int t = 0; t < np; ++t) {
                    const int tn = (t + 1 < np) ? t + 1 : t;
                    const int ct_n = s_cnt[tn], c_n = st_c[tn * PP_CAP + sl];
                    const float x_n = st_x[tn * PP_CAP + sl];
                    if (lane < ct && lane < PP_CAP) acc[c] = __fadd_rn(acc[c], x);
                    if (ct > PP_CAP) {                            // long sparse row: remaining cells straight from HBM
                        const long long pt = p0 + t;
                        const int jt = (int)(pt - (pt / a.nB) * a.nB);
                        const int* mb = a.mapB + (size_t)jt * a.ldB;
                        for (int q = PP_CAP + lane; q < ct; q += 32) {
                            const int2 cell = a.cells[s_base[t] + q];
                            const int cq = mb[cell.x];
                            acc[cq] = __fadd_rn(acc[cq], __fmul_rn(s_w[t], __int_as_float(cell.y)));
                        }
                    }
                    __syncwarp();
                    ct = ct_n; c = c_n; x = x_n;
                }
            }
            __syncthreads();
        }
    }
    float* out = a.dense + (size_t)r * (a.l2 + 1);
    for (int c = tid; c <= a.l2; c += PP_THREADS) out[c] = acc[c];
}


// MEA dynamic programme over the dense profile posterior (ProbabilisticModel::computeAlignment ProbabilisticModel.cpp:345-421)
// as a skewed wavefront inside ONE CTA: thread t owns C consecutive columns and is one row behind thread t-1, so the value
// it needs from its left neighbour was produced one step earlier.  Scores are exact (each cell is the reference's own
// max-of-three with its tie order), the 2-bit choices go to HBM and the host walks them back.
constexpr int MEA_MAXC = 8;
constexpr int MEA_DEPTH = 4;
struct MeaArgs {
    const float* dense; int l1, l2, C, T, B;
    unsigned char* tb;                 // [(l1)][T][B] 2-bit choices (0 = diagonal, 1 = left, 2 = up), rows 1..l1
    float* score;                      // the maximum sum = DP value of cell (l1, l2)
};

__global__ void __launch_bounds__(1024) k_mea_wavefront(MeaArgs a) {
    extern __shared__ float sm[];
    float* row = sm;                                              // T*C floats
    float2* edge = (float2*)(sm + (size_t)a.T * a.C);             // [2][T]
    const int t = threadIdx.x, C = a.C, j0 = t * C;
    const int W = a.l2 + 1;
    for (int c = 0; c < C; ++c) row[j0 + c] = 0.0f;
    edge[t] = make_float2(0.0f, 0.0f);
    edge[a.T + t] = make_float2(0.0f, 0.0f);
    // posteriors of the rows this thread reaches in the next MEA_DEPTH steps, in registers (one row per step; the loads of
    // row i + MEA_DEPTH are issued when row i is consumed, far enough ahead to cover the L2 latency)
    float pf[MEA_DEPTH][MEA_MAXC];
    auto load_row = [&](int i, float (&dst)[MEA_MAXC]) {
#pragma unroll
        for (int c = 0; c < MEA_MAXC; ++c) dst[c] = 0.0f;
        if (i >= 1 && i <= a.l1) {
            const float* p = a.dense + (size_t)i * W + j0;
#pragma unroll
            for (int c = 0; c < MEA_MAXC; ++c) if (c < C && j0 + c < W) dst[c] = p[c];
        }
    };
#pragma unroll
    for (int u = 0; u < MEA_DEPTH; ++u) load_row(u - t + 1, pf[u]);
    __syncthreads();
    const int steps = a.l1 + a.T - 1;
    for (int s0 = 0; s0 < steps; s0 += MEA_DEPTH) {
#pragma unroll
        for (int u = 0; u < MEA_DEPTH; ++u) {
            const int s = s0 + u;
            const int i = s - t + 1;
            if (s < steps && i >= 1 && i <= a.l1 && j0 < W) {
                float d = 0.0f, l = 0.0f;
                if (t > 0) { const float2 e = edge[((s + 1) & 1) * a.T + t - 1]; d = e.x; l = e.y; }
                unsigned bits = 0;
                float o = 0.0f, v = 0.0f;
#pragma unroll
                for (int c = 0; c < MEA_MAXC; ++c) {
                    if (c < C && j0 + c < W) {
                        const int j = j0 + c;
                        o = row[j];
                        if (j == 0) v = 0.0f;
                        else {
                            const float x1 = __fadd_rn(pf[u][c], d);
                            unsigned dir;
                            if (x1 >= l) { if (x1 >= o) { v = x1; dir = 0; } else { v = o; dir = 2; } }
                            else if (l >= o) { v = l; dir = 1; }
                            else { v = o; dir = 2; }
                            bits |= dir << (2 * c);
                        }
                        d = o; l = v; row[j] = v;
                    }
                }
                edge[(s & 1) * a.T + t] = make_float2(o, v);
                unsigned char* q = a.tb + ((size_t)(i - 1) * a.T + t) * a.B;
                q[0] = (unsigned char)(bits & 0xff);
                if (a.B > 1) q[1] = (unsigned char)(bits >> 8);
            }
            load_row(i + MEA_DEPTH, pf[u]);
            __syncthreads();
        }
    }
    if (a.score && j0 <= a.l2 && a.l2 < j0 + C) *a.score = row[a.l2];   // all rows done (the loop ends with a barrier)
}

__global__ void k_gather_dense(const float* __restrict__ dense, const long long* __restrict__ off, float* __restrict__ out, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = dense[off[k]];
}

size_t pp_smem_bytes(int l2, int nA) {
    const size_t accn = ((size_t)l2 + 1 + 3) & ~(size_t)3;
    return accn * 4 + (size_t)PP_THREADS * PP_CAP * 8 + (size_t)PP_THREADS * (4 + 4 + 8) + (size_t)nA * 8;
}

class DeviceProfilePosterior : public qptail::ProfilePosterior {
public:
    explicit DeviceProfilePosterior(mlp_ctx* c) : ctx(c) {}
    ~DeviceProfilePosterior() override {
        if (d_blob) cudaFree(d_blob);
        if (h_blob) cudaFreeHost(h_blob);
        if (d_dense) cudaFree(d_dense);
        if (h_dense) cudaFreeHost(h_dense);
        if (d_tb) cudaFree(d_tb);
        if (h_tb) cudaFreeHost(h_tb);
    }
    // `offsets` (optional): element offsets into the dense matrix that k_gather_dense will read; they ride in the same
    // host->device copy as the mappings (d_offsets points at them afterwards)
    int launch_profile(const qptail::Profile& A, const qptail::Profile& B, const qptail::WeightSpec& ws,
                       const std::vector<long long>* offsets = nullptr) {
        const int nA = A.count(), nB = B.count(), l1 = A.length(), l2 = B.length();
        int ldB = 1;
        for (int j = 0; j < nB; ++j) ldB = std::max(ldB, ctx->len[B.ids[j]] + 1);
        // blob layout: doubles first, then ints
        const size_t n_dbl = (size_t)nA + nB;
        const size_t n_inv = (size_t)(l1 + 1) * nA, n_map = (size_t)nB * ldB;
        const size_t ng = offsets ? offsets->size() : 0;
        const size_t off_at = (n_dbl * 8 + (n_inv + n_map + nA + nB) * 4 + 7) & ~(size_t)7;
        const size_t bytes = off_at + ng * 8;
        if (bytes > blob_cap) {
            if (d_blob) cudaFree(d_blob);
            if (h_blob) cudaFreeHost(h_blob);
            d_blob = nullptr; h_blob = nullptr;
            blob_cap = bytes + bytes / 2;
            CK(cudaMalloc(&d_blob, blob_cap));
            CK(cudaHostAlloc(&h_blob, blob_cap, cudaHostAllocDefault));
        }
        const size_t dn = (size_t)(l1 + 1) * (l2 + 1);
        if (dn > dense_cap) {
            if (d_dense) cudaFree(d_dense);
            if (h_dense) cudaFreeHost(h_dense);
            d_dense = nullptr; h_dense = nullptr;
            dense_cap = dn + dn / 2;
            CK(cudaMalloc(&d_dense, dense_cap * 4));
            CK(cudaHostAlloc(&h_dense, dense_cap * 4, cudaHostAllocDefault));
        }
        double* wA = (double*)h_blob;
        double* wB = wA + nA;
        int* invA = (int*)(wB + nB);
        int* mapB = invA + n_inv;
        int* idsA = mapB + n_map;
        int* idsB = idsA + nA;
        const double total = ws.total(A, B);                      // the reference's normaliser, in its order and precision
        for (int i = 0; i < nA; ++i) { wA[i] = ws.weight_of(A.ids[i]); idsA[i] = A.ids[i]; }
        for (int j = 0; j < nB; ++j) { wB[j] = ws.weight_of(B.ids[j]); idsB[j] = B.ids[j]; }
#pragma omp parallel for schedule(static) if ((long long)nA * l1 >= (1 << 16))
        for (int i = 0; i < nA; ++i) {                            // [i][column], row-major: sequential writes
            const char* row = A.rows[i].data();
            int* dst = invA + (size_t)i * (l1 + 1);
            int k = 0;
            dst[0] = 0;
            for (int c = 0; c < l1; ++c) dst[c + 1] = (row[c] != '-') ? ++k : 0;
        }
#pragma omp parallel for schedule(static) if ((long long)nB * l2 >= (1 << 16))
        for (int j = 0; j < nB; ++j) {
            const std::string& row = B.rows[j];
            int* m = mapB + (size_t)j * ldB;
            int k = 0;
            m[0] = 0;
            for (int c = 0; c < l2; ++c) if (row[c] != '-') m[++k] = c + 1;
        }
        const size_t smem = pp_smem_bytes(l2, nA);
        if (smem > 200 * 1024) { ctx->err = "profile too long for the row accumulator"; return MLP_E_UNSUPPORTED; }
        if (smem > smem_set) {
            CK(cudaFuncSetAttribute(k_profile_posterior, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set = smem;
        }
        if (ng) memcpy((char*)h_blob + off_at, offsets->data(), ng * 8);
        d_offsets = reinterpret_cast<const long long*>((const char*)d_blob + off_at);
        CK(cudaMemcpyAsync(d_blob, h_blob, bytes, cudaMemcpyHostToDevice, ctx->stream));
        PPArgs a;
        a.wA = (const double*)d_blob; a.wB = a.wA + nA;
        a.invA = (const int*)(a.wB + nB); a.mapB = a.invA + n_inv; a.idsA = a.mapB + n_map; a.idsB = a.idsA + nA;
        a.total = total; a.wmode = ws.mode; a.nA = nA; a.nB = nB; a.l1 = l1; a.l2 = l2; a.ldB = ldB; a.n = ctx->n;
        const CsrSetDev& S = ctx->set[ctx->cur];
        a.rp_off = ctx->d_rp_off; a.nz_off = S.nz_off; a.rp_pool = S.rp_pool; a.cells = S.cells;
        a.dense = d_dense;
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        k_profile_posterior<<<l1 + 1, PP_THREADS, smem, ctx->stream>>>(a);
        CK(cudaGetLastError());
        ctx->stats.launches += 1;
        ctx->stats.h2d_bytes += (int64_t)bytes;
        ctx->stats.pairs += (int64_t)nA * nB;
        return 0;
    }

    int finish_timing() {
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        ctx->stats.ms_total += ms;
        return 0;
    }

    int build(const qptail::Profile& A, const qptail::Profile& B, const qptail::WeightSpec& ws, const float** out) override {
        int rc = launch_profile(A, B, ws);
        if (rc < 0) return rc;
        const size_t dn = (size_t)(A.length() + 1) * (B.length() + 1);
        CK(cudaMemcpyAsync(h_dense, d_dense, dn * 4, cudaMemcpyDeviceToHost, ctx->stream));
        rc = finish_timing();
        if (rc < 0) return rc;
        ctx->stats.d2h_bytes += (int64_t)dn * 4;
        *out = h_dense;
        return 0;
    }

    // profile posterior + MEA wavefront on the device, 2-bit choices back to the host, traceback there
    int build_and_align(const qptail::Profile& A, const qptail::Profile& B, const qptail::WeightSpec& ws, std::string& path) override {
        return align_common(A, B, ws, path, nullptr, nullptr, nullptr);
    }
    int build_align_score(const qptail::Profile& A, const qptail::Profile& B, const qptail::WeightSpec& ws, const std::vector<long long>& offsets,
                          std::string& path, float* score, std::vector<float>& values) override {
        return align_common(A, B, ws, path, &offsets, score, &values);
    }
    int align_common(const qptail::Profile& A, const qptail::Profile& B, const qptail::WeightSpec& ws, std::string& path,
                     const std::vector<long long>* offsets, float* score, std::vector<float>* values) {
        const int l1 = A.length(), l2 = B.length();
        const int W = l2 + 1;
        if (W > 1024 * MEA_MAXC || l1 < 1 || l2 < 1) return 1;    // very wide profile: host dynamic programme on the dense matrix
        int T = std::min(1024, ((W + 3) / 4 + 31) / 32 * 32);
        const int C = (W + T - 1) / T;
        const int Bb = (C + 3) / 4;
        const size_t smem = ((size_t)T * C + 4 * (size_t)T) * 4;
        if (smem > 200 * 1024) return 1;
        int rc = launch_profile(A, B, ws, offsets);
        if (rc < 0) return rc;
        // result buffer: [2-bit choices, tbn bytes][pad to 16][score, 16 bytes][ng gathered values]: one device->host copy
        const size_t tbn = (size_t)l1 * T * Bb;
        const size_t ng = offsets ? offsets->size() : 0;
        const bool extras = score || ng;
        const size_t xoff = (tbn + 15) & ~(size_t)15;
        const size_t out_bytes = extras ? xoff + 16 + ng * 4 : tbn;
        if (out_bytes > tb_cap) {
            if (d_tb) cudaFree(d_tb);
            if (h_tb) cudaFreeHost(h_tb);
            d_tb = nullptr; h_tb = nullptr; tb_cap = 0;
            const size_t cap = out_bytes + out_bytes / 2 + 4096;
            CK(cudaMalloc(&d_tb, cap));
            CK(cudaHostAlloc(&h_tb, cap, cudaHostAllocDefault));
            tb_cap = cap;
        }
        if (smem > mea_smem_set) {
            CK(cudaFuncSetAttribute(k_mea_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mea_smem_set = smem;
        }
        MeaArgs m;
        float* d_score = extras ? reinterpret_cast<float*>(d_tb + xoff) : nullptr;
        float* d_val = reinterpret_cast<float*>(d_tb + xoff + 16);
        m.dense = d_dense; m.l1 = l1; m.l2 = l2; m.C = C; m.T = T; m.B = Bb; m.tb = d_tb; m.score = score ? d_score : nullptr;
        k_mea_wavefront<<<1, T, smem, ctx->stream>>>(m);
        CK(cudaGetLastError());
        ctx->stats.launches += 1;
        if (ng) {
            k_gather_dense<<<(unsigned)((ng + 255) / 256), 256, 0, ctx->stream>>>(d_dense, d_offsets, d_val, (int)ng);
            CK(cudaGetLastError());
            ctx->stats.launches += 1;
        }
        CK(cudaMemcpyAsync(h_tb, d_tb, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        rc = finish_timing();
        if (rc < 0) return rc;
        ctx->stats.d2h_bytes += (int64_t)out_bytes;
        if (score) *score = *reinterpret_cast<const float*>(h_tb + xoff);
        if (values) { values->resize(ng); if (ng) memcpy(values->data(), h_tb + xoff + 16, ng * 4); }
        path.clear();
        path.reserve((size_t)l1 + l2);
        int r = l1, c = l2;
        while (r != 0 || c != 0) {
            int dir;
            if (r == 0) dir = 1;
            else if (c == 0) dir = 2;
            else {
                const int t = c / C, k = c - t * C;
                dir = (h_tb[((size_t)(r - 1) * T + t) * Bb + (k >> 2)] >> ((k & 3) * 2)) & 3;
            }
            if (dir == 1) { --c; path.push_back('Y'); }
            else if (dir == 2) { --r; path.push_back('X'); }
            else { --c; --r; path.push_back('B'); }
        }
        std::reverse(path.begin(), path.end());
        return 0;
    }

private:
    mlp_ctx* ctx;
    void* d_blob = nullptr; void* h_blob = nullptr; size_t blob_cap = 0;
    float* d_dense = nullptr; float* h_dense = nullptr; size_t dense_cap = 0;
    size_t smem_set = 48 * 1024, mea_smem_set = 48 * 1024;
    unsigned char* d_tb = nullptr; unsigned char* h_tb = nullptr; size_t tb_cap = 0;   // traceback choices + score + gathered values (cpnp refinement)
    const long long* d_offsets = nullptr;                                              // inside d_blob, set by launch_profile
};

// One provider per context: its pinned / device staging buffers only ever grow, so a context that aligns many families
// (directory mode of the two executables) does not pay cudaHostAlloc / cudaMalloc again for every family.
DeviceProfilePosterior& provider_of(mlp_ctx* ctx) {
    if (!ctx->tail_prov) {
        ctx->tail_prov = new DeviceProfilePosterior(ctx);
        ctx->tail_prov_free = [](void* p) { delete static_cast<DeviceProfilePosterior*>(p); };
    }
    return *static_cast<DeviceProfilePosterior*>(ctx->tail_prov);
}

}  // namespace

extern "C" int mlp_qp_finish_alignment(mlp_ctx* ctx, const float* weights, const int32_t* left, const int32_t* right,
                                       int ref_iters, uint32_t ref_seed, char** rows_out, int32_t* aln_len) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }   // a split exchange must have landed before the set is read
    if (!rows_out || !aln_len) { ctx->err = "null output"; return MLP_E_ARG; }
    if (ctx->n < 1) { ctx->err = "no sequences"; return MLP_E_STATE; }
    if (ctx->n > 1) {
        if (!weights || !left || !right) { ctx->err = "weights and tree required"; return MLP_E_ARG; }
        if (!ctx->have_sets || ctx->flavour_of_set != MLP_QP) { ctx->err = "no QuickProbs sparse set on the device"; return MLP_E_STATE; }
        if (ctx->world > 1 && !ctx->nccl_comm) { ctx->err = "sharded set: call mlp_exchange first"; return MLP_E_STATE; }
    }
    CK(cudaSetDevice(ctx->device));
    ctx->stats = mlp_stage_stats{};
    std::vector<uint8_t> letters(ctx->codes_h.size());
    for (size_t k = 0; k < letters.size(); ++k) letters[k] = (uint8_t)('A' + ctx->codes_h[k]);
    std::vector<int32_t> len(ctx->len.begin(), ctx->len.end());
    DeviceProfilePosterior& prov = provider_of(ctx);
    qptail::TailOptions opt;
    opt.ref_iters = ref_iters;
    opt.ref_seed = ref_seed;
    qptail::Profile out;
    std::string err;
    const int rc = qptail::run_tail(ctx->n, len.data(), letters.data(), weights, left, right, prov, opt, out, err);
    if (rc < 0) { if (ctx->err.empty() || rc != MLP_E_CUDA) ctx->err = err; return rc; }
    const int L = out.length();
    char* buf = (char*)malloc((size_t)ctx->n * (size_t)std::max(L, 1));
    if (!buf) { ctx->err = "out of host memory"; return MLP_E_NOMEM; }
    for (int i = 0; i < ctx->n; ++i) memcpy(buf + (size_t)i * L, out.rows[i].data(), (size_t)L);
    *rows_out = buf;
    *aln_len = L;
    return MLP_OK;
}

extern "C" int mlp_cpnp_finish_alignment(mlp_ctx* ctx, const int32_t* iweights, const int32_t* left, const int32_t* right,
                                         int refine_reps, int pid, char** rows_out, int32_t* aln_len, int32_t* order_out) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }   // a split exchange must have landed before the set is read
    if (!rows_out || !aln_len) { ctx->err = "null output"; return MLP_E_ARG; }
    if (ctx->n < 2) { ctx->err = "no sequences"; return MLP_E_STATE; }
    if (!iweights || !left || !right) { ctx->err = "weights and tree required"; return MLP_E_ARG; }
    if (!ctx->have_sets || (ctx->flavour_of_set != MLP_CPNP_P0 && ctx->flavour_of_set != MLP_CPNP_P1)) { ctx->err = "no c_p_np_aln sparse set on the device"; return MLP_E_STATE; }
    if (ctx->world > 1 && !ctx->nccl_comm) { ctx->err = "sharded set: call mlp_exchange first"; return MLP_E_STATE; }
    CK(cudaSetDevice(ctx->device));
    ctx->stats = mlp_stage_stats{};
    std::vector<uint8_t> letters(ctx->codes_h.size());
    for (size_t k = 0; k < letters.size(); ++k) letters[k] = (uint8_t)('A' + ctx->codes_h[k]);
    std::vector<int32_t> len(ctx->len.begin(), ctx->len.end());
    DeviceProfilePosterior& prov = provider_of(ctx);
    qptail::Profile out;
    std::string err;
    const int rc = qptail::run_cpnp_tail(ctx->n, len.data(), letters.data(), iweights, left, right, prov, refine_reps, pid, out, err);
    if (rc < 0) { if (ctx->err.empty() || rc != MLP_E_CUDA) ctx->err = err; return rc; }
    const int L = out.length();
    char* buf = (char*)malloc((size_t)ctx->n * (size_t)std::max(L, 1));
    if (!buf) { ctx->err = "out of host memory"; return MLP_E_NOMEM; }
    for (int i = 0; i < ctx->n; ++i) {
        memcpy(buf + (size_t)i * L, out.rows[i].data(), (size_t)L);
        if (order_out) order_out[i] = out.ids[i];
    }
    *rows_out = buf;
    *aln_len = L;
    return MLP_OK;
}

// c_p_np_aln -p 1: the alignment graph is host work on a copy of the relaxed set (a scalar, order-dependent greedy over the
// sorted cells), the similar-set refinement after it runs its profile posteriors and MEA sweeps on the resident set.
extern "C" int mlp_cpnp_np_finish_alignment(mlp_ctx* ctx, int refine_reps, int64_t seed, char** rows_out, int32_t* aln_len) {
    if (!ctx) return MLP_E_ARG;
    if (ctx->exch_pending) { const int rce = mlp_exchange_end(ctx); if (rce != MLP_OK) return rce; }   // a split exchange must have landed before the set is read
    if (!rows_out || !aln_len) { ctx->err = "null output"; return MLP_E_ARG; }
    if (ctx->n < 2) { ctx->err = "no sequences"; return MLP_E_STATE; }
    if (!ctx->have_sets || (ctx->flavour_of_set != MLP_CPNP_P0 && ctx->flavour_of_set != MLP_CPNP_P1)) { ctx->err = "no c_p_np_aln sparse set on the device"; return MLP_E_STATE; }
    if (ctx->world > 1 && !ctx->nccl_comm) { ctx->err = "sharded set: call mlp_exchange first"; return MLP_E_STATE; }
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->n;
    const size_t nn = (size_t)n * n;
    std::vector<int64_t> rp_off(nn), nz_off(nn);
    int64_t rp_total = 0, cells_used = 0;
    int rc = mlp_csr_layout(ctx, rp_off.data(), &rp_total, &cells_used);
    if (rc != MLP_OK) return rc;
    std::vector<int32_t> rp_pool((size_t)rp_total);
    std::vector<int2> cells((size_t)cells_used + 1);
    rc = mlp_get_csr_raw(ctx, nz_off.data(), nullptr, rp_pool.data(), cells.data());
    if (rc != MLP_OK) return rc;
    std::vector<float> dist(nn);
    rc = mlp_get_distances(ctx, dist.data());
    if (rc != MLP_OK) return rc;
    ctx->stats = mlp_stage_stats{};
    std::vector<uint8_t> letters(ctx->codes_h.size());
    for (size_t k = 0; k < letters.size(); ++k) letters[k] = (uint8_t)('A' + ctx->codes_h[k]);
    std::vector<int32_t> len(ctx->len.begin(), ctx->len.end());
    const qptail::HostCsrView view{n, len.data(), rp_off.data(), nz_off.data(), rp_pool.data(), cells.data()};
    DeviceProfilePosterior& prov = provider_of(ctx);
    qptail::Profile out;
    std::string err;
    rc = qptail::run_cpnp_np_tail(view, letters.data(), dist.data(), prov, refine_reps, (long long)seed, out, err);
    if (rc < 0) { if (ctx->err.empty() || rc != MLP_E_CUDA) ctx->err = err; return rc; }
    const int L = out.length();
    char* buf = (char*)malloc((size_t)n * (size_t)std::max(L, 1));
    if (!buf) { ctx->err = "out of host memory"; return MLP_E_NOMEM; }
    for (int i = 0; i < n; ++i) memcpy(buf + (size_t)i * L, out.rows[i].data(), (size_t)L);
    *rows_out = buf;
    *aln_len = L;
    return MLP_OK;
}
