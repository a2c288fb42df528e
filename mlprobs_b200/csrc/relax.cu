// Consistency transformation as a masked sparse x sparse product, one warp per output pair (x, y).
//
// Reference: cpnp MSA::DoRelaxation / Relax / Relax1 (MSA.cpp:1172-1360), QP ConsistencyStage::doRelaxation /
// relax (ConsistencyStage.cpp:133-300).  Both densify S_xy, scatter-add S_xz * S_zy for every third sequence z
// into the dense matrix and finally zero everything outside the old pattern of S_xy.  Only cells of the old
// pattern survive, so this kernel computes exactly those: each lane owns ONE output cell (r, c) and accumulates
//      acc = init * S_xy[r][c] ;  for z ascending: for (q, v1) in row r of S_xz (q ascending): acc += [w_z *] v1 * S_zy[q][c]
// which is the order in which the reference's scatter-adds reach that cell (z ascending, then z-residue ascending;
// Relax1 and the transposed Relax visit a given cell in the same order).  Float multiply and add stay separate.
// Both orientations of every matrix are resident, so S_xz and S_zy are read directly for any z.
#include "posterior.cuh"


// ---- chunked, shared-memory staged version -------------------------------------------------------------------
// A chunk is RELAX_G*32 consecutive cells of S_xy (row-major), i.e. a handful of rows r0..r1 and a narrow column span
// cmin..cmax.  For every third sequence z the warp stages exactly the slices it needs:
//   rows r0..r1 of S_xz (one contiguous cell range, coalesced)                         -> smem A (sparse, sorted by q)
//   rows cmin..cmax of S_yz (= columns cmin..cmax of S_zy) restricted to q in [qmin,qmax] -> smem W, a DENSE window
// and every lane walks row r of A in ascending q and adds [w*]S_xz[r][q]*W[c][q].  Absent entries of W are +0, and
// acc + (+0) == acc exactly, so the result and the order of the non-zero adds are the reference's (q ascending).
// Chunks whose slices do not fit fall back to a sorted merge-join straight from global memory.
#define RELAX_G 4
#define RELAX_RMAX 64       // rows per chunk that fit the staged row-pointer slice
#define RELAX_CAPA 384      // staged cells of S_xz
#define RELAX_WIN 2304      // floats in the dense (column x z-residue) window of S_zy
struct __align__(16) RelaxSmem {
    float W[RELAX_WIN];
    int2 A[RELAX_CAPA];
    int rpA[RELAX_RMAX + 2];
    int pad[2];
};

__device__ __forceinline__ float merge_join(const int2* __restrict__ pa, const int2* __restrict__ ea,
                                            const int2* __restrict__ pb, const int2* __restrict__ eb,
                                            float acc, float w, bool weighted) {
    if (pa >= ea || pb >= eb) return acc;
    int2 x = *pa, y = *pb;
    for (;;) {
        if (x.x == y.x) {
            const float v1 = __int_as_float(x.y), v2 = __int_as_float(y.y);
            const float prod = weighted ? __fmul_rn(__fmul_rn(w, v1), v2) : __fmul_rn(v1, v2);   // ConsistencyStage.cpp:294 / MSA.cpp:1316
            acc = __fadd_rn(acc, prod);
            if (++pa >= ea || ++pb >= eb) break;
            x = *pa; y = *pb;
        } else if (x.x < y.x) {
            if (++pa >= ea) break;
            x = *pa;
        } else {
            if (++pb >= eb) break;
            y = *pb;
        }
    }
    return acc;
}

__global__ void __launch_bounds__(MLP_BLOCK) k_relax(RelaxArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    RelaxSmem& sm = reinterpret_cast<RelaxSmem*>(smem_raw)[warp];
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* wk = a.wk_scratch + gw * a.wk_stride;          // [n] weight of the m-th accepted z
    int* kl = reinterpret_cast<int*>(wk + a.n);           // [n] index  of the m-th accepted z
    const int n = a.n;
    const bool weighted = (a.flavour == 0);
    for (;;) {
        int ti = 0;
        if (lane == 0) ti = atomicAdd(a.counter, 1);
        ti = __shfl_sync(MLP_FULL, ti, 0);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        const int i = t.a, j = t.b;
        const long long sIJ = (long long)i * n + j;
        const int* rp_ij = a.in.rp_pool + a.rp_off[sIJ];
        const int2* c_ij = a.in.cells + a.in.nz_off[sIJ];
        const int nnz = a.in.nz_cnt[sIJ];
        int* orp = a.out.rp_pool + a.rp_off[sIJ];
        for (int r = lane; r <= t.L1 + 1; r += 32) orp[r] = 0;

        // ---- list of third sequences and the normaliser
        int nk = 0;
        float norm;
        if (weighted) {
            // ConsistencyStage.cpp:181-216: z accepted <=> max(d[x][z], d[y][z]) <= selectivity (Deterministic filter)
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                bool ok = false;
                if (k < n && k != i && k != j) ok = fmaxf(a.seldist[(long long)i * n + k], a.seldist[(long long)j * n + k]) <= a.selectivity;
                const unsigned m = __ballot_sync(MLP_FULL, ok);
                if (ok) kl[nk + __popc(m & ((1u << lane) - 1u))] = k;
                nk += __popc(m);
            }
            float wi_wj = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(__fsub_rn(a.selfweight, 1.0f), (float)nk), a.selectivity));
            wi_wj = __fmul_rn(wi_wj, __fadd_rn(a.weights[i], a.weights[j]));
            __syncwarp();
            for (int m = lane; m < nk; m += 32) wk[m] = __fdiv_rn(a.weights[kl[m]], wi_wj);
            __syncwarp();
            float sumW = 1.0f;   // ConsistencyStage.cpp:213: sequential float sum in z order
            if (lane == 0) for (int m = 0; m < nk; ++m) sumW = __fadd_rn(sumW, wk[m]);
            norm = __shfl_sync(MLP_FULL, sumW, 0);
        } else {
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                const bool ok = (k < n && k != i && k != j);
                const unsigned m = __ballot_sync(MLP_FULL, ok);
                if (ok) kl[nk + __popc(m & ((1u << lane) - 1u))] = k;
                nk += __popc(m);
            }
            norm = (float)n;   // MSA.cpp:1234
        }
        __syncwarp();

        long long obase = 0;
        if (lane == 0) {
            obase = (long long)atomicAdd(a.out.cursor, (unsigned long long)nnz);
            a.out.nz_off[sIJ] = obase;
            if (obase + nnz > a.out.cap) atomicOr(a.err, 2);
        }
        obase = __shfl_sync(MLP_FULL, obase, 0);
        const bool room = (obase + nnz <= a.out.cap);
        int kept_total = 0;
        int rowhint = 1;
        for (int c0 = 0; c0 < nnz; c0 += 32 * RELAX_G) {
            int rr[RELAX_G], cc[RELAX_G];
            float acc[RELAX_G];
            bool okc[RELAX_G];
            int cmin = 0x7fffffff, cmax = 0, rmin = 0x7fffffff, rmax = 0;
#pragma unroll
            for (int g = 0; g < RELAX_G; ++g) {
                const int cidx = c0 + g * 32 + lane;
                okc[g] = cidx < nnz;
                rr[g] = 1; cc[g] = 0; acc[g] = 0.0f;
                if (okc[g]) {
                    const int2 cell = c_ij[cidx];
                    cc[g] = cell.x;
                    const float v0 = __int_as_float(cell.y);
                    acc[g] = weighted ? v0 : __fadd_rn(v0, v0);   // MSA.cpp:1211-1213 doubles the matrix first
                    int lo = rowhint, hi = t.L1;
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rp_ij[mid] <= cidx) lo = mid; else hi = mid - 1; }
                    rr[g] = lo;
                    cmin = min(cmin, cc[g]); cmax = max(cmax, cc[g]); rmin = min(rmin, lo); rmax = max(rmax, lo);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                cmin = min(cmin, __shfl_xor_sync(MLP_FULL, cmin, d)); cmax = max(cmax, __shfl_xor_sync(MLP_FULL, cmax, d));
                rmin = min(rmin, __shfl_xor_sync(MLP_FULL, rmin, d)); rmax = max(rmax, __shfl_xor_sync(MLP_FULL, rmax, d));
            }
            const int R = rmax - rmin + 1, CW = cmax - cmin + 1;
            for (int m = 0; m < nk; ++m) {
                const int k = kl[m];
                const float w = weighted ? wk[m] : 1.0f;
                const long long sIK = (long long)i * n + k, sJK = (long long)j * n + k;
                const int* rp_ik = a.in.rp_pool + a.rp_off[sIK];
                const int2* c_ik = a.in.cells + a.in.nz_off[sIK];
                const int* rp_jk = a.in.rp_pool + a.rp_off[sJK];
                const int2* c_jk = a.in.cells + a.in.nz_off[sJK];
                bool staged = false;
                int a0 = 0, qmin = 0, QW = 1;
                if (R <= RELAX_RMAX) {
                    for (int x = lane; x <= R; x += 32) sm.rpA[x] = rp_ik[rmin + x];
                    __syncwarp();
                    a0 = sm.rpA[0];
                    const int na = sm.rpA[R] - a0;
                    if (na <= RELAX_CAPA) {
                        int qlo = 0x7fffffff, qhi = -1;
                        for (int x = lane; x < na; x += 32) { const int2 e = c_ik[a0 + x]; sm.A[x] = e; qlo = min(qlo, e.x); qhi = max(qhi, e.x); }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) { qlo = min(qlo, __shfl_xor_sync(MLP_FULL, qlo, d)); qhi = max(qhi, __shfl_xor_sync(MLP_FULL, qhi, d)); }
                        if (na == 0) { qlo = 1; qhi = 1; }
                        qmin = qlo;
                        QW = (qhi - qlo + 1) | 1;                  // odd stride: lanes on neighbouring columns hit different banks
                        if (CW * QW <= RELAX_WIN) {
                            staged = true;
                            {
                                float4* w4 = reinterpret_cast<float4*>(sm.W);
                                const int n4 = (CW * QW + 3) >> 2;
                                for (int x = lane; x < n4; x += 32) w4[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            __syncwarp();
                            for (int c = cmin + lane; c <= cmax; c += 32) {   // one lane per row of S_yz
                                const int b = rp_jk[c], e = rp_jk[c + 1];
                                float* wrow = sm.W + (c - cmin) * QW - qlo;
                                for (int x = b; x < e; ++x) {
                                    const int2 cell = c_jk[x];
                                    if (cell.x >= qlo && cell.x <= qhi) wrow[cell.x] = __int_as_float(cell.y);
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if (staged) {
#pragma unroll
                    for (int g = 0; g < RELAX_G; ++g)
                        if (okc[g]) {
                            const int ra = rr[g] - rmin;
                            const int2* pa = sm.A + (sm.rpA[ra] - a0);
                            const int2* ea = sm.A + (sm.rpA[ra + 1] - a0);
                            const float* wrow = sm.W + (cc[g] - cmin) * QW - qmin;
                            float ac = acc[g];
                            for (; pa < ea; ++pa) {
                                const int2 e = *pa;
                                const float v1 = __int_as_float(e.y), v2 = wrow[e.x];
                                const float prod = weighted ? __fmul_rn(__fmul_rn(w, v1), v2) : __fmul_rn(v1, v2);   // ConsistencyStage.cpp:294 / MSA.cpp:1316
                                ac = __fadd_rn(ac, prod);
                            }
                            acc[g] = ac;
                        }
                } else {
#pragma unroll
                    for (int g = 0; g < RELAX_G; ++g)
                        if (okc[g])
                            acc[g] = merge_join(c_ik + rp_ik[rr[g]], c_ik + rp_ik[rr[g] + 1],
                                                c_jk + rp_jk[cc[g]], c_jk + rp_jk[cc[g] + 1], acc[g], w, weighted);
                }
                __syncwarp();
            }
#pragma unroll
            for (int g = 0; g < RELAX_G; ++g) {
                const float v = __fdiv_rn(acc[g], norm);
                const bool keep = okc[g] && (v >= a.cutoff);
                const unsigned km = __ballot_sync(MLP_FULL, keep);
                if (keep && room) {
                    const long long d = obase + kept_total + __popc(km & ((1u << lane) - 1u));
                    a.out.cells[d] = make_int2(cc[g], __float_as_int(weighted ? dev_quantize_u16(v) : v));
                    atomicAdd(&orp[rr[g] + 1], 1);
                }
                kept_total += __popc(km);
            }
            rowhint = max(rmax, 1);   // cells are row-major: the next chunk starts at or after this row
        }
        __syncwarp();
        __threadfence_block();
        // row counts -> row pointers
        int run = 0;
        for (int base = 1; base <= t.L1; base += 32) {
            const int r = base + lane;
            int inc = (r <= t.L1) ? orp[r + 1] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(MLP_FULL, inc, d); if (lane >= d) inc += o; }
            if (r <= t.L1) orp[r + 1] = run + inc;
            run += __shfl_sync(MLP_FULL, inc, 31);
        }
        if (lane == 0) a.out.nz_cnt[sIJ] = kept_total;
        __syncwarp();
    }
}

static size_t relax_smem() { return sizeof(RelaxSmem) * (MLP_BLOCK / 32); }

cudaError_t relax_launch(const RelaxArgs& a, int grid, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)relax_smem());
    if (e != cudaSuccess) return e;
    k_relax<<<grid, MLP_BLOCK, relax_smem(), st>>>(a);
    return cudaGetLastError();
}

int relax_max_blocks_per_sm() {
    int nb = 0;
    cudaFuncSetAttribute(k_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)relax_smem());
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_relax, MLP_BLOCK, relax_smem()) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
