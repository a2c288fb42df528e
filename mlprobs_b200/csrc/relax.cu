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


__global__ void __launch_bounds__(MLP_BLOCK) k_relax(RelaxArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* wk = a.wk_scratch ? a.wk_scratch + gw * a.wk_stride : nullptr;
    const int n = a.n;
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

        float norm;   // divisor applied before masking: N (cpnp, MSA.cpp:1234) or sumW (QP, ConsistencyStage.cpp:226)
        if (a.flavour == 0) {
            // ConsistencyStage.cpp:181-216: accepted z <=> max(d[i][z], d[j][z]) <= selectivity (Deterministic filter)
            int acc_cnt = 0;
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                bool ok = false;
                if (k < n && k != i && k != j) ok = fmaxf(a.seldist[(long long)i * n + k], a.seldist[(long long)j * n + k]) <= a.selectivity;
                acc_cnt += __popc(__ballot_sync(MLP_FULL, ok));
                if (k < n) wk[k] = ok ? 1.0f : -1.0f;
            }
            float wi_wj = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(__fsub_rn(a.selfweight, 1.0f), (float)acc_cnt), a.selectivity));
            wi_wj = __fmul_rn(wi_wj, __fadd_rn(a.weights[i], a.weights[j]));
            __syncwarp();
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                if (k < n && wk[k] > 0.0f) wk[k] = __fdiv_rn(a.weights[k], wi_wj);
            }
            __syncwarp();
            float sumW = 1.0f;   // sequential float sum in z order
            if (lane == 0) for (int k = 0; k < n; ++k) { const float w = wk[k]; if (w >= 0.0f) sumW = __fadd_rn(sumW, w); }
            norm = __shfl_sync(MLP_FULL, sumW, 0);
        } else {
            norm = (float)n;
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
        for (int c0 = 0; c0 < nnz; c0 += 32) {
            const int cidx = c0 + lane;
            const bool ok = cidx < nnz;
            int r = 1, y = 0;
            float acc = 0.0f;
            if (ok) {
                const int2 cell = c_ij[cidx];
                y = cell.x;
                const float v0 = __int_as_float(cell.y);
                acc = (a.flavour == 0) ? v0 : __fadd_rn(v0, v0);   // MSA.cpp:1211-1213 doubles the matrix first
                int lo = rowhint, hi = t.L1;
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rp_ij[mid] <= cidx) lo = mid; else hi = mid - 1; }
                r = lo;
            }
            for (int k = 0; k < n; ++k) {
                if (k == i || k == j) continue;
                float w = 1.0f;
                if (a.flavour == 0) { w = wk[k]; if (w < 0.0f) continue; }
                const long long sIK = (long long)i * n + k, sKJ = (long long)k * n + j;
                const int* rp_ik = a.in.rp_pool + a.rp_off[sIK];
                const int2* c_ik = a.in.cells + a.in.nz_off[sIK];
                const int* rp_kj = a.in.rp_pool + a.rp_off[sKJ];
                const int2* c_kj = a.in.cells + a.in.nz_off[sKJ];
                if (ok) {
                    const int b = rp_ik[r], e = rp_ik[r + 1];
                    for (int u = b; u < e; ++u) {
                        const int2 xz = c_ik[u];
                        const int b2 = rp_kj[xz.x], e2 = rp_kj[xz.x + 1];
                        for (int q = b2; q < e2; ++q) {
                            const int2 zy = c_kj[q];
                            if (zy.x >= y) {
                                if (zy.x == y) {
                                    const float v1 = __int_as_float(xz.y), v2 = __int_as_float(zy.y);
                                    const float prod = (a.flavour == 0) ? __fmul_rn(__fmul_rn(w, v1), v2) : __fmul_rn(v1, v2);
                                    acc = __fadd_rn(acc, prod);
                                }
                                break;
                            }
                        }
                    }
                }
            }
            acc = __fdiv_rn(acc, norm);
            const bool keep = ok && (acc >= a.cutoff);
            const unsigned km = __ballot_sync(MLP_FULL, keep);
            if (keep && room) {
                const long long d = obase + kept_total + __popc(km & ((1u << lane) - 1u));
                a.out.cells[d] = make_int2(y, __float_as_int(a.flavour == 0 ? dev_quantize_u16(acc) : acc));
                atomicAdd(&orp[r + 1], 1);
            }
            kept_total += __popc(km);
            rowhint = __shfl_sync(MLP_FULL, r, 31);
            if (rowhint < 1) rowhint = 1;
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

cudaError_t relax_launch(const RelaxArgs& a, int grid, cudaStream_t st) {
    k_relax<<<grid, MLP_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}

int relax_max_blocks_per_sm() {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_relax, MLP_BLOCK, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
