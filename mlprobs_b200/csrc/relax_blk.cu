// Consistency transformation, CTA-per-pair version: masked sparse x sparse product with TMA-staged operand slices.
//
// Reference: cpnp MSA::DoRelaxation / Relax / Relax1 (MSA.cpp:1172-1360), QP ConsistencyStage::doRelaxation / relax
// (ConsistencyStage.cpp:133-300).  Only cells of the old pattern of S_xy survive the reference's final mask, so exactly
// those are computed: every thread owns RB_G consecutive cells (r, c) of S_xy and accumulates
//      acc = init * S_xy[r][c] ;  for z ascending: for (q, v1) in row r of S_xz (q ascending): acc += [w_z *] v1 * S_zy[q][c]
// which is the order in which the reference's scatter-adds reach that cell.  Float multiply and add stay separate.
//
// Work decomposition: one CTA (256 threads) per output pair, pairs from an atomic queue.  A pair is cut into bands of up to
// 1024 consecutive cells (a few dozen rows r0..r1 and a column span cmin..cmax).  For every accepted third sequence z
// the CTA needs two CONTIGUOUS pieces of the pooled set: rows r0..r1 of S_xz and rows cmin..cmax of S_yz (both orientations
// of every matrix are stored), plus their row-pointer slices.  One thread issues four cp.async.bulk (TMA) copies per z into a
// double-buffered shared-memory stage, signalled through an mbarrier, one z ahead of the compute.  From the staged S_yz
// rows the CTA builds per-row dense strips W[c][q - first_q(c)] so that the inner loop is one LDS per (cell, q) instead
// of a sorted merge.  Absent entries are +0 and acc + (+0) == acc exactly (acc is never -0: it starts positive and only
// non-negative products are added), so skipping them or adding them gives the reference's bits.
// Slices that do not fit the stage (very long rows, diffuse matrices) take a merge-join straight from global memory.
#include "posterior.cuh"

#define RB_THREADS 256
#define RB_G 4
#define RB_BAND (RB_THREADS * RB_G)
#define RB_CAPA 1536        // staged cells of S_xz per stage
#define RB_CAPB 2048        // staged cells of S_yz per stage
#define RB_RMAX 256         // rows of S_xy per band
#define RB_CWMAX 320        // columns of S_xy per band (= rows of S_yz staged)
#define RB_WCAP 6144        // floats in the strip pool
// Cell -> thread mapping inside a band: a warp owns 128 consecutive cells and lane l takes cells l, l+32, l+64, l+96 of them,
// so the 32 cells a warp works on at a time are consecutive (about three rows of S_xy: similar row lengths of S_xz, the row
// walks of the lanes end together).  Round 1 gave every thread 4 consecutive cells: a warp then spanned a dozen rows and ran
// at 19 of 32 lanes.
#define RB_CELL(tid, g) ((((tid) >> 5) << 7) + ((g) << 5) + ((tid) & 31))

struct __align__(16) RbDesc {          // one accepted third sequence for the current band (80 bytes, in global scratch)
    long long a_src, b_src;            // first cell of the staged slices in the cell pool (even index = 16-byte aligned)
    long long rpa_src, rpb_src;        // first int of the staged row-pointer slices (multiple of 4)
    int a_n, b_n, rpa_n, rpb_n;        // cells (even) / ints (multiple of 4) to copy
    int a_base, b_base;                // row-pointer value - base = index inside the staged cell slice
    int rpa_skew, rpb_skew;            // position of row r0 / cmin inside the staged row-pointer slice
    float w; int k; int fits; int pad;
};

struct __align__(16) RbSmem {
    int2 A[2][RB_CAPA];
    int2 B[2][RB_CAPB];
    float W[RB_WCAP];
    int4 meta[RB_CWMAX];               // per staged row of S_yz: first q, strip offset, span, 1 = no strip (merge instead)
    int rpA[2][RB_RMAX + 8];
    int rpB[2][RB_CWMAX + 8];
    unsigned long long bar[2];
    int red[RB_THREADS / 32][4];
    int scan[RB_THREADS / 32];
    int wtop, task, nk, room;
    float norm;
    long long obase;
};

__device__ __forceinline__ unsigned rb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rb_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(rb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool rb_mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(rb_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ float rb_merge(const int2* pa, const int2* ea, const int2* pb, const int2* eb, float acc, float w, bool weighted) {
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

template <bool WEIGHTED>
__global__ void __launch_bounds__(RB_THREADS, 2) k_relax_blk(RelaxArgs a) {
    extern __shared__ __align__(16) unsigned char rb_raw[];
    RbSmem& sm = *reinterpret_cast<RbSmem*>(rb_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const int npad = (n + 3) & ~3;
    float* wk = a.wk_scratch + (long long)blockIdx.x * a.wk_stride;   // [npad] weight of the m-th accepted z
    int* kl = reinterpret_cast<int*>(wk + npad);                      // [npad] index  of the m-th accepted z
    RbDesc* desc = reinterpret_cast<RbDesc*>(kl + npad);              // [n]    slice descriptors of the current band
    constexpr bool weighted = WEIGHTED;              // QuickProbs (weights, selectivity) vs cpnp (all z, unweighted): compile-time
    if (tid == 0) {
        rb_mbar_init(&sm.bar[0], 1);
        rb_mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned phase[2] = {0u, 0u};
    for (;;) {
        __syncthreads();
        if (tid == 0) sm.task = atomicAdd(a.counter, 1);
        __syncthreads();
        const int ti = sm.task;
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        const int i = t.a, j = t.b;
        const long long sIJ = (long long)i * n + j;
        const int* rp_ij = a.in.rp_pool + a.rp_off[sIJ];
        const int2* c_ij = a.in.cells + a.in.nz_off[sIJ];
        const int nnz = a.in.nz_cnt[sIJ];
        int* orp = a.out.rp_pool + a.rp_off[sIJ];
        for (int r = tid; r <= t.L1 + 1; r += RB_THREADS) orp[r] = 0;

        // ---- accepted third sequences and the normaliser (warp 0; ordered list, sequential float sum as the reference)
        if (warp == 0) {
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
            if (lane == 0) {
                sm.nk = nk;
                sm.norm = norm;
                const long long ob = (long long)atomicAdd(a.out.cursor, (unsigned long long)nnz);
                a.out.nz_off[sIJ] = ob;
                sm.obase = ob;
                sm.room = (ob + nnz <= a.out.cap) ? 1 : 0;
                if (ob + nnz > a.out.cap) atomicOr(a.err, 2);
                sm.wtop = 0;
            }
        }
        __syncthreads();
        const int nk = sm.nk;
        const float norm = sm.norm;
        const long long obase = sm.obase;
        const bool room = sm.room != 0;
        int kept_total = 0;
        int rowhint = 1;

        for (int c0 = 0; c0 < nnz;) {
            int nb = min(RB_BAND, nnz - c0);
            int rr[RB_G], cc[RB_G];
            float acc[RB_G];
#pragma unroll
            for (int g = 0; g < RB_G; ++g) {
                const int cidx = c0 + RB_CELL(tid, g);
                rr[g] = 1; cc[g] = 0; acc[g] = 0.0f;
                if (cidx < c0 + nb) {
                    const int2 cell = c_ij[cidx];
                    cc[g] = cell.x;
                    const float v0 = __int_as_float(cell.y);
                    acc[g] = weighted ? v0 : __fadd_rn(v0, v0);   // MSA.cpp:1211-1213 doubles the matrix first
                    int lo = (g == 0) ? rowhint : rr[g - 1], hi = t.L1;
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rp_ij[mid] <= cidx) lo = mid; else hi = mid - 1; }
                    rr[g] = lo;
                }
            }
            // ---- extents of the band; halve it until the row / column spans fit the stage
            int rmin, rmax, cmin, cmax;
            for (;;) {
                int v0 = 0x7fffffff, v1 = 0, v2 = 0x7fffffff, v3 = 0;
#pragma unroll
                for (int g = 0; g < RB_G; ++g)
                    if (RB_CELL(tid, g) < nb) { v0 = min(v0, rr[g]); v1 = max(v1, rr[g]); v2 = min(v2, cc[g]); v3 = max(v3, cc[g]); }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    v0 = min(v0, __shfl_xor_sync(MLP_FULL, v0, d)); v1 = max(v1, __shfl_xor_sync(MLP_FULL, v1, d));
                    v2 = min(v2, __shfl_xor_sync(MLP_FULL, v2, d)); v3 = max(v3, __shfl_xor_sync(MLP_FULL, v3, d));
                }
                __syncthreads();
                if (lane == 0) { sm.red[warp][0] = v0; sm.red[warp][1] = v1; sm.red[warp][2] = v2; sm.red[warp][3] = v3; }
                __syncthreads();
                rmin = sm.red[0][0]; rmax = sm.red[0][1]; cmin = sm.red[0][2]; cmax = sm.red[0][3];
#pragma unroll
                for (int w8 = 1; w8 < RB_THREADS / 32; ++w8) {
                    rmin = min(rmin, sm.red[w8][0]); rmax = max(rmax, sm.red[w8][1]);
                    cmin = min(cmin, sm.red[w8][2]); cmax = max(cmax, sm.red[w8][3]);
                }
                if ((rmax - rmin + 1 <= RB_RMAX && cmax - cmin + 1 <= RB_CWMAX) || nb <= 32) break;
                nb >>= 1;
            }
            const int R = rmax - rmin + 1, CW = cmax - cmin + 1;
            const bool band_fits = (R <= RB_RMAX && CW <= RB_CWMAX);
            bool okc[RB_G];
#pragma unroll
            for (int g = 0; g < RB_G; ++g) okc[g] = RB_CELL(tid, g) < nb;

            // ---- slice descriptors of every accepted z for this band (one thread per z: the dependent loads overlap)
            for (int m = tid; m < nk; m += RB_THREADS) {
                const int k = kl[m];
                const long long sIK = (long long)i * n + k, sJK = (long long)j * n + k;
                const long long rpoA = a.rp_off[sIK], rpoB = a.rp_off[sJK];
                RbDesc d;
                d.k = k;
                d.w = weighted ? wk[m] : 1.0f;
                d.pad = 0;
                d.fits = 0;
                d.a_src = d.b_src = d.rpa_src = d.rpb_src = 0;
                d.a_n = d.b_n = d.rpa_n = d.rpb_n = 0;
                d.a_base = d.b_base = d.rpa_skew = d.rpb_skew = 0;
                if (band_fits) {
                    const int a_lo = a.in.rp_pool[rpoA + rmin], a_hi = a.in.rp_pool[rpoA + rmax + 1];
                    const int b_lo = a.in.rp_pool[rpoB + cmin], b_hi = a.in.rp_pool[rpoB + cmax + 1];
                    const long long ga = a.in.nz_off[sIK] + a_lo, gb = a.in.nz_off[sJK] + b_lo;
                    const int ska = (int)(ga & 1), skb = (int)(gb & 1);
                    d.a_src = ga - ska; d.b_src = gb - skb;
                    d.a_n = (ska + (a_hi - a_lo) + 1) & ~1;
                    d.b_n = (skb + (b_hi - b_lo) + 1) & ~1;
                    d.a_base = a_lo - ska; d.b_base = b_lo - skb;
                    const long long gra = rpoA + rmin, grb = rpoB + cmin;
                    d.rpa_skew = (int)(gra & 3); d.rpb_skew = (int)(grb & 3);
                    d.rpa_src = gra - d.rpa_skew; d.rpb_src = grb - d.rpb_skew;
                    d.rpa_n = (d.rpa_skew + R + 1 + 3) & ~3;
                    d.rpb_n = (d.rpb_skew + CW + 1 + 3) & ~3;
                    d.fits = (d.a_n <= RB_CAPA && d.b_n <= RB_CAPB) ? 1 : 0;
                }
                desc[m] = d;
            }
            __syncthreads();

            auto issue = [&](int m) {       // thread 0 only: TMA copies of z = kl[m] into stage m & 1
                const RbDesc d = desc[m];
                if (!d.fits) return;
                const int s = m & 1;
                const unsigned bytes = (unsigned)(d.a_n + d.b_n) * 8u + (unsigned)(d.rpa_n + d.rpb_n) * 4u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rb_mbar_expect_tx(&sm.bar[s], bytes);
                if (d.a_n) rb_bulk_g2s(sm.A[s], a.in.cells + d.a_src, (unsigned)d.a_n * 8u, &sm.bar[s]);
                if (d.b_n) rb_bulk_g2s(sm.B[s], a.in.cells + d.b_src, (unsigned)d.b_n * 8u, &sm.bar[s]);
                rb_bulk_g2s(sm.rpA[s], a.in.rp_pool + d.rpa_src, (unsigned)d.rpa_n * 4u, &sm.bar[s]);
                rb_bulk_g2s(sm.rpB[s], a.in.rp_pool + d.rpb_src, (unsigned)d.rpb_n * 4u, &sm.bar[s]);
            };
            if (tid == 0 && nk > 0) issue(0);

            for (int m = 0; m < nk; ++m) {
                const int s = m & 1;
                const int4 dq = reinterpret_cast<const int4*>(desc + m)[3];   // a_base, b_base, rpa_skew, rpb_skew
                const int4 dw = reinterpret_cast<const int4*>(desc + m)[4];   // w, k, fits, pad
                const float w = __int_as_float(dw.x);
                if (tid == 0 && m + 1 < nk) issue(m + 1);
                if (dw.z) {
                    {   // wait for the stage (bounded spin: a lost copy must not hang the device)
                        unsigned spins = 0;
                        while (!rb_mbar_try_wait(&sm.bar[s], phase[s])) { if (++spins > (1u << 24)) { atomicOr(a.err, 4); break; } }
                        phase[s] ^= 1u;
                    }
                    const int* rpB = sm.rpB[s] + dq.w;
                    const int2* Bs = sm.B[s];
                    for (int x = tid; x < CW; x += RB_THREADS) {                // strips: one thread per staged row of S_yz
                        const int b = rpB[x] - dq.y, e = rpB[x + 1] - dq.y;
                        int4 me = make_int4(0, 0, 0, 0);
                        if (e > b) {
                            const int first = Bs[b].x, span = Bs[e - 1].x - first + 1;
                            const int off = (span <= a.wide_span) ? atomicAdd(&sm.wtop, span) : RB_WCAP;
                            if (off + span <= RB_WCAP) {
                                float* strip = sm.W + off;
                                for (int q = 0; q < span; ++q) strip[q] = 0.0f;
                                for (int y = b; y < e; ++y) { const int2 cell = Bs[y]; strip[cell.x - first] = __int_as_float(cell.y); }
                                me = make_int4(first, off, span, 0);
                            } else me = make_int4(first, b, e, 1);
                        }
                        sm.meta[x] = me;
                    }
                    __syncthreads();
                    if (tid == 0) sm.wtop = 0;
                    const int* rpA = sm.rpA[s] + dq.z;
                    const int2* As = sm.A[s];
#pragma unroll
                    for (int g = 0; g < RB_G; ++g) {
                        if (!okc[g]) continue;
                        const int ra = rr[g] - rmin;
                        const int2* pa = As + (rpA[ra] - dq.x);
                        const int2* ea = As + (rpA[ra + 1] - dq.x);
                        const int4 me = sm.meta[cc[g] - cmin];
                        float ac = acc[g];
                        if (me.w == 0) {
                            const float* strip = sm.W + me.y - me.x;
                            const unsigned span = (unsigned)me.z;
                            for (; pa < ea; ++pa) {
                                const int2 e = *pa;
                                if ((unsigned)(e.x - me.x) < span) {
                                    const float v1 = __int_as_float(e.y), v2 = strip[e.x];
                                    const float prod = weighted ? __fmul_rn(__fmul_rn(w, v1), v2) : __fmul_rn(v1, v2);   // ConsistencyStage.cpp:294 / MSA.cpp:1316
                                    ac = __fadd_rn(ac, prod);
                                }
                            }
                        } else ac = rb_merge(pa, ea, Bs + me.y, Bs + me.z, ac, w, weighted);
                        acc[g] = ac;
                    }
                    __syncthreads();
                } else {
                    // slices too large for the stage: sorted merge straight from global memory
                    const int k = dw.y;
                    const long long sIK = (long long)i * n + k, sJK = (long long)j * n + k;
                    const int* rp_ik = a.in.rp_pool + a.rp_off[sIK];
                    const int2* c_ik = a.in.cells + a.in.nz_off[sIK];
                    const int* rp_jk = a.in.rp_pool + a.rp_off[sJK];
                    const int2* c_jk = a.in.cells + a.in.nz_off[sJK];
#pragma unroll
                    for (int g = 0; g < RB_G; ++g)
                        if (okc[g])
                            acc[g] = rb_merge(c_ik + rp_ik[rr[g]], c_ik + rp_ik[rr[g] + 1],
                                              c_jk + rp_jk[cc[g]], c_jk + rp_jk[cc[g] + 1], acc[g], w, weighted);
                }
            }

            // ---- normalise, threshold, ordered compaction of the band
            float vv[RB_G];
            bool keep[RB_G];
#pragma unroll
            for (int g = 0; g < RB_G; ++g) {
                vv[g] = __fdiv_rn(acc[g], norm);
                keep[g] = okc[g] && (vv[g] >= a.cutoff);
            }
            // cells of a warp in row-major order: pass g = 0..3, lane ascending inside a pass
            int rank[RB_G], wtot = 0;
#pragma unroll
            for (int g = 0; g < RB_G; ++g) {
                const unsigned bm = __ballot_sync(MLP_FULL, keep[g]);
                rank[g] = wtot + __popc(bm & ((1u << lane) - 1u));
                wtot += __popc(bm);
            }
            __syncthreads();
            if (lane == 0) sm.scan[warp] = wtot;
            __syncthreads();
            int before = 0, band_total = 0;
#pragma unroll
            for (int w8 = 0; w8 < RB_THREADS / 32; ++w8) { const int s8 = sm.scan[w8]; if (w8 < warp) before += s8; band_total += s8; }
            if (room) {
#pragma unroll
                for (int g = 0; g < RB_G; ++g)
                    if (keep[g]) {
                        a.out.cells[obase + kept_total + before + rank[g]] = make_int2(cc[g], __float_as_int(weighted ? dev_quantize_u16(vv[g]) : vv[g]));
                        atomicAdd(&orp[rr[g] + 1], 1);
                    }
            }
            kept_total += band_total;
            rowhint = max(rmax, 1);
            c0 += nb;
        }
        __syncthreads();
        // row counts -> row pointers (warp 0; the counts were accumulated with L2 atomics, read them past L1)
        if (warp == 0) {
            int run = 0;
            for (int base = 1; base <= t.L1; base += 32) {
                const int r = base + lane;
                int inc = (r <= t.L1) ? __ldcg(orp + r + 1) : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(MLP_FULL, inc, d); if (lane >= d) inc += o; }
                if (r <= t.L1) orp[r + 1] = run + inc;
                run += __shfl_sync(MLP_FULL, inc, 31);
            }
            if (lane == 0) a.out.nz_cnt[sIJ] = kept_total;
        }
    }
}

size_t relax_blk_smem() { return sizeof(RbSmem); }
int relax_blk_threads() { return RB_THREADS; }
long long relax_blk_scratch_words(int n) { const long long npad = (n + 3) & ~3; return 2 * npad + (long long)n * (sizeof(RbDesc) / 4); }

cudaError_t relax_blk_launch(const RelaxArgs& a, int grid, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_relax_blk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)relax_blk_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_relax_blk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)relax_blk_smem());
    if (e != cudaSuccess) return e;
    if (a.flavour == 0) k_relax_blk<true><<<grid, RB_THREADS, relax_blk_smem(), st>>>(a);
    else k_relax_blk<false><<<grid, RB_THREADS, relax_blk_smem(), st>>>(a);
    return cudaGetLastError();
}

int relax_blk_max_blocks_per_sm() {
    int nb = 0;
    cudaFuncSetAttribute(k_relax_blk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)relax_blk_smem());
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_relax_blk<true>, RB_THREADS, relax_blk_smem()) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
