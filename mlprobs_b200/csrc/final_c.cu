// Merge + MEA score + threshold to CSR, register-band sweep (sweep_c.cuh), compiled per model mix (MODE):
//   0  QuickProbs default (5-state HMM + partition function, PosteriorStage.cpp:156-196, PackedSparseMatrix.cpp:40-83)
//   1 / 2 / 4  c_p_np_aln -p 0 with one model: 5-state HMM / partition function / local HMM (MSA.cpp:1009-1011: the posterior as it is)
//   7  c_p_np_aln -p 0 with all three models: sqrt(((v5^2 + vp^2) + vl^2) / 3)  (MSA.cpp:997-1001)
// The general version (-p 1's merge order and traceback layer, dense dumps) stays k_final_t in posterior.cu.
// cpnp's merge also covers row 0 / column 0, but neither the MEA recursion nor the sparse matrix reads them (ProbabilisticModel.h:
// 826-836 starts at 1, SparseMatrix.h:69 asserts they are below the cutoff), so they are not computed here.
//   QuickProbs: p(i,j) = sqrt((v5^2 + vp^2) * 0.5),  v5 = EXP(min(0, (F+B)(i,j) - total)),  vp = partition posterior;  row 0 / column 0 are 0
//   MEA score: s(i,j) = max(p + s(i-1,j-1), s(i,j-1), s(i-1,j));  distance = 1 - s(L1,L2) / min(L1,L2)
//   cells with p >= cutoff are kept (quantised to QuickProbs' uint16 fixed point)
// States: 0 = MEA row score, 1 = kept cells so far in this row (exact small integer in a float, travels with the row).
// Row 0 is virtual (all zero).  A column's cutoff is a per-column register: NaN outside 1..L2, so padding can never be kept.
// Kept cells of a row are collected in a lane-local bit mask and staged once per row; EXP / sqrt are skipped for a whole
// warp step when none of its lanes has a posterior that can be non-zero.
#include "posterior.cuh"
#include "sweep_c.cuh"
#ifndef MLP_MINB_FINAL
#define MLP_MINB_FINAL 6   // minimum resident CTAs per SM the register allocation is held to (measured)
#endif

namespace {

__device__ __forceinline__ int next_task_c(const KArgs& a, int lane) {
    int ti = 0;
    if (lane == 0) ti = atomicAdd(a.counter, 1);
    return __shfl_sync(MLP_FULL, ti, 0) + a.task_begin;
}

template <int C, int MODE>
struct FinalQ {
    typedef float T;
    typedef float TIN;
    enum { NS = 2, NIN = (MODE == 0 ? 2 : (MODE == 7 ? 3 : 1)), REV = 0, ROW_LO = 1, USES_S1 = 0 };
    __device__ __forceinline__ int row_residue(int) const { return -1; }
    const float* S5; const float* P; const float* SL; const ExpLut* elut;
    float total5, totalL, cutoff; int L1, L2;
    int* rowcnt; int4* stage; int stage_cap; int* stage_n;
    float cut[C]; bool col0; int cL2;       // cL2: index of column L2 inside this lane's strip, -1 if it is elsewhere
    float pv[C]; unsigned hitmask; float cnt_in;
    float score; bool has_score;
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ float load_in(int k, long long idx) const {
        if (MODE == 0) return k == 0 ? S5[idx] : P[idx];
        if (MODE == 7) return k == 0 ? S5[idx] : (k == 1 ? P[idx] : SL[idx]);
        return MODE == 1 ? S5[idx] : (MODE == 2 ? P[idx] : SL[idx]);
    }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            cut[c] = (j >= 1 && j <= L2) ? cutoff : __int_as_float(0x7fc00000);
        }
        col0 = (jbase == 0);
        cL2 = (L2 >= jbase && L2 < jbase + C) ? (L2 - jbase) : -1;
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = 0.0f; st[1] = 0.0f; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = 0.0f; e[1] = 0.0f; }
    __device__ __forceinline__ void begin_row(int, int) { hitmask = 0u; }
    __device__ __forceinline__ void cell(int c, int, int, long long, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&in)[NIN], T (&nw)[NS]) {
        float p = 0.0f;
        if (MODE == 0) {
            const float x = fminf(0.0f, __fsub_rn(in[0], total5));      // ProbabilisticModel.h:483 / ParallelProbabilisticModel.cpp:262
            const float vp = in[NIN > 1 ? 1 : 0];
            // v5 is exactly 0 for x <= -16 (ScoreType.h EXP) and sqrt((0 + 0) * 0.5) is +0: skip the whole evaluation when no lane of the warp needs it
            if (__any_sync(__activemask(), (x > -16.0f) || (vp != 0.0f))) {
                const float v5 = dev_exp_lut(x, elut);
                // sqrt.rn of 0 (most lanes of a step) would take the compiler's out-of-line special-operand path: feed it 1 and select 0
                const float s2 = __fmul_rn(__fadd_rn(__fmul_rn(v5, v5), __fmul_rn(vp, vp)), 0.5f);   // PosteriorStage.cpp:169-177
                const float rt = __fsqrt_rn(s2 == 0.0f ? 1.0f : s2);
                p = (s2 == 0.0f) ? 0.0f : rt;
            }
        } else if (MODE == 7) {
            const float x5 = fminf(0.0f, __fsub_rn(in[0], total5));
            const float vp = in[NIN > 1 ? 1 : 0];
            const float xl = fminf(0.0f, __fsub_rn(in[NIN > 2 ? 2 : 0], totalL));
            // cpnp's partition posterior has no 0.001 filter (MSAPartProbs.cpp:294-298 is commented out), so vp is non-zero almost
            // everywhere: each EXP is skipped on its own when no lane of the warp can have a non-zero value (x <= -16 gives exactly 0)
            float q5 = 0.0f, ql = 0.0f;
            if (__any_sync(__activemask(), x5 > -16.0f)) { const float v5 = dev_exp_lut(x5, elut); q5 = __fmul_rn(v5, v5); }
            if (__any_sync(__activemask(), xl > -16.0f)) { const float vl = dev_exp_lut(xl, elut); ql = __fmul_rn(vl, vl); }
            // MSA.cpp:997-1001: sqrt(((dbl^2 + glob^2) + loc^2) / 3) in the order the reference adds them: 5-state, partition, local
            const float sq = __fadd_rn(__fadd_rn(q5, __fmul_rn(vp, vp)), ql);
            // Far from the alignment vp is ~1e-20 and sq / 3 is a denormal float: div.rn and sqrt.rn then take their out-of-line
            // special-operand paths, divergently -- 30 % of this kernel's instructions in the first profile.  Such a cell has
            // p < 2^-60; the only reader of p besides the cutoff test is `p + S(i-1,j-1)`, and that sum IS S(i-1,j-1) whenever
            // S(i-1,j-1) >= 2^-30 (p is below half an ulp of it).  There the cell is treated as p = 0 -- same bits everywhere;
            // where the score so far is itself that small (first rows of unalignable ends) the exact value is computed.
            const bool tiny = (sq < 0x1p-122f) && (diag[0] >= 0x1p-30f);
            if (__any_sync(__activemask(), !tiny && sq != 0.0f)) {
                const float s3 = __fdiv_rn(tiny ? 3.0f : sq, 3.0f);
                const float rt = __fsqrt_rn(s3 == 0.0f ? 1.0f : s3);
                p = (tiny || s3 == 0.0f) ? 0.0f : rt;
            }
        } else if (MODE == 2) {
            p = in[0];
        } else {
            const float x = fminf(0.0f, __fsub_rn(in[0], MODE == 1 ? total5 : totalL));
            if (__any_sync(__activemask(), x > -16.0f)) p = dev_exp_lut(x, elut);
        }
        if (c == 0 && col0) p = 0.0f;                                  // column 0 is forced to 0
        if (c == 0) cnt_in = col0 ? 0.0f : carry[1];
        const float sc = fmaxf(fmaxf(__fadd_rn(p, diag[0]), carry[0]), old[0]);
        pv[c] = p;
        if (p >= cut[c]) hitmask |= (1u << c);                         // SparseMatrix.h:89 / PackedSparseMatrix.cpp:68
        nw[0] = sc; nw[1] = 0.0f;
    }
    __device__ __forceinline__ void end_row(int i, int jbase, const T (&band)[C][NS], T (&carry)[NS]) {
        const int nh = __popc(hitmask);
        const float cnt_out = __fadd_rn(cnt_in, (float)nh);
        carry[1] = cnt_out;
        if (hitmask) {
            int k = atomicAdd(stage_n, nh);
            int pos = (int)cnt_in;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if ((hitmask >> c) & 1u) {
                    if (k < stage_cap) stage[k] = make_int4(i, pos, jbase + c, __float_as_int(pv[c]));
                    ++k; ++pos;
                }
        }
        if (cL2 >= 0) {
            rowcnt[i + 1] = (int)cnt_out;
            if (i == L1) {
#pragma unroll
                for (int c = 0; c < C; ++c) if (c == cL2) { has_score = true; score = band[c][0]; }
            }
        }
    }
};

template <int C, int MODE>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_FINAL) k_final_c(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    ExpLut* elut = reinterpret_cast<ExpLut*>(smem);
    exp_lut_fill(elut, threadIdx.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* stage_n = reinterpret_cast<int*>(smem + MLP_FINAL_TABLE_BYTES) + warp * 4;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    int4* stage = a.stage + gw * a.stage_cap;
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = nullptr; cx.s2 = nullptr; cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        const long long slotAB = (long long)t.a * a.n + t.b;
        int* rowptr = a.out.rp_pool + a.rp_off[slotAB];
        if (lane == 0) { *stage_n = 0; rowptr[0] = 0; rowptr[1] = 0; }
        __syncwarp();
        FinalQ<C, MODE> m;
        m.elut = elut; m.S5 = a.layerS5 + t.off; m.P = a.layerP + t.off; m.SL = a.layerSL + t.off;   // layers a mode does not read are never dereferenced
        m.total5 = a.pout[ti].total5; m.totalL = a.pout[ti].totalL; m.cutoff = a.cutoff; m.L1 = t.L1; m.L2 = t.L2;
        m.rowcnt = rowptr; m.stage = stage; m.stage_cap = a.stage_cap; m.stage_n = stage_n;
        m.has_score = false; m.score = 0.0f; m.hitmask = 0u; m.cnt_in = 0.0f;
        run_sweep_c<FinalQ<C, MODE>, C>(m, cx, edge, smem + MLP_FINAL_TABLE_BYTES + 64 + warp * MLP_SWEEP_RING_BYTES(2, 4));
        if (m.has_score) {
            const float dist = __fsub_rn(1.0f, __fdiv_rn(m.score, (float)min(t.L1, t.L2)));   // PosteriorStage.cpp:194
            a.pout[ti].mea = m.score;
            a.dist[(long long)t.a * a.n + t.b] = dist;
            a.dist[(long long)t.b * a.n + t.a] = dist;
        }
        __syncwarp();
        __threadfence_block();
        // exclusive scan of the per-row counts -> row pointers (row i occupies rowptr[i]..rowptr[i+1])
        int run = 0;
        for (int base = 1; base <= t.L1; base += 32) {
            const int i = base + lane;
            const int v = (i <= t.L1) ? rowptr[i + 1] : 0;
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(MLP_FULL, inc, d); if (lane >= d) inc += o; }
            if (i <= t.L1) rowptr[i + 1] = run + inc;
            run += __shfl_sync(MLP_FULL, inc, 31);
        }
        const int nnz = run;
        const int staged = *stage_n;
        long long basecell = 0;
        if (lane == 0) {
            basecell = (long long)atomicAdd(a.out.cursor, (unsigned long long)nnz);
            a.out.nz_off[slotAB] = basecell;
            a.out.nz_cnt[slotAB] = (staged <= a.stage_cap && basecell + nnz <= a.out.cap) ? nnz : 0;   // nothing written -> nothing published
            if (staged > a.stage_cap) atomicOr(a.err, 1);
            if (basecell + nnz > a.out.cap) atomicOr(a.err, 2);
        }
        basecell = __shfl_sync(MLP_FULL, basecell, 0);
        __syncwarp();
        if (staged <= a.stage_cap && basecell + nnz <= a.out.cap) {
            for (int k = lane; k < staged; k += 32) {
                const int4 r = stage[k];
                const long long d = basecell + rowptr[r.x] + r.y;
                const float v = __int_as_float(r.w);
                a.out.cells[d] = make_int2(r.z, __float_as_int(MODE == 0 ? dev_quantize_u16(v) : v));   // only QuickProbs stores uint16 fixed point
            }
        }
        __syncwarp();
    }
}

}  // namespace

template <int MODE> static void (*final_c_pick(int C))(KArgs) {
    switch (C) {
        case 1: return k_final_c<1, MODE>; case 2: return k_final_c<2, MODE>; case 3: return k_final_c<3, MODE>; case 4: return k_final_c<4, MODE>;
        case 5: return k_final_c<5, MODE>; case 6: return k_final_c<6, MODE>; case 7: return k_final_c<7, MODE>; case 8: return k_final_c<8, MODE>;
    }
    return nullptr;
}

// mode: 0 = QuickProbs (5-state + partition), otherwise the model mask of a c_p_np_aln -p 0 run (1, 2, 4 or 7)
void (*final_c_kernel(int C, int mode))(KArgs) {
    switch (mode) {
        case 0: return final_c_pick<0>(C); case 1: return final_c_pick<1>(C); case 2: return final_c_pick<2>(C);
        case 4: return final_c_pick<4>(C); case 7: return final_c_pick<7>(C);
    }
    return nullptr;
}
