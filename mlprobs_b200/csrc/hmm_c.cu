// 5-state pair-HMM forward / backward sweeps (cpnp ProbabilisticModel.h:153-493 flag=true, QP ParallelProbabilisticModel.cpp:40-269),
// register-band version (sweep_c.cuh).  State order (reference numbering): 0 = M, 1 = X1, 2 = Y1, 3 = X2, 4 = Y2.
// Same arithmetic, same order of every LOG_ADD as k_hmm_fwd / k_hmm_bwd of round 1; what changed is the skeleton:
// the previous row lives in registers, the column's insert score and match-table offset are per-column registers, the
// initialisation cells (i <= 1, j <= 1) and the final cell are handled under a per-row flag, and the virtual column L2+1 of
// the backward sweep needs no test at all (LOG_ZERO + anything finite is LOG_ZERO again in float, LOG_ADD of two
// LOG_ZEROs is LOG_ZERO: the padding stays LOG_ZERO by itself).
#include "posterior.cuh"
#include "sweep_c.cuh"
// minimum resident CTAs per SM the register allocation is held to (measured at 1000 x 300: the forward sweep wants its 128
// registers, the backward sweep runs best at 96)
#ifndef MLP_MINB_HMM_FWD
#define MLP_MINB_HMM_FWD 4
#endif
#ifndef MLP_MINB_HMM_BWD
#define MLP_MINB_HMM_BWD 5
#endif

__constant__ DevScalars c_sc_hmm;

cudaError_t hmm_c_set_scalars(const DevScalars& s, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_sc_hmm, &s, sizeof(DevScalars), 0, cudaMemcpyHostToDevice, st);
}

namespace {

__device__ __forceinline__ int next_task_c(const KArgs& a, int lane) {
    int ti = 0;
    if (lane == 0) ti = atomicAdd(a.counter, 1);
    return __shfl_sync(MLP_FULL, ti, 0) + a.task_begin;
}

__device__ __forceinline__ void load_hmm_tables_c(unsigned char* smem, const KArgs& a, float*& match, float*& ins, LogAddLut*& lut) {
    match = reinterpret_cast<float*>(smem);
    ins = match + 676;
    lut = reinterpret_cast<LogAddLut*>(smem + 2816);
    for (int k = threadIdx.x; k < 676; k += blockDim.x) match[k] = a.match[k];
    for (int k = threadIdx.x; k < 26; k += blockDim.x) ins[k] = a.ins[k];
    log_add_lut_fill(lut, threadIdx.x);
    __syncthreads();
}

template <int C>
struct HmmFwdC {
    typedef float T;
    typedef float TIN;
    enum { NS = 5, NIN = 0, REV = 0, ROW_LO = 0, USES_S1 = 1 };
    __device__ __forceinline__ int row_residue(int i) const { return i - 1; }
    const float* match; const float* ins; unsigned lutb;
    float* F; const uint8_t* s1; const uint8_t* s2; int L1, L2;
    float t0q[5], tqq[5], tq0[5];
    int r2[C]; float ins2[C];
    float ins1; const float* mrow; bool init_row;
    float fin[5]; bool has_fin;
    __device__ __forceinline__ void load_consts() {
#pragma unroll
        for (int q = 0; q < 5; ++q) { t0q[q] = c_sc_hmm.t0q[q]; tqq[q] = c_sc_hmm.tqq[q]; tq0[q] = c_sc_hmm.tq0[q]; }
    }
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ float load_in(int, long long) const { return 0.0f; }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            r2[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0;
            ins2[c] = ins[r2[c]];
        }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        ins1 = ins[r1]; mrow = match + r1 * 26;
        init_row = (i <= 1);
    }
    __device__ __forceinline__ void cell(int c, int i, int j, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&)[1], T (&nw)[NS]) {
        // ProbabilisticModel.h:213-245 / ParallelProbabilisticModel.cpp:91-113
        const float e = mrow[r2[c]];
        float m = __fadd_rn(diag[0], tq0[0]);
        m = dev_log_add_lutb(m, __fadd_rn(diag[1], tq0[1]), lutb);
        m = dev_log_add_lutb(m, __fadd_rn(diag[2], tq0[2]), lutb);
        m = dev_log_add_lutb(m, __fadd_rn(diag[3], tq0[3]), lutb);
        m = dev_log_add_lutb(m, __fadd_rn(diag[4], tq0[4]), lutb);
        m = __fadd_rn(m, e);
        float x1 = __fadd_rn(ins1, dev_log_add_lutb(__fadd_rn(old[0], t0q[1]), __fadd_rn(old[1], tqq[1]), lutb));
        float x2 = __fadd_rn(ins1, dev_log_add_lutb(__fadd_rn(old[0], t0q[3]), __fadd_rn(old[3], tqq[3]), lutb));
        float y1 = __fadd_rn(ins2[c], dev_log_add_lutb(__fadd_rn(carry[0], t0q[2]), __fadd_rn(carry[2], tqq[2]), lutb));
        float y2 = __fadd_rn(ins2[c], dev_log_add_lutb(__fadd_rn(carry[0], t0q[4]), __fadd_rn(carry[4], tqq[4]), lutb));
        if (init_row && j <= 1) {   // initialisation cells, ProbabilisticModel.h:173-184 (the recurrence is skipped there)
            m = (i == 1 && j == 1) ? __fadd_rn(c_sc_hmm.init[0], e) : MLP_LOG_ZERO;
            x1 = (i == 1 && j == 0) ? __fadd_rn(c_sc_hmm.init[1], ins1) : MLP_LOG_ZERO;
            x2 = (i == 1 && j == 0) ? __fadd_rn(c_sc_hmm.init[3], ins1) : MLP_LOG_ZERO;
            y1 = (i == 0 && j == 1) ? __fadd_rn(c_sc_hmm.init[2], ins2[c]) : MLP_LOG_ZERO;
            y2 = (i == 0 && j == 1) ? __fadd_rn(c_sc_hmm.init[4], ins2[c]) : MLP_LOG_ZERO;
        }
        nw[0] = m; nw[1] = x1; nw[2] = y1; nw[3] = x2; nw[4] = y2;
        F[idx] = m;
    }
    __device__ __forceinline__ void end_row(int i, int jbase, const T (&band)[C][NS], T (&)[NS]) {
        if (i == L1) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (jbase + c == L2) {
                    has_fin = true;
                    fin[0] = band[c][0]; fin[1] = band[c][1]; fin[2] = band[c][2]; fin[3] = band[c][3]; fin[4] = band[c][4];
                }
        }
    }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_HMM_FWD) k_hmm_fwd_c(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables_c(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    HmmFwdC<C> m;
    m.match = match; m.ins = ins; m.lutb = log_add_lut_bias(lut);
    m.load_consts();
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = a.residues + a.seq_off[t.a]; cx.s2 = a.residues + a.seq_off[t.b];
        cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        m.F = a.layerS5 + t.off; m.s1 = cx.s1; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2; m.has_fin = false;
        run_sweep_c<HmmFwdC<C>, C>(m, cx, edge, smem + MLP_HMM_TABLE_BYTES + 128 + warp * MLP_SWEEP_RING_BYTES(5, 4));
        if (m.has_fin) {   // total forward probability, ProbabilisticModel.h:415-419
            float tF = MLP_LOG_ZERO;
#pragma unroll
            for (int k = 0; k < 5; ++k) tF = dev_log_add(tF, __fadd_rn(m.fin[k], c_sc_hmm.init[k]));
            a.pout[ti].tF5 = tF;
        }
    }
}

template <int C>
struct HmmBwdC {
    typedef float T;
    typedef float TIN;
    enum { NS = 5, NIN = 1, REV = 1, ROW_LO = 0, USES_S1 = 1 };
    __device__ __forceinline__ int row_residue(int i) const { return i; }     // residue i+1 of the row sequence
    const float* match; const float* ins; unsigned lutb;
    float* F;      // in: forward M, out: F + B (ProbabilisticModel.h:483 evaluates (F+B)-total)
    float* cap;    // [0]=B_M(1,1) [1]=B_X1(1,0) [2]=B_Y1(0,1) [3]=B_X2(1,0) [4]=B_Y2(0,1)
    const uint8_t* s1; const uint8_t* s2; int L1, L2;
    float t0q[5], tqq[5], tq0[5];
    int r2[C]; float ins2[C];
    float ins1; const float* mrow; bool special_row;
    __device__ __forceinline__ void load_consts() {
#pragma unroll
        for (int q = 0; q < 5; ++q) { t0q[q] = c_sc_hmm.t0q[q]; tqq[q] = c_sc_hmm.tqq[q]; tq0[q] = c_sc_hmm.tq0[q]; }
    }
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ float load_in(int, long long idx) const { return F[idx]; }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;                       // the transition out of (i, j) emits residue j+1 of the column sequence
            r2[c] = (j + 1 >= 1 && j + 1 <= L2) ? s2[j] : 0;
            ins2[c] = ins[r2[c]];
        }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        ins1 = ins[r1]; mrow = match + r1 * 26;
        special_row = (i <= 1) || (i == L1);
    }
    __device__ __forceinline__ void cell(int c, int i, int j, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&in)[1], T (&nw)[NS]) {
        // ProbabilisticModel.h:340-379 / ParallelProbabilisticModel.cpp:196-218, same LOG_PLUS_EQUALS order
        const float pxy = __fadd_rn(diag[0], mrow[r2[c]]);
        float bm = __fadd_rn(pxy, tq0[0]);
        float x1 = __fadd_rn(pxy, tq0[1]);
        float y1 = __fadd_rn(pxy, tq0[2]);
        float x2 = __fadd_rn(pxy, tq0[3]);
        float y2 = __fadd_rn(pxy, tq0[4]);
        const float a1 = __fadd_rn(old[1], ins1);
        bm = dev_log_add_lutb(bm, __fadd_rn(a1, t0q[1]), lutb);
        x1 = dev_log_add_lutb(x1, __fadd_rn(a1, tqq[1]), lutb);
        const float a2 = __fadd_rn(old[3], ins1);
        bm = dev_log_add_lutb(bm, __fadd_rn(a2, t0q[3]), lutb);
        x2 = dev_log_add_lutb(x2, __fadd_rn(a2, tqq[3]), lutb);
        const float b1 = __fadd_rn(carry[2], ins2[c]);
        bm = dev_log_add_lutb(bm, __fadd_rn(b1, t0q[2]), lutb);
        y1 = dev_log_add_lutb(y1, __fadd_rn(b1, tqq[2]), lutb);
        const float b2 = __fadd_rn(carry[4], ins2[c]);
        bm = dev_log_add_lutb(bm, __fadd_rn(b2, t0q[4]), lutb);
        y2 = dev_log_add_lutb(y2, __fadd_rn(b2, tqq[4]), lutb);
        if (special_row) {
            if (i == L1 && j == L2) { bm = c_sc_hmm.init[0]; x1 = c_sc_hmm.init[1]; y1 = c_sc_hmm.init[2]; x2 = c_sc_hmm.init[3]; y2 = c_sc_hmm.init[4]; }
            if (i <= 1 && j <= 1) {
                if (i == 1 && j == 1) cap[0] = bm;
                if (i == 1 && j == 0) { cap[1] = x1; cap[3] = x2; }
                if (i == 0 && j == 1) { cap[2] = y1; cap[4] = y2; }
            }
        }
        nw[0] = bm; nw[1] = x1; nw[2] = y1; nw[3] = x2; nw[4] = y2;
        F[idx] = __fadd_rn(in[0], bm);
    }
    __device__ __forceinline__ void end_row(int, int, const T (&)[C][NS], T (&)[NS]) const {}
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_HMM_BWD) k_hmm_bwd_c(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables_c(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* cap = reinterpret_cast<float*>(smem + MLP_HMM_TABLE_BYTES) + warp * 8;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    HmmBwdC<C> m;
    m.match = match; m.ins = ins; m.lutb = log_add_lut_bias(lut); m.cap = cap;
    m.load_consts();
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = a.residues + a.seq_off[t.a]; cx.s2 = a.residues + a.seq_off[t.b];
        cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        m.F = a.layerS5 + t.off; m.s1 = cx.s1; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2;
        run_sweep_c<HmmBwdC<C>, C>(m, cx, edge, smem + MLP_HMM_TABLE_BYTES + 128 + warp * MLP_SWEEP_RING_BYTES(5, 4));
        __syncwarp();
        if (lane == 0) {   // ProbabilisticModel.h:421-432 / ParallelProbabilisticModel.cpp:226-231, then :453 and PosteriorStage.cpp:142
            const int r1 = cx.s1[0], r2 = cx.s2[0];
            float tB = __fadd_rn(__fadd_rn(c_sc_hmm.init[0], match[r1 * 26 + r2]), cap[0]);
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc_hmm.init[1], ins[r1]), cap[1]));
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc_hmm.init[2], ins[r2]), cap[2]));
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc_hmm.init[3], ins[r1]), cap[3]));
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc_hmm.init[4], ins[r2]), cap[4]));
            float total = __fdiv_rn(__fadd_rn(a.pout[ti].tF5, tB), 2.0f);
            if (a.flavour == 0 && total == 0.0f) total = 1.0f;   // ParallelProbabilisticModel.cpp:252-254
            a.pout[ti].total5 = total;
        }
        __syncwarp();
    }
}

typedef void (*KFn)(KArgs);
template <int C> KFn pick(int kernel) { return kernel == MLP_K_HMM_FWD ? (KFn)k_hmm_fwd_c<C> : (KFn)k_hmm_bwd_c<C>; }

}  // namespace

void (*hmm_c_kernel(int kernel, int C))(KArgs) {
    switch (C) {
        case 1: return pick<1>(kernel); case 2: return pick<2>(kernel); case 3: return pick<3>(kernel); case 4: return pick<4>(kernel);
        case 5: return pick<5>(kernel); case 6: return pick<6>(kernel); case 7: return pick<7>(kernel); case 8: return pick<8>(kernel);
    }
    return nullptr;
}
