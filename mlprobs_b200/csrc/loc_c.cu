// cpnp's 3-state local pair-HMM (ProbabilisticModel.h flag=false branches: forward :167-255, backward :339-379, totals :434-451),
// register-band sweeps (sweep_c.cuh).  States: 0 = M, 1 = X, 2 = Y.  Same arithmetic and the same order of every LOG_ADD as
// k_loc_fwd / k_loc_bwd (posterior.cu), which stay as the fallback (MLP_OLD_SWEEP bit 32) and for the dense debug dumps.
//
//   k_loc_fwd_c<C>    forward sweep, writes F_M in slot layout
//   k_loc_bwd_c<C>    backward sweep, F_M -> F_M + B_M in place, and the Z term of every cell (:445-446) in slot layout
//   k_loc_cand_c<C>   forward-direction pass over one of those two layers: the row-major candidate lists of the Z chain
//   k_loc_replay      the sequential LOG_PLUS_EQUALS chain over the candidate lists, one THREAD per pair
//
// The model's total is a sequential row-major LOG_PLUS_EQUALS chain over all (i, j >= 1) cells (:434-451) and has to be replayed in
// that order to get the reference's bits.  Round 1 wrote a row-major copy of the layer from inside the wavefront (32 lanes on 32
// different rows: 32 sectors per store instruction) and replayed it with one WARP per pair (every firing cell costs a warp-wide
// ballot / shuffle / LOG_ADD round: a third of the kernel's instructions).  Here:
//   * a cell v changes the running sum s only if s < v or s - v < 7.5, so a cell that lies 8.5 below any LOWER BOUND of s cannot
//     fire.  The bound has to follow the sum, not the largest cell: thousands of cells of similar size lift the sum ~9 above the
//     largest of them, and a bound from the maximum alone keeps 54 % of the cells of the forward chain (measured) where 9 % fire.
//     So both sweeps carry one more state along the row -- the LOG_ADD of the row's chain terms -- and the lane that finishes the row
//     stores it (rowaux R[i]).  k_loc_cand_c then walks the layer in the forward direction; its head lane accumulates
//     PB(i) = LOG_ADD over R(1..i-1), hands it along the row with the running candidate count, and every lane appends the cells
//     above PB(i) - 8.5 at the right place of the row's list: rows come out in column order without a transposition.
//   * k_loc_replay walks the lists with one thread per pair: 32 independent chains per warp instead of one, no warp-wide rounds.
//     It applies the reference's own firing test to every candidate, so a candidate that does not fire costs only the test.
//   * The arithmetic that produced PB does not matter, because the one assumption -- s >= PB(i) - 0.5 while row i is processed -- is
//     CHECKED by the replay at the start of every row and after every candidate.  By induction a skipped cell then satisfies
//     s - v >= 8 and provably does not fire; if the check ever fails the error word gets bit 16 and the host re-runs the batch with
//     the round-1 kernels, which make no assumption.
#include "posterior.cuh"
#include "sweep_c.cuh"
#ifndef MLP_MINB_LOC_FWD
#define MLP_MINB_LOC_FWD 5
#endif
#ifndef MLP_MINB_LOC_BWD
#define MLP_MINB_LOC_BWD 5
#endif
#ifndef MLP_MINB_LOC_CAND
#define MLP_MINB_LOC_CAND 8
#endif

__constant__ DevScalars c_sc_loc;
__device__ unsigned long long g_loc_dbg[4];   // developer counters: candidates / firing cells of the forward chain, of the backward chain

cudaError_t loc_c_set_scalars(const DevScalars& s, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_sc_loc, &s, sizeof(DevScalars), 0, cudaMemcpyHostToDevice, st);
}

namespace {

// Row sums of the chain terms feed only the BOUND of the Z chain (checked by k_loc_replay), so they need not be the reference's
// LOG_ADD: base-2 log-sum-exp on the special-function unit, 8 instructions instead of 17.
#define MLP_LOG2E 1.4426950408889634f
#define MLP_LN2 0.6931471805599453f
__device__ __forceinline__ float loc_bound_add2(float acc2, float v) {
    const float v2 = __fmul_rn(v, MLP_LOG2E);
    const float mx = fmaxf(acc2, v2), mn = fminf(acc2, v2);
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fsub_rn(mn, mx)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(__fadd_rn(1.0f, e)));
    return __fadd_rn(mx, l);
}

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

// ---------------------------------------------------------------------------------------------------------------- forward
// Row 0 and column 0 are LOG_ZERO (:167-168): X and Y get there by themselves (everything that flows into them is LOG_ZERO,
// LOG_ZERO + finite is LOG_ZERO again in float and LOG_ADD of two LOG_ZEROs is LOG_ZERO), only M has a term without a predecessor and is
// forced.  Cell (1,1) needs no special case either: its three recurrence terms are LOG_ZERO and leave the first term unchanged.
template <int C>
struct LocFwdC {
    typedef float T;
    typedef float TIN;
    enum { NS = 4, NIN = 0, REV = 0, ROW_LO = 0, USES_S1 = 1 };   // state 3: LOG_ADD of the row's cells so far (bound of the Z chain, see above)
    __device__ __forceinline__ int row_residue(int i) const { return i - 1; }
    const float* match; const float* ins; unsigned lutb;
    float* F; float* R; const uint8_t* s2; int L1, L2;
    int r2[C]; float ins2[C]; bool valid[C]; int cL2;
    float ins1; const float* mrow; bool row0, col0;
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ float load_in(int, long long) const { return 0.0f; }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            r2[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0;
            ins2[c] = ins[r2[c]];
            valid[c] = (j >= 1 && j <= L2);
        }
        col0 = (jbase == 0);
        cL2 = (L2 >= jbase && L2 < jbase + C) ? (L2 - jbase) : -1;
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = st[1] = st[2] = st[3] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = e[3] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void begin_row(int i, int r1) { ins1 = ins[r1]; mrow = match + r1 * 26; row0 = (i == 0); }
    __device__ __forceinline__ void cell(int c, int, int, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&)[1], T (&nw)[NS]) {
        // ProbabilisticModel.h:210-211,222-227: base = ((m - a) - b); M = (base - 2r) (+) sum_k ((base + F_k) + lt[k][0]) - 2r
        const float base = __fsub_rn(__fsub_rn(mrow[r2[c]], ins1), ins2[c]);
        float m = __fsub_rn(base, c_sc_loc.r2);
        m = dev_log_add_lutb(m, __fsub_rn(__fadd_rn(__fadd_rn(base, diag[0]), c_sc_loc.lt00), c_sc_loc.r2), lutb);
        m = dev_log_add_lutb(m, __fsub_rn(__fadd_rn(__fadd_rn(base, diag[1]), c_sc_loc.lt10), c_sc_loc.r2), lutb);
        m = dev_log_add_lutb(m, __fsub_rn(__fadd_rn(__fadd_rn(base, diag[2]), c_sc_loc.lt20), c_sc_loc.r2), lutb);
        // :238-241, :252-255
        const float x = dev_log_add_lutb(__fsub_rn(__fadd_rn(old[0], c_sc_loc.lt01), c_sc_loc.r), __fsub_rn(__fadd_rn(old[1], c_sc_loc.lt11), c_sc_loc.r), lutb);
        const float y = dev_log_add_lutb(__fsub_rn(__fadd_rn(carry[0], c_sc_loc.lt02), c_sc_loc.r), __fsub_rn(__fadd_rn(carry[2], c_sc_loc.lt22), c_sc_loc.r), lutb);
        if (row0 || (c == 0 && col0)) m = MLP_LOG_ZERO;
        nw[0] = m; nw[1] = x; nw[2] = y;
        nw[3] = loc_bound_add2(carry[3], valid[c] ? m : MLP_LOG_ZERO);   // a LOG_ZERO term leaves the sum unchanged
        F[idx] = m;
    }
    __device__ __forceinline__ void end_row(int i, int, const T (&band)[C][NS], T (&)[NS]) const {
        if (cL2 >= 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) if (c == cL2) R[i] = __fmul_rn(band[c][3], MLP_LN2);
        }
    }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_LOC_FWD) k_loc_fwd_c(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables_c(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    LocFwdC<C> m;
    m.match = match; m.ins = ins; m.lutb = log_add_lut_bias(lut);
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = a.residues + a.seq_off[t.a]; cx.s2 = a.residues + a.seq_off[t.b];
        cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        m.F = a.layerSL + t.off; m.R = a.rowaux + (long long)ti * a.rowaux_stride; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2;
        run_sweep_c<LocFwdC<C>, C>(m, cx, edge, smem + MLP_HMM_TABLE_BYTES + 128 + warp * MLP_SWEEP_RING_BYTES(4, 4));
    }
}

// --------------------------------------------------------------------------------------------------------------- backward
// B_M starts at LOG_ONE in every cell (:339); the reference's three guarded blocks (i < L1 && j < L2, i < L1, j < L2) are data here:
// the virtual row L1+1 and everything right of column L2 are LOG_ZERO, and a LOG_ZERO term leaves a LOG_ADD's other operand
// unchanged.  For the padding columns to STAY LOG_ZERO their own B_M must not start at LOG_ONE: the start value is a per-column
// register (LOG_ONE for columns 0..L2, LOG_ZERO beyond).
template <int C>
struct LocBwdC {
    typedef float T;
    typedef float TIN;
    enum { NS = 4, NIN = 1, REV = 1, ROW_LO = 0, USES_S1 = 2 };   // state 3: LOG_ADD of the row's Z terms so far (travels right to left)
    __device__ __forceinline__ int row_residue(int i) const { return i; }     // residue i+1 of the row sequence; the ring's entry of row i-1 is residue i
    const float* match; const float* ins; unsigned lutb;
    float* F; float* VB; float* R; const uint8_t* s2; int L1, L2;
    int r2n[C], r2c[C]; float ins2n[C], ins2c[C], bm0[C]; bool valid[C]; int c1;
    float ins1n, ins1c; const float* mrown; const float* mrowc;
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ float load_in(int, long long idx) const { return F[idx]; }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            r2n[c] = (j + 1 >= 1 && j + 1 <= L2) ? s2[j] : 0;          // the transition out of (i, j) emits residue j+1
            r2c[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0;              // the cell's own residue (Z term)
            ins2n[c] = ins[r2n[c]]; ins2c[c] = ins[r2c[c]];
            bm0[c] = (j <= L2) ? 0.0f : MLP_LOG_ZERO;
            valid[c] = (j >= 1 && j <= L2);
        }
        c1 = (1 >= jbase && 1 < jbase + C) ? (1 - jbase) : -1;       // column 1 ends a row of the reverse sweep
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = st[1] = st[2] = st[3] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = e[3] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void begin_row(int i, int rr) {
        const int rn = rr & 255, rc = (i >= 1) ? ((rr >> 8) & 255) : 0;
        ins1n = ins[rn]; mrown = match + rn * 26;
        ins1c = ins[rc]; mrowc = match + rc * 26;
    }
    __device__ __forceinline__ void cell(int c, int, int, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&in)[1], T (&nw)[NS]) {
        // ProbabilisticModel.h:339-379 flag=false, same LOG_PLUS_EQUALS order
        const float pxy = __fsub_rn(__fsub_rn(__fadd_rn(diag[0], mrown[r2n[c]]), ins1n), ins2n[c]);
        float bm = dev_log_add_lutb(bm0[c], __fsub_rn(__fadd_rn(pxy, c_sc_loc.lt00), c_sc_loc.r2), lutb);
        float x = __fsub_rn(__fadd_rn(pxy, c_sc_loc.lt10), c_sc_loc.r2);
        float y = __fsub_rn(__fadd_rn(pxy, c_sc_loc.lt20), c_sc_loc.r2);
        bm = dev_log_add_lutb(bm, __fsub_rn(__fadd_rn(old[1], c_sc_loc.lt01), c_sc_loc.r), lutb);
        x = dev_log_add_lutb(x, __fsub_rn(__fadd_rn(old[1], c_sc_loc.lt11), c_sc_loc.r), lutb);
        bm = dev_log_add_lutb(bm, __fsub_rn(__fadd_rn(carry[2], c_sc_loc.lt02), c_sc_loc.r), lutb);
        y = dev_log_add_lutb(y, __fsub_rn(__fadd_rn(carry[2], c_sc_loc.lt22), c_sc_loc.r), lutb);
        nw[0] = bm; nw[1] = x; nw[2] = y;
        // Z term of this cell, :445-446: (((B_M + m) - a) - b) - 2r with the cell's own residues (read only for i, j >= 1)
        const float vb = __fsub_rn(__fsub_rn(__fsub_rn(__fadd_rn(bm, mrowc[r2c[c]]), ins1c), ins2c[c]), c_sc_loc.r2);
        VB[idx] = vb;
        nw[3] = loc_bound_add2(carry[3], valid[c] ? vb : MLP_LOG_ZERO);
        F[idx] = __fadd_rn(in[0], bm);
    }
    __device__ __forceinline__ void end_row(int i, int, const T (&band)[C][NS], T (&)[NS]) const {
        if (c1 >= 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) if (c == c1) R[i] = __fmul_rn(band[c][3], MLP_LN2);
        }
    }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_LOC_BWD) k_loc_bwd_c(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables_c(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    LocBwdC<C> m;
    m.match = match; m.ins = ins; m.lutb = log_add_lut_bias(lut);
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = a.residues + a.seq_off[t.a]; cx.s2 = a.residues + a.seq_off[t.b];
        cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        m.F = a.layerSL + t.off; m.VB = a.layerVB + t.off; m.R = a.rowaux + (long long)ti * a.rowaux_stride; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2;
        run_sweep_c<LocBwdC<C>, C>(m, cx, edge, smem + MLP_HMM_TABLE_BYTES + 128 + warp * MLP_SWEEP_RING_BYTES(4, 4));
    }
}

// ------------------------------------------------------------------------------------------------------- candidate lists
// States: 0 = PB(i), the bound of the running sum when the chain enters row i (constant along the row), 1 = candidates of this row so
// far (exact small integer in a float); both travel along the row.  Per pair, rowaux holds three arrays of `as` floats: R (row
// sums, from the sweep), PB and the candidate count of every row; row i's candidates are LC[i*(L2+1) + 0 .. count).
template <int C>
struct LocCandC {
    typedef float T;
    typedef float TIN;
    enum { NS = 2, NIN = 1, REV = 0, ROW_LO = 1, USES_S1 = 0 };
    __device__ __forceinline__ int row_residue(int) const { return -1; }
    const float* src; float* LC; const float* R; float* PB; float* CNT; int L1, L2; bool lane0;
    bool valid[C]; int cL2; bool head;
    float pb, rcur;         // head lane only: PB of the current row, R of the current row (fetched one row ahead)
    float* row;
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ float load_in(int, long long idx) const { return src[idx]; }
    __device__ __forceinline__ void begin_block(int, int cbi, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) valid[c] = (jbase + c >= 1 && jbase + c <= L2);
        cL2 = (L2 >= jbase && L2 < jbase + C) ? (L2 - jbase) : -1;
        head = lane0 && cbi == 0;
        pb = MLP_LOG_ZERO;
        rcur = head ? R[1] : 0.0f;
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = MLP_LOG_ZERO; st[1] = 0.0f; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = pb; e[1] = 0.0f; }   // head lane, first column block: the row starts here
    __device__ __forceinline__ void begin_row(int i, int) { row = LC + (long long)i * (L2 + 1); }
    __device__ __forceinline__ void cell(int c, int, int, long long, const T (&)[NS], const T (&carry)[NS], const T (&)[NS],
                                         const TIN (&in)[1], T (&nw)[NS]) {
        float cnt = carry[1];
        const float v = in[0];
        if (valid[c] && v > __fadd_rn(carry[0], -8.5f)) { row[(int)cnt] = v; cnt = __fadd_rn(cnt, 1.0f); }
        nw[0] = carry[0]; nw[1] = cnt;
    }
    __device__ __forceinline__ void end_row(int i, int, const T (&band)[C][NS], T (&)[NS]) {
        if (head) {
            PB[i] = pb;
            pb = dev_log_add(pb, rcur);
            if (i + 1 <= L1) rcur = R[i + 1];
        }
        if (cL2 >= 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) if (c == cL2) CNT[i] = band[c][1];
        }
    }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_LOC_CAND) k_loc_cand_c(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    LocCandC<C> m;
    m.lane0 = (lane == 0);
    const long long as = a.rowaux_stride / 3;
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = nullptr; cx.s2 = nullptr; cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        float* aux = a.rowaux + (long long)ti * a.rowaux_stride;
        m.src = (a.loc_phase == 0 ? a.layerSL : a.layerVB) + t.off; m.LC = a.layerLC + t.off; m.L1 = t.L1; m.L2 = t.L2;
        m.R = aux; m.PB = aux + as; m.CNT = aux + 2 * as;
        run_sweep_c<LocCandC<C>, C>(m, cx, edge, smem + 64 + warp * MLP_SWEEP_RING_BYTES(2, 4));
    }
}

// ------------------------------------------------------------------------------------------------------------ the Z chain
// ProbabilisticModel.h:434-451 over the candidate lists, one thread per pair.  phase 0: forward total -> tFL; phase 1: backward total,
// then totalL = (tF + tB) / 2 (:453).  Candidates are fetched eight at a time, the next eight while the current ones are applied (the
// lists of the 32 pairs of a warp lie a multiple of 4 KB apart when the sequences have equal lengths: single loads thrash the L1 sets).
__global__ void __launch_bounds__(128) k_loc_replay(KArgs a) {
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= a.ntasks) return;
    const PairTask t = a.tasks[ti];
    const float* LC = a.layerLC + t.off;
    const long long as = a.rowaux_stride / 3;
    const float* PB = a.rowaux + (long long)ti * a.rowaux_stride + as;
    const float* CNT = PB + as;
    const int W = t.L2 + 1;
    float sum = MLP_LOG_ZERO;
    bool bad = false;
    unsigned ncand = 0, nfire = 0;
    float pbn = PB[1], cn = CNT[1];
    for (int i = 1; i <= t.L1; ++i) {
        const float* row = LC + (long long)i * W;
        const int cnt = (int)cn;
        const float floor_i = __fadd_rn(pbn, -0.5f);      // the running sum must stay above this while row i is processed
        if (i < t.L1) { pbn = PB[i + 1]; cn = CNT[i + 1]; }
        ncand += cnt;
        if (!(sum >= floor_i)) bad = true;
        float nx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) nx[u] = (u < cnt) ? row[u] : MLP_LOG_ZERO;
        for (int k = 0; k < cnt; k += 8) {
            float cv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cv[u] = nx[u];
#pragma unroll
            for (int u = 0; u < 8; ++u) nx[u] = (k + 8 + u < cnt) ? row[k + 8 + u] : MLP_LOG_ZERO;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float v = cv[u];          // padding of the last chunk is LOG_ZERO, which never fires
                // LOG_PLUS_EQUALS leaves the sum alone when the cell is LOG_ZERO or 7.5 below it (ScoreType.h:279-285)
                if (!(sum >= v && (v == MLP_LOG_ZERO || __fsub_rn(sum, v) >= 7.5f))) {
                    sum = dev_log_add(sum, v); ++nfire;
                    if (!(sum >= floor_i)) bad = true;
                }
            }
        }
    }
    if (!(fabsf(sum) < 4.0e6f) && sum != MLP_LOG_ZERO) bad = true;   // the 8.5 margin of the candidate test must survive its own rounding
    if (bad || (a.loc_debug & 2)) atomicOr(a.err, 16);
    if (a.loc_debug & 1) { atomicAdd(&g_loc_dbg[2 * a.loc_phase], (unsigned long long)ncand); atomicAdd(&g_loc_dbg[2 * a.loc_phase + 1], (unsigned long long)nfire); }
    if (a.loc_phase == 0) a.pout[ti].tFL = sum;
    else a.pout[ti].totalL = __fdiv_rn(__fadd_rn(a.pout[ti].tFL, sum), 2.0f);
}

typedef void (*KFn)(KArgs);
template <int C> KFn pick(int kernel) {
    return kernel == MLP_K_LOCAL_FWD ? (KFn)k_loc_fwd_c<C> : (kernel == MLP_K_LOCAL_BWD ? (KFn)k_loc_bwd_c<C> : (KFn)k_loc_cand_c<C>);
}

}  // namespace

void (*loc_c_kernel(int kernel, int C))(KArgs) {
    switch (C) {
        case 1: return pick<1>(kernel); case 2: return pick<2>(kernel); case 3: return pick<3>(kernel); case 4: return pick<4>(kernel);
        case 5: return pick<5>(kernel); case 6: return pick<6>(kernel); case 7: return pick<7>(kernel); case 8: return pick<8>(kernel);
    }
    return nullptr;
}

cudaError_t loc_debug_counters(unsigned long long out[4]) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_loc_dbg, sizeof(unsigned long long) * 4);
    unsigned long long z[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_loc_dbg, z, sizeof(z));
    return e;
}

cudaError_t loc_replay_launch(const KArgs& a, cudaStream_t st) {
    k_loc_replay<<<(a.ntasks + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}
