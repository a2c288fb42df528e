// Striped wavefront skeleton shared by every pair-DP kernel (pair-HMM forward/backward, partition
// function forward/reverse, merge+MEA+sparsify).
//
// One warp owns one sequence pair.  The (L1+1) x (L2+1) cell grid is cut into column blocks of 32*C columns;
// inside a block lane l owns the C consecutive columns [l*C, l*C+C) and walks the rows, one row per step,
// skewed by its lane index: at step t lane l is on row  t - l  (forward)  or  L1 - t + 31 - l  (reverse).
// Lane l therefore needs, for its first column, the values lane l-1 (forward) / l+1 (reverse) produced one
// step earlier: ONE warp shuffle per carried state per step, no block barrier, no global hand-off.
// The previous row of the strip lives in a per-warp shared-memory band laid out [c][state][lane]
// (lane-contiguous -> conflict free).  Between column blocks the boundary column travels through a small
// per-warp global "edge" array.
//
// Dense layers written for later sweeps use the slot layout [cb][slot][c][lane] with slot = row + lane.
// Forward step t and reverse step (L1 + 31 - t) touch the SAME slot on every lane, so each dense access is a
// fully coalesced 128-byte line per c, in both directions.
#pragma once
#define MLP_SWEEP_MAXC 8   // columns per lane the sweep kernels are unrolled for (ctx.h: kCmaxLimit)
#include "dev_common.cuh"

template <int N, class T>
__device__ __forceinline__ void shfl_vec(T (&dst)[N], const T (&src)[N], int srclane) {
#pragma unroll
    for (int s = 0; s < N; ++s) dst[s] = __shfl_sync(MLP_FULL, src[s], srclane);
}

struct SweepCtx {
    const PairTask* task;
    const uint8_t* s1;     // residues (letter-'A') of the row sequence, 0-based
    const uint8_t* s2;     // residues of the column sequence
    int lane;
    int L1, L2, C, nb, T;  // T = L1 + 32 slots per column block
    long long off;         // dense layer offset of the task
};

// Model concept:
//   typedef T; enum { NS, REV, COLMASK };  // NS states per cell, COLMASK = states kept in the row band
//   void band_init(T (&st)[NS], int j)                  value of the virtual row before the first (row -1 / L1+1)
//   void edge_init(T (&e)[NS], int i)                   value of the virtual column before the first (col -1 / L2+1)
//   void begin_row(int i, int r1)                       per-row setup (r1 = residue the model needs for this row)
//   void cell(int i, int j, int c, int buf, int r2, int slot /* element index inside the pair's dense layer */, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS], T (&nw)[NS])
//   int row_residue_index(int i) / col_residue_index(int j)   1-based residue used at row i / column j (0 = none)
//   void begin_block(int cb, int cbi)                    called at the start of every column block
//   void step_sync()                                    called by ALL lanes at the top of every step (warp-wide reductions)
//   void prefetch(int slotbase, int C, int buf)   issue cp.async of the dense inputs of one wavefront slot into
//                                                       staging buffer `buf` (models without dense inputs: no-op);
//                                                       cell() receives `buf` and reads its inputs from there
// rank of state s among the states kept in the row band (compile-time after unrolling)
template <int COLMASK>
__device__ __forceinline__ constexpr int band_rank(int s) { return __builtin_popcount(COLMASK & ((1 << s) - 1)); }

template <class M>
__device__ __forceinline__ void run_sweep(M& m, const SweepCtx& cx, typename M::T* band /* [NS][Cmax][32] per warp */,
                                          uint8_t* colres /* [Cmax][32] per warp */, int Cmax,
                                          typename M::T* edgebuf /* [(L1+1)][NS] per warp, only if nb>1 */) {
    typedef typename M::T T;
    constexpr int NS = M::NS;
    constexpr int NSB = __builtin_popcount((int)M::COLMASK);   // states kept in the row band
    const int lane = cx.lane;
    const int C = cx.C;
    const int src = M::REV ? (lane + 1) : (lane - 1);
    const bool first_lane = M::REV ? (lane == 31) : (lane == 0);
    const bool last_lane = M::REV ? (lane == 0) : (lane == 31);

    for (int cbi = 0; cbi < cx.nb; ++cbi) {
        const int cb = M::REV ? (cx.nb - 1 - cbi) : cbi;
        const int jbase = cb * 32 * C + lane * C;
        m.begin_block(cb, cbi);
        const bool lane_has_cols = (jbase <= cx.L2 + 1);   // column L2+1 is the virtual column of reverse sweeps
        // stage the strip: boundary row values + residues of my columns
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            T st[NS];
            m.band_init(st, j);
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if ((M::COLMASK >> s) & 1) band[(c * NSB + band_rank<M::COLMASK>(s)) * 32 + lane] = st[s];
            const int rj = m.col_residue_index(j);
            colres[c * 32 + lane] = (rj >= 1 && rj <= cx.L2) ? cx.s2[rj - 1] : (uint8_t)0;
        }
        __syncwarp();
        T myout[NS], diag_in[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) { myout[s] = (T)0; diag_in[s] = (T)0; }
        // software pipeline of the dense inputs: the slot of step t+1 is fetched (cp.async) while step t computes
        {
            const int i0 = M::REV ? (cx.L1 + 31 - lane) : (0 - lane);
            if (lane_has_cols && i0 >= 0 && i0 <= cx.L1)
                m.prefetch(((cb * cx.T + (M::REV ? (cx.L1 + 31) : 0)) * C) * 32 + lane, C, 0);
        }

        for (int t = 0; t < cx.T; ++t) {
            const int i = M::REV ? (cx.L1 - t + 31 - lane) : (t - lane);
            const bool in_rows = (i >= 0 && i <= cx.L1);
            const bool active = in_rows && lane_has_cols;
            m.step_sync();
            if (M::NIN > 0) {
                cp_async_wait_all();
                const int inext = M::REV ? (i - 1) : (i + 1);
                if (lane_has_cols && inext >= 0 && inext <= cx.L1 && t + 1 < cx.T)
                    m.prefetch(((cb * cx.T + (M::REV ? (cx.L1 + 31 - (t + 1)) : (t + 1))) * C) * 32 + lane, C, (t + 1) & 1);
            }
            T in[NS];
            shfl_vec<NS, T>(in, myout, src);
            if (first_lane && in_rows) {
                if (cbi == 0) m.edge_init(in, i);
                else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) in[s] = edgebuf[(long long)i * NS + s];
                }
            }
            if (active) {
                T carry[NS], diag[NS];
                const bool first_row = M::REV ? (i == cx.L1) : (i == 0);
                if (first_row) {
                    // diagonal predecessor of my first column lies in the virtual row: take the band_init value of column jbase-1 / jbase+C
                    m.band_init(diag, M::REV ? (jbase + C) : (jbase - 1));
                } else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) diag[s] = diag_in[s];
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) carry[s] = in[s];
                const int ri = m.row_residue_index(i);
                const int r1 = (ri >= 1 && ri <= cx.L1) ? cx.s1[ri - 1] : 0;
                m.begin_row(i, r1);
                const int slotbase = ((cb * cx.T + (M::REV ? (cx.L1 + 31 - t) : t)) * C) * 32 + lane;   // element index inside this pair's layer
                auto do_cell = [&](const int c) {
                    const int j = jbase + c;
                    T old[NS], nw[NS];
                    T* bcell = band + (c * NSB) * 32 + lane;   // [c][state][lane]: one address per cell, states at immediate offsets
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        if ((M::COLMASK >> s) & 1) old[s] = bcell[band_rank<M::COLMASK>(s) * 32]; else old[s] = (T)0;
                    m.cell(i, j, c, t & 1, colres[c * 32 + lane], slotbase + c * 32, old, carry, diag, nw);
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if ((M::COLMASK >> s) & 1) bcell[band_rank<M::COLMASK>(s) * 32] = nw[s];
                        diag[s] = old[s];
                        carry[s] = nw[s];
                    }
                };
                if (M::UNROLLC) {
                    // fully unrolled over the 8 columns a lane can own (c is a compile-time constant in every copy, so the band,
                    // residue and slot addresses are immediates); copies beyond C are skipped by a warp-uniform branch.  Measured
                    // per kernel at 1000 x 300: helps the partition and final sweeps, not the 5-state ones.
#pragma unroll
                    for (int cc = 0; cc < MLP_SWEEP_MAXC; ++cc) {
                        const int c = M::REV ? (MLP_SWEEP_MAXC - 1 - cc) : cc;
                        if (c < C) do_cell(c);
                    }
                } else {
                    for (int cc = 0; cc < C; ++cc) do_cell(M::REV ? (C - 1 - cc) : cc);
                }
                m.end_row();
#pragma unroll
                for (int s = 0; s < NS; ++s) myout[s] = carry[s];
                if (last_lane && cbi + 1 < cx.nb) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) edgebuf[(long long)i * NS + s] = myout[s];
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) diag_in[s] = in[s];
        }
        __syncwarp();
    }
}
