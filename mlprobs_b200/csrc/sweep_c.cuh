// Striped wavefront skeleton, register-band version (round 2).  Same decomposition and the same dense "slot" layout as
// sweep.cuh (one warp = one pair, lane l owns C consecutive columns of a 32*C-column block and runs one row behind lane
// l-1 / l+1), but C is a COMPILE-TIME constant:
//   * the previous row of the lane's strip lives in registers (band[c][state]) instead of a shared-memory band
//     (no LDS/STS per cell, no per-warp shared memory at all besides the CTA's tables);
//   * per-column constants (residue of the column, terminal-gap factors, validity) are set up once per column block by the
//     model (begin_block) and are plain registers in the unrolled row body, so interior cells carry no boundary tests;
//   * dense inputs of the NEXT step are fetched with ordinary loads into registers one step ahead (the C loads of a step
//     are independent and stay in flight while the current row is computed) -- no cp.async staging, no wait per step;
//   * rows below M::ROW_LO are virtual: their values come from band_init and they are never computed or stored.
//
// Model concept (all members inlined; the model itself is templated on C):
//   typedef T (state type), TIN (dense input type); enum { NS, NIN, REV, ROW_LO, KEEP /* bit s: state s is read from the
//   previous row (kept in the band) */ };
//   void begin_block(int cb, int cbi, int jbase)        per-column constants for columns jbase .. jbase+C-1
//   void band_init(T (&st)[NS], int j)                  virtual row before the first computed one (ROW_LO-1 / L1+1)
//   void edge_init(T (&e)[NS], int i)                   virtual column before the first one (-1 / 32*C*nb)
//   int  row_residue(int i)                             0-based index into the row sequence of the residue row i needs (-1: none)
//   void begin_row(int i, int r1)                       per-row setup; r1 = that residue (fetched one step ahead), enum USES_S1
//                                                       (USES_S1 == 2: bits 8..15 of r1 carry the residue fetched for row i-1 as well)
//   TIN  load_in(int k, long long idx)                  dense input layer k at element idx of this pair's layer
//   void cell<c>(i, j, idx, old, carry, diag, in, nw)   one cell; idx = element index of the cell in the pair's layers
//   void end_row(int i, int jbase, const T (&band)[C][NS], T (&carry)[NS])   after the row's cells; may adjust what the next lane receives
//   void step_sync()                                    called by all lanes at the top of every step
#pragma once
#include "dev_common.cuh"
#define MLP_SWEEP_RING_BYTES(NS, TSIZE) (128 * (NS) * (TSIZE) + 128)

struct SweepCtx2 {
    const uint8_t* s1;     // residues (letter - 'A') of the row sequence, 0-based
    const uint8_t* s2;     // residues of the column sequence
    int lane, L1, L2, nb;
};

template <class M, int C>
__device__ __forceinline__ void run_sweep_c(M& m, const SweepCtx2& cx, typename M::T* edgebuf /* [(L1+1)][NS] per warp, nb > 1 only */,
                                            unsigned char* wsm /* per-warp shared memory, MLP_SWEEP_RING_BYTES(NS, sizeof(T)) */) {
    typedef typename M::T T;
    typedef typename M::TIN TIN;
    constexpr int NS = M::NS, NIN = M::NIN, ROW_LO = M::ROW_LO;
    constexpr bool REV = M::REV != 0;
    const int lane = cx.lane;
    const int src = REV ? (lane + 1) : (lane - 1);
    const bool first_lane = REV ? (lane == 31) : (lane == 0);
    const bool last_lane = REV ? (lane == 0) : (lane == 31);
    const int T_slots = cx.L1 + 32;                    // slots per column block in the dense layout (row + lane)
    const int steps = cx.L1 + 1 - ROW_LO + 31;

    for (int cbi = 0; cbi < cx.nb; ++cbi) {
        const int cb = REV ? (cx.nb - 1 - cbi) : cbi;
        const int jbase = cb * 32 * C + lane * C;
        const bool lane_has_cols = (jbase <= cx.L2 + 1);   // column L2+1 is the virtual column of the reverse sweeps
        m.begin_block(cb, cbi, jbase);
        T band[C][NS];
#pragma unroll
        for (int c = 0; c < C; ++c) m.band_init(band[c], jbase + c);
        // Lanes without columns never compute: what their neighbour reads from them must be the model's "nothing flows in"
        // value (LOG_ZERO for the log-space models), exactly what the edge supplies.
        T myout[NS], diag_in[NS];
        m.edge_init(myout, -1);
#pragma unroll
        for (int s = 0; s < NS; ++s) diag_in[s] = myout[s];
        // element index of (slot, c = 0, this lane) inside the pair's layers: ((cb*T + slot)*C + c)*32 + lane
        const long long blk0 = (long long)cb * T_slots * (C * 32) + lane;
        // The hand-off column of the previous column block (edgebuf, global memory) and the residue of the row are needed at the
        // very start of a step, and by then the store stream of the sweep has pushed them out of L2 (round-2 profile: these
        // dependent loads were 20-36 % of all stall samples).  They go through a per-warp shared-memory ring of 128 rows that
        // the warp refills 32 rows at a time, 32 steps before the first of them is needed: lane l fetches the row the head
        // lane reaches l steps into the chunk.  A row's slot (row & 127) is rewritten only after every lane has left it.
        T* ering = reinterpret_cast<T*>(wsm);
        uint8_t* rring = wsm + 128 * NS * sizeof(T);
        auto fill = [&](int t0) {          // rows the head lane visits at steps t0 .. t0+31
            const int r = REV ? (cx.L1 - (t0 + lane)) : (ROW_LO + t0 + lane);
            if (r >= ROW_LO && r <= cx.L1) {
                if (cbi > 0) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) ering[(r & 127) * NS + s] = edgebuf[(long long)r * NS + s];
                }
                if (M::USES_S1) {
                    const int ri = m.row_residue(r);
                    rring[r & 127] = (ri >= 0 && ri < cx.L1) ? cx.s1[ri] : (uint8_t)0;
                }
            }
        };
        __syncwarp();
        fill(0); fill(32);
        __syncwarp();
        TIN nxt[NIN > 0 ? NIN : 1][C];
        if (NIN > 0) {
            const int i0 = REV ? (cx.L1 + 31 - lane) : (ROW_LO - lane);
            const int slot0 = REV ? (cx.L1 + 31) : ROW_LO;
            if (lane_has_cols && i0 >= ROW_LO && i0 <= cx.L1) {
#pragma unroll
                for (int k = 0; k < NIN; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c) nxt[k][c] = m.load_in(k, blk0 + (long long)slot0 * (C * 32) + c * 32);
            }
        }

        for (int t = 0; t < steps; ++t) {
            const int i = REV ? (cx.L1 - t + 31 - lane) : (ROW_LO + t - lane);
            const int slot = REV ? (cx.L1 + 31 - t) : (ROW_LO + t);
            const bool in_rows = (i >= ROW_LO && i <= cx.L1);
            const bool active = in_rows && lane_has_cols;
            m.step_sync();
            if ((t & 31) == 0 && t > 0) { fill(t + 32); __syncwarp(); }   // rows of steps t+32 .. t+63 (t .. t+31 were fetched 32 steps ago)
            TIN cur[NIN > 0 ? NIN : 1][C];
            if (NIN > 0) {
#pragma unroll
                for (int k = 0; k < NIN; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c) cur[k][c] = nxt[k][c];
                const int inext = REV ? (i - 1) : (i + 1);
                const int snext = REV ? (slot - 1) : (slot + 1);
                if (lane_has_cols && inext >= ROW_LO && inext <= cx.L1) {
#pragma unroll
                    for (int k = 0; k < NIN; ++k)
#pragma unroll
                        for (int c = 0; c < C; ++c) nxt[k][c] = m.load_in(k, blk0 + (long long)snext * (C * 32) + c * 32);
                }
            }
            T in[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) in[s] = __shfl_sync(MLP_FULL, myout[s], src);
            if (first_lane && in_rows) {
                if (cbi == 0) m.edge_init(in, i);
                else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) in[s] = ering[(i & 127) * NS + s];
                }
            }
            if (active) {
                T carry[NS], diag[NS];
                const bool first_row = REV ? (i == cx.L1) : (i == ROW_LO);
                if (first_row) m.band_init(diag, REV ? (jbase + C) : (jbase - 1));   // diagonal predecessor in the virtual row
                else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) diag[s] = diag_in[s];
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) carry[s] = in[s];
                // USES_S1 == 2: the model also wants the residue the ring holds for row i-1 (bits 8..15; only meaningful for i > ROW_LO)
                m.begin_row(i, M::USES_S1 == 2 ? ((int)rring[i & 127] | ((int)rring[(i - 1) & 127] << 8)) : (M::USES_S1 ? (int)rring[i & 127] : 0));
                const long long idx0 = blk0 + (long long)slot * (C * 32);
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    const int c = REV ? (C - 1 - cc) : cc;
                    T nw[NS];
                    TIN inp[NIN > 0 ? NIN : 1];
#pragma unroll
                    for (int k = 0; k < (NIN > 0 ? NIN : 1); ++k) inp[k] = (NIN > 0) ? cur[k][c] : (TIN)0;
                    m.cell(c, i, jbase + c, idx0 + c * 32, band[c], carry, diag, inp, nw);
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        diag[s] = band[c][s];
                        band[c][s] = nw[s];
                        carry[s] = nw[s];
                    }
                }
                m.end_row(i, jbase, band, carry);
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
