// Pair-DP kernels of the posterior stage (one warp per sequence pair, striped wavefront: sweep.cuh).
//   k_part_fwd / k_part_rev   partition-function posterior   (QP PartitionFunction.cpp:71-291, cpnp MSAPartProbs.cpp:78-660)
//   k_hmm_fwd  / k_hmm_bwd    5-state pair-HMM               (cpnp ProbabilisticModel.h:153-493 flag=true, QP ParallelProbabilisticModel.cpp:40-269)
//   k_loc_fwd  / k_loc_bwd    3-state local pair-HMM + row-major Z replay (cpnp ProbabilisticModel.h flag=false)
//   k_final                   merge + MEA score + threshold to CSR (PosteriorStage.cpp:156-196, MSA.cpp:992-1023, SparseMatrix.h:55-98)
//   k_transpose               second orientation of each CSR matrix (PackedSparseMatrix.cpp:93-140, SparseMatrix.h:205-248)
#include "posterior.cuh"
#include "sweep.cuh"
#include "sweep_c.cuh"

__constant__ DevScalars c_sc;

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int next_task(int* counter, int lane) {
    int ti = 0;
    if (lane == 0) ti = atomicAdd(counter, 1);
    return __shfl_sync(MLP_FULL, ti, 0);
}

__device__ __forceinline__ SweepCtx make_ctx(const PairTask& t, const KArgs& a, int lane) {
    SweepCtx cx;
    cx.task = &t;
    cx.s1 = a.residues + a.seq_off[t.a];
    cx.s2 = a.residues + a.seq_off[t.b];
    cx.lane = lane;
    cx.L1 = t.L1; cx.L2 = t.L2; cx.C = t.C; cx.nb = t.nb; cx.T = t.L1 + 32;
    cx.off = t.off;
    return cx;
}

// per-warp shared-memory carve-up: [tables (CTA)] [warp0: band | stage | colres | cap] [warp1: ...]
// stage = NIN double-buffered wavefront slots of dense input (cp.async targets), element size sizeof(T)
template <class T, int NSB /* states kept in the band */, int NIN>
__device__ __forceinline__ void warp_smem(unsigned char* base, int tables_bytes, int Cmax, int warp,
                                          T*& band, T*& stage, uint8_t*& colres, float*& cap) {
    const int band_bytes = NSB * Cmax * 32 * (int)sizeof(T);
    const int stage_bytes = NIN * 2 * Cmax * 32 * (int)sizeof(T);
    const int per_warp = band_bytes + stage_bytes + Cmax * 32 + 64;
    unsigned char* p = base + tables_bytes + (size_t)warp * ((per_warp + 15) & ~15);
    band = reinterpret_cast<T*>(p);
    stage = reinterpret_cast<T*>(p + band_bytes);
    colres = p + band_bytes + stage_bytes;
    cap = reinterpret_cast<float*>(p + band_bytes + stage_bytes + Cmax * 32);
}

// HMM kernels: CTA-shared tables = match[676] | ins[26] (+pad to 704 floats) | LogAddLut (256 B)
__device__ __forceinline__ void load_hmm_tables(unsigned char* smem, const KArgs& a, float*& match, float*& ins, LogAddLut*& lut) {
    match = reinterpret_cast<float*>(smem);
    ins = match + 676;
    lut = reinterpret_cast<LogAddLut*>(smem + 2816);
    for (int k = threadIdx.x; k < 676; k += blockDim.x) match[k] = a.match[k];
    for (int k = threadIdx.x; k < 26; k += blockDim.x) ins[k] = a.ins[k];
    log_add_lut_fill(lut, threadIdx.x);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ 5-state HMM
// state order (reference numbering): 0 = M, 1 = X1, 2 = Y1, 3 = X2, 4 = Y2
struct HmmFwd {
    __device__ __forceinline__ void begin_block(int, int) const {}
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ void end_row() const {}
    typedef float T;
    enum { NS = 5, REV = 0, COLMASK = 0x1f, NIN = 0, UNROLLC = 0 };
    const float* match; const float* ins; const LogAddLut* lut; unsigned lutb;
    float* F;
    int L1, L2;
    float ins1; const float* mrow;
    float fin[5]; bool has_fin;
    float t0q[5], tqq[5], tq0[5];   // transitions hoisted out of constant memory into registers
    __device__ __forceinline__ void load_consts() {
#pragma unroll
        for (int q = 0; q < 5; ++q) { t0q[q] = c_sc.t0q[q]; tqq[q] = c_sc.tqq[q]; tq0[q] = c_sc.tq0[q]; }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ int row_residue_index(int i) const { return i; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j; }
    __device__ __forceinline__ void begin_row(int, int r1) { ins1 = ins[r1]; mrow = match + r1 * 26; }
    __device__ __forceinline__ void prefetch(int, int, int) const {}
    __device__ __forceinline__ void cell(int i, int j, int, int, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        // ProbabilisticModel.h:213-245 / ParallelProbabilisticModel.cpp:91-113
        float m = __fadd_rn(diag[0], tq0[0]);
        m = dev_log_add_lutb(m, __fadd_rn(diag[1], tq0[1]), lutb);
        m = dev_log_add_lutb(m, __fadd_rn(diag[2], tq0[2]), lutb);
        m = dev_log_add_lutb(m, __fadd_rn(diag[3], tq0[3]), lutb);
        m = dev_log_add_lutb(m, __fadd_rn(diag[4], tq0[4]), lutb);
        m = __fadd_rn(m, mrow[r2]);
        const float ins2 = ins[r2];
        float x1 = __fadd_rn(ins1, dev_log_add_lutb(__fadd_rn(old[0], t0q[1]), __fadd_rn(old[1], tqq[1]), lutb));
        float x2 = __fadd_rn(ins1, dev_log_add_lutb(__fadd_rn(old[0], t0q[3]), __fadd_rn(old[3], tqq[3]), lutb));
        float y1 = __fadd_rn(ins2, dev_log_add_lutb(__fadd_rn(carry[0], t0q[2]), __fadd_rn(carry[2], tqq[2]), lutb));
        float y2 = __fadd_rn(ins2, dev_log_add_lutb(__fadd_rn(carry[0], t0q[4]), __fadd_rn(carry[4], tqq[4]), lutb));
        if (i <= 1 && j <= 1) {   // initialisation cells, ProbabilisticModel.h:173-184 (the recurrence is skipped there)
            m = (i == 1 && j == 1) ? __fadd_rn(c_sc.init[0], mrow[r2]) : MLP_LOG_ZERO;
            x1 = (i == 1 && j == 0) ? __fadd_rn(c_sc.init[1], ins1) : MLP_LOG_ZERO;
            x2 = (i == 1 && j == 0) ? __fadd_rn(c_sc.init[3], ins1) : MLP_LOG_ZERO;
            y1 = (i == 0 && j == 1) ? __fadd_rn(c_sc.init[2], ins2) : MLP_LOG_ZERO;
            y2 = (i == 0 && j == 1) ? __fadd_rn(c_sc.init[4], ins2) : MLP_LOG_ZERO;
        }
        nw[0] = m; nw[1] = x1; nw[2] = y1; nw[3] = x2; nw[4] = y2;
        F[slot] = m;
        if (i == L1 && j == L2) { has_fin = true; fin[0] = m; fin[1] = x1; fin[2] = y1; fin[3] = x2; fin[4] = y2; }
    }
};

__global__ void __launch_bounds__(MLP_BLOCK) k_hmm_fwd(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* band; float* stage; uint8_t* colres; float* cap;
    warp_smem<float, 5, 0>(smem, MLP_HMM_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    HmmFwd m;
    m.match = match; m.ins = ins; m.lut = lut; m.lutb = log_add_lut_bias(lut); m.F = a.layerS5;
    m.load_consts();
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        m.F = a.layerS5 + t.off; m.L1 = t.L1; m.L2 = t.L2; m.has_fin = false;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        if (m.has_fin) {   // total forward probability, ProbabilisticModel.h:415-419
            float tF = MLP_LOG_ZERO;
#pragma unroll
            for (int k = 0; k < 5; ++k) tF = dev_log_add(tF, __fadd_rn(m.fin[k], c_sc.init[k]));   // once per pair: plain version
            a.pout[ti].tF5 = tF;
        }
    }
}

struct HmmBwd {
    __device__ __forceinline__ void begin_block(int, int) const {}
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ void end_row() const {}
    typedef float T;
    enum { NS = 5, REV = 1, COLMASK = 0x0b, NIN = 1, UNROLLC = 0 };   // keep B_M, X1, X2 of row i+1; Y1, Y2 travel along the row
    const float* match; const float* ins; const LogAddLut* lut; unsigned lutb;
    float* F;      // in: forward M, out: F + B (ProbabilisticModel.h:483 evaluates (F+B)-total)
    float* stage; int Cmax;
    float* cap;    // [0]=B_M(1,1) [1]=B_X1(1,0) [2]=B_Y1(0,1) [3]=B_X2(1,0) [4]=B_Y2(0,1)
    int L1, L2, lane;
    float ins1; const float* mrow;
    float t0q[5], tqq[5], tq0[5];
    __device__ __forceinline__ void load_consts() {
#pragma unroll
        for (int q = 0; q < 5; ++q) { t0q[q] = c_sc.t0q[q]; tqq[q] = c_sc.tqq[q]; tq0[q] = c_sc.tq0[q]; }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = MLP_LOG_ZERO;
    }
    __device__ __forceinline__ int row_residue_index(int i) const { return i + 1; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j + 1; }
    __device__ __forceinline__ void begin_row(int, int r1) { ins1 = ins[r1]; mrow = match + r1 * 26; }
    __device__ __forceinline__ void prefetch(int slotbase, int C, int buf) const {
        for (int c = 0; c < C; ++c) cp_async4(stage + (buf * Cmax + c) * 32 + lane, F + slotbase + c * 32);
    }
    __device__ __forceinline__ void cell(int i, int j, int c, int buf, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        if (j > L2) {   // virtual column L2+1 (and padding): nothing flows in from the right
#pragma unroll
            for (int s = 0; s < NS; ++s) nw[s] = MLP_LOG_ZERO;
            return;
        }
        // ProbabilisticModel.h:340-379 / ParallelProbabilisticModel.cpp:196-218, same LOG_PLUS_EQUALS order
        const float ins2 = ins[r2];
        const float pxy = __fadd_rn(diag[0], mrow[r2]);
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
        const float b1 = __fadd_rn(carry[2], ins2);
        bm = dev_log_add_lutb(bm, __fadd_rn(b1, t0q[2]), lutb);
        y1 = dev_log_add_lutb(y1, __fadd_rn(b1, tqq[2]), lutb);
        const float b2 = __fadd_rn(carry[4], ins2);
        bm = dev_log_add_lutb(bm, __fadd_rn(b2, t0q[4]), lutb);
        y2 = dev_log_add_lutb(y2, __fadd_rn(b2, tqq[4]), lutb);
        if (i == L1 && j == L2) { bm = c_sc.init[0]; x1 = c_sc.init[1]; y1 = c_sc.init[2]; x2 = c_sc.init[3]; y2 = c_sc.init[4]; }
        nw[0] = bm; nw[1] = x1; nw[2] = y1; nw[3] = x2; nw[4] = y2;
        F[slot] = __fadd_rn(stage[(buf * Cmax + c) * 32 + lane], bm);
        if (i <= 1 && j <= 1) {
            if (i == 1 && j == 1) cap[0] = bm;
            if (i == 1 && j == 0) { cap[1] = x1; cap[3] = x2; }
            if (i == 0 && j == 1) { cap[2] = y1; cap[4] = y2; }
        }
    }
};

__global__ void __launch_bounds__(MLP_BLOCK) k_hmm_bwd(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* band; float* stage; uint8_t* colres; float* cap;
    warp_smem<float, 3, 1>(smem, MLP_HMM_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    HmmBwd m;
    m.match = match; m.ins = ins; m.lut = lut; m.lutb = log_add_lut_bias(lut); m.F = a.layerS5; m.cap = cap; m.stage = stage; m.Cmax = a.Cmax; m.lane = lane;
    m.load_consts();
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        m.F = a.layerS5 + t.off; m.L1 = t.L1; m.L2 = t.L2;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        __syncwarp();
        if (lane == 0) {   // ProbabilisticModel.h:421-432 / ParallelProbabilisticModel.cpp:226-231, then :453 and PosteriorStage.cpp:142
            const int r1 = cx.s1[0], r2 = cx.s2[0];
            float tB = __fadd_rn(__fadd_rn(c_sc.init[0], match[r1 * 26 + r2]), cap[0]);
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc.init[1], ins[r1]), cap[1]));
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc.init[2], ins[r2]), cap[2]));
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc.init[3], ins[r1]), cap[3]));
            tB = dev_log_add(tB, __fadd_rn(__fadd_rn(c_sc.init[4], ins[r2]), cap[4]));
            float total = __fdiv_rn(__fadd_rn(a.pout[ti].tF5, tB), 2.0f);
            if (a.flavour == 0 && total == 0.0f) total = 1.0f;   // ParallelProbabilisticModel.cpp:252-254
            a.pout[ti].total5 = total;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ partition function
// states: 0 = Zm, 1 = H (gap run along the row: Ze in QP, Zf in cpnp), 2 = V (gap run down the column)
// The 3-term sums are (Zm+H)+V or (Zm+V)+H depending on which reference statement is reproduced:
//   forward Zm and zz: QP (Zm+H)+V  PartitionFunction.cpp:137,139 ; cpnp (Zm+V)+H  MSAPartProbs.cpp:583,589
//   reverse Zm:        QP (Zm+V)+H  PartitionFunction.cpp:257     ; cpnp (Zm+H)+V  MSAPartProbs.cpp:283
template <bool H_FIRST>
__device__ __forceinline__ double sum3(double zm, double h, double v) {
    return H_FIRST ? __dadd_rn(__dadd_rn(zm, h), v) : __dadd_rn(__dadd_rn(zm, v), h);
}

// Rescaled variant (SC = true, cpnp flavour): the reference runs this recursion in 80-bit long double (range 1e+-4932),
// FP64 would overflow for similar sequences longer than ~600.  Every row i carries a power-of-two scale 2^-e_i: stored
// values are Z * 2^-e_i.  Scaling by powers of two commutes with IEEE rounding, so mantissas are exactly those of an
// unbounded-range double.  e_i is chosen by the lane that starts the row (column 0 forward, column L2+1 reverse) from a
// warp-wide maximum of the exponents seen in the last <= 32 rows and travels with the row (4th carried value, exact small
// integer in a double); it changes only when the magnitude drifted by more than 2^48.  QP's own FP64 code is reproduced
// unscaled (SC = false) because its silent overflow is part of its result.
__device__ __forceinline__ double pow2d(int k) { return __hiloint2double((1023 + k) << 20, 0); }   // |k| <= 1022
__device__ __forceinline__ int exp_of(double x) { return ((__double2hiint(x) >> 20) & 0x7ff) - 1023; }
#define MLP_EXP_NONE (-1000000)

template <bool SC>
struct PartFwdT {
    typedef double T;
    enum { NS = SC ? 4 : 3, REV = 0, COLMASK = 0x7, NIN = 0, UNROLLC = 1 };
    const double* sub; double* Z; int* rowexp; int L1, L2, W; bool qp;   // qp == !SC: QuickProbs order (Zm+H)+V, cpnp (Zm+V)+H
    const double* srow; double zz; bool has_zz; int zexp; double o0, e0;
    // scale bookkeeping (SC): seen_exp = largest true exponent in the row this lane finished last (this column block)
    int seen_exp, row_seen, gmax, e_prev, e_row, cb, cbi; double f;
    __device__ __forceinline__ void reset() { has_zz = false; zz = 0; zexp = 0; seen_exp = MLP_EXP_NONE; row_seen = MLP_EXP_NONE; gmax = MLP_EXP_NONE; e_prev = 0; e_row = 0; f = 1.0; cb = 0; cbi = 0; }
    __device__ __forceinline__ void begin_block(int cb_, int cbi_) { cb = cb_; cbi = cbi_; e_prev = 0; seen_exp = MLP_EXP_NONE; gmax = MLP_EXP_NONE; }
    __device__ __forceinline__ void step_sync() { if (SC) gmax = __reduce_max_sync(MLP_FULL, seen_exp); }
    __device__ __forceinline__ void prefetch(int, int, int) const {}
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = 0;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const {
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = 0;
    }
    __device__ __forceinline__ int row_residue_index(int i) const { return i; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j; }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        srow = sub + r1 * 26; e_row = MLP_EXP_NONE; row_seen = MLP_EXP_NONE;
        o0 = (i == L1) ? 1.0 : c_sc.go; e0 = (i == L1) ? 1.0 : c_sc.ge;   // H-type gap is terminal (exp(0)) in the last row
    }
    __device__ __forceinline__ void cell(int i, int j, int, int, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        double fz = 1.0, fc = 1.0;   // factors that bring the diagonal / left operands into this row's scale
        if (SC) {
            const bool origin = (j == cb * W);        // first column of the column block: this cell fixes the row's scale
            if (origin) {
                int e = (i == 0) ? 0 : e_prev;
                int t = gmax;
                if (cbi > 0) {   // what flows in from the previous column block
                    const double cm = fmax(fmax(carry[0], carry[1]), carry[2]);
                    if (cm > 0.0) t = max(t, exp_of(cm) + (int)carry[3]);
                }
                if (i > 0 && t != MLP_EXP_NONE && (t - e > 48 || e - t > 48)) e = t;
                nw[3] = (double)e;
                rowexp[cb * (L1 + 1) + i] = e;
            } else nw[3] = carry[3];
            if (e_row == MLP_EXP_NONE) {   // first cell of this lane in this row
                e_row = (int)nw[3];
                const int d = e_prev - e_row;
                f = (d == 0) ? 1.0 : pow2d(max(min(d, 1000), -1000));
            }
            fz = f;
            if (origin && cbi > 0) {       // operands handed over by the previous column block carry that block's exponents
                fc = pow2d(max(min((int)carry[3] - e_row, 1000), -1000));
                fz = pow2d(max(min((int)diag[3] - e_row, 1000), -1000));
            }
        }
        if (i == 0 || j == 0 || j > L2) {   // boundary: Zm(0,0)=1, H(0,j>=1)=1, V(i>=1,0)=1 (terminal gaps are exp(0))
            nw[0] = (i == 0 && j == 0) ? 1.0 : 0.0;
            nw[1] = (i == 0 && j >= 1 && j <= L2) ? 1.0 : 0.0;
            nw[2] = (j == 0 && i >= 1) ? (SC ? pow2d(max(min(-e_row, 1000), -1000)) : 1.0) : 0.0;
            Z[slot] = nw[0];
        } else {
            const double score = srow[r2];
            double h = __dadd_rn(__dmul_rn(carry[0], o0), __dmul_rn(carry[1], e0));
            double v = __dadd_rn(__dmul_rn(old[0], c_sc.go), __dmul_rn(old[2], c_sc.ge));
            if (j == L2) v = __dadd_rn(old[0], old[2]);   // V-type gap is terminal in the last column: x * exp(0) == x
            double zm = __dmul_rn(sum3<!SC>(diag[0], diag[1], diag[2]), score);
            if (SC) {   // power-of-two rescaling: exact
                if (fc != 1.0) h = __dmul_rn(h, fc);
                if (f != 1.0) v = __dmul_rn(v, f);
                if (fz != 1.0) zm = __dmul_rn(zm, fz);
            }
            nw[0] = zm; nw[1] = h; nw[2] = v;
            Z[slot] = zm;
            if (SC) row_seen = max(row_seen, exp_of(zm) + e_row);
            if (i == L1 && j == L2) { has_zz = true; zz = sum3<!SC>(zm, h, v); zexp = SC ? e_row : 0; }
        }
    }
    __device__ __forceinline__ void end_row() { if (SC) { e_prev = e_row; if (row_seen != MLP_EXP_NONE) seen_exp = row_seen; } }
};

template <bool SC>
__global__ void __launch_bounds__(MLP_BLOCK) k_part_fwd_t(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* sub = reinterpret_cast<double*>(smem);
    for (int k = threadIdx.x; k < 676; k += blockDim.x) sub[k] = a.sub[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* band; double* stage; uint8_t* colres; float* cap;
    warp_smem<double, 3, 0>(smem, MLP_PART_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    double* edge = a.edge_d ? a.edge_d + gw * a.edge_stride : nullptr;
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        PartFwdT<SC> m;
        m.reset();
        m.sub = sub; m.Z = a.layerZ + t.off; m.L1 = t.L1; m.L2 = t.L2; m.W = 32 * t.C; m.qp = (a.flavour == 0);
        m.rowexp = SC ? a.rowexp + (long long)ti * a.rowexp_stride : nullptr;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        if (m.has_zz) { a.pout[ti].Zpart = m.zz; a.pout[ti].zexp = m.zexp; }
    }
}

template <bool SC>
struct PartRevT {
    typedef double T;
    enum { NS = SC ? 4 : 3, REV = 1, COLMASK = 0x7, NIN = 1, UNROLLC = 1 };
    const double* sub; const double* Z; float* P; const int* rowexp; int L1, L2, W, nb; bool qp; double Ztot; int zexp;
    const double* srow; double o0, e0;
    double* stage; int Cmax, lane;
    int seen_exp, row_seen, gmax, e_prev, e_row, fexp, cb, cbi; double f;
    __device__ __forceinline__ void reset() { seen_exp = MLP_EXP_NONE; row_seen = MLP_EXP_NONE; gmax = MLP_EXP_NONE; e_prev = 0; e_row = 0; f = 1.0; fexp = 0; cb = 0; cbi = 0; }
    __device__ __forceinline__ void begin_block(int cb_, int cbi_) { cb = cb_; cbi = cbi_; e_prev = 0; seen_exp = MLP_EXP_NONE; gmax = MLP_EXP_NONE; }
    __device__ __forceinline__ void step_sync() { if (SC) gmax = __reduce_max_sync(MLP_FULL, seen_exp); }
    __device__ __forceinline__ void prefetch(int slotbase, int C, int buf) const {
        for (int c = 0; c < C; ++c) cp_async8(stage + (buf * Cmax + c) * 32 + lane, Z + slotbase + c * 32);
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int j) const {   // virtual row L1+1 (scale exponent 0)
        st[0] = (j == L2 + 1) ? 1.0 : 0.0;
        st[1] = (j >= 1 && j <= L2) ? 1.0 : 0.0;
        st[2] = 0.0;
        if (SC) st[3] = 0.0;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int i) const {    // virtual column L2+1 when it lies outside the strips
        e[0] = 0.0; e[1] = 0.0; e[2] = (i >= 1 && i <= L1) ? 1.0 : 0.0;
        if (SC) e[3] = 0.0;
    }
    __device__ __forceinline__ int row_residue_index(int i) const { return i; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j; }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        srow = sub + r1 * 26; e_row = MLP_EXP_NONE; row_seen = MLP_EXP_NONE;
        o0 = (i == 1) ? 1.0 : c_sc.go; e0 = (i == 1) ? 1.0 : c_sc.ge;   // H-type terminal at the first row
        if (SC) fexp = rowexp[cb * (L1 + 1) + i];
    }
    __device__ __forceinline__ void cell(int i, int j, int c, int buf, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        double fz = 1.0, fc = 1.0;
        if (SC) {
            // origin of the row inside this column block: the virtual column L2+1 in the last block, the block's last column otherwise
            const bool origin = (cb == nb - 1) ? (j == L2 + 1) : (j == cb * W + W - 1);
            if (origin) {
                int e = e_prev;
                int t = gmax;
                if (cbi > 0) {
                    const double cm = fmax(fmax(carry[0], carry[1]), carry[2]);
                    if (cm > 0.0) t = max(t, exp_of(cm) + (int)carry[3]);
                }
                if (t != MLP_EXP_NONE && (t - e > 48 || e - t > 48)) e = t;
                nw[3] = (double)e;
            } else nw[3] = (j > L2 + 1) ? 0.0 : carry[3];
            if (e_row == MLP_EXP_NONE && j <= L2 + 1) {
                e_row = (int)nw[3];
                const int d = e_prev - e_row;
                f = (d == 0) ? 1.0 : pow2d(max(min(d, 1000), -1000));
            }
            fz = f;
            if (origin && cbi > 0) {
                fc = pow2d(max(min((int)carry[3] - e_row, 1000), -1000));
                fz = pow2d(max(min((int)diag[3] - e_row, 1000), -1000));
            }
        }
        if (j > L2) {
            nw[0] = 0.0; nw[1] = 0.0;
            nw[2] = (j == L2 + 1 && i >= 1) ? (SC ? pow2d(max(min(-e_row, 1000), -1000)) : 1.0) : 0.0;
            return;
        }
        if (i == 0 || j == 0) { nw[0] = 0.0; nw[1] = 0.0; nw[2] = 0.0; P[slot] = 0.0f; return; }
        const double score = srow[r2];
        double v = __dadd_rn(__dmul_rn(old[0], c_sc.go), __dmul_rn(old[2], c_sc.ge));
        if (j == 1) v = __dadd_rn(old[0], old[2]);   // V-type terminal at the first column
        double h = __dadd_rn(__dmul_rn(carry[0], o0), __dmul_rn(carry[1], e0));
        double zm = __dmul_rn(sum3<SC>(diag[0], diag[1], diag[2]), score);
        if (SC) {
            if (fc != 1.0) h = __dmul_rn(h, fc);
            if (f != 1.0) v = __dmul_rn(v, f);
            if (fz != 1.0) zm = __dmul_rn(zm, fz);
        }
        nw[0] = zm; nw[1] = h; nw[2] = v;
        if (SC) row_seen = max(row_seen, exp_of(zm) + e_row);
        // PartitionFunction.cpp:259-270 / MSAPartProbs.cpp:286-297
        double tmp = __dmul_rn(stage[(buf * Cmax + c) * 32 + lane], zm);
        tmp = __ddiv_rn(tmp, __dmul_rn(score, Ztot));
        if (SC) tmp = scalbn(tmp, fexp + e_row - zexp);
        float p = (float)tmp;
        if (!SC && !(p <= 1.0f && (double)p >= 0.001)) p = 0.0f;   // QuickProbs only (SC == cpnp)
        P[slot] = p;
    }
    __device__ __forceinline__ void end_row() { if (SC) { e_prev = e_row; if (row_seen != MLP_EXP_NONE) seen_exp = row_seen; } }
};

template <bool SC>
__global__ void __launch_bounds__(MLP_BLOCK) k_part_rev_t(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* sub = reinterpret_cast<double*>(smem);
    for (int k = threadIdx.x; k < 676; k += blockDim.x) sub[k] = a.sub[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* band; double* stage; uint8_t* colres; float* cap;
    warp_smem<double, 3, 1>(smem, MLP_PART_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    double* edge = a.edge_d ? a.edge_d + gw * a.edge_stride : nullptr;
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        PartRevT<SC> m;
        m.reset();
        m.stage = stage; m.Cmax = a.Cmax; m.lane = lane;
        m.sub = sub; m.Z = a.layerZ + t.off; m.P = a.layerP + t.off; m.L1 = t.L1; m.L2 = t.L2; m.W = 32 * t.C; m.nb = t.nb; m.qp = (a.flavour == 0);
        m.Ztot = a.pout[ti].Zpart; m.zexp = a.pout[ti].zexp;
        m.rowexp = SC ? a.rowexp + (long long)ti * a.rowexp_stride : nullptr;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
    }
}

// ------------------------------------------------------------------------------------------------ local 3-state HMM
// states: 0 = M, 1 = X, 2 = Y  (ProbabilisticModel.h flag=false branches)
struct LocFwd {
    __device__ __forceinline__ void begin_block(int, int) const {}
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ void end_row() const {}
    typedef float T;
    enum { NS = 3, REV = 0, COLMASK = 0x7, NIN = 0, UNROLLC = 0 };
    const float* match; const float* ins; const LogAddLut* lut; unsigned lutb; float* F; int L1, L2;
    float* RM;   // row-major copy of F_M for the sequential Z replay (lives in the not-yet-used Z-term layer)
    float ins1; const float* mrow;
    __device__ __forceinline__ void prefetch(int, int, int) const {}
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = st[1] = st[2] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = MLP_LOG_ZERO; }
    __device__ __forceinline__ int row_residue_index(int i) const { return i; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j; }
    __device__ __forceinline__ void begin_row(int, int r1) { ins1 = ins[r1]; mrow = match + r1 * 26; }
    __device__ __forceinline__ void cell(int i, int j, int, int, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        // ProbabilisticModel.h:210-211,222-227: base = ((m - a) - b); M = (base - 2r) (+) sum_k ((base + F_k) + lt[k][0]) - 2r
        const float base = __fsub_rn(__fsub_rn(mrow[r2], ins1), ins[r2]);
        float m = __fsub_rn(base, c_sc.r2);
        m = dev_log_add_lutb(m, __fsub_rn(__fadd_rn(__fadd_rn(base, diag[0]), c_sc.lt00), c_sc.r2), lutb);
        m = dev_log_add_lutb(m, __fsub_rn(__fadd_rn(__fadd_rn(base, diag[1]), c_sc.lt10), c_sc.r2), lutb);
        m = dev_log_add_lutb(m, __fsub_rn(__fadd_rn(__fadd_rn(base, diag[2]), c_sc.lt20), c_sc.r2), lutb);
        // :238-241, :252-255
        float x = dev_log_add_lutb(__fsub_rn(__fadd_rn(old[0], c_sc.lt01), c_sc.r), __fsub_rn(__fadd_rn(old[1], c_sc.lt11), c_sc.r), lutb);
        float y = dev_log_add_lutb(__fsub_rn(__fadd_rn(carry[0], c_sc.lt02), c_sc.r), __fsub_rn(__fadd_rn(carry[2], c_sc.lt22), c_sc.r), lutb);
        if (i == 0 || j == 0) m = MLP_LOG_ZERO;
        if (i == 0) x = MLP_LOG_ZERO;
        if (j == 0) y = MLP_LOG_ZERO;
        if (i <= 1 && j <= 1) {
            m = (i == 1 && j == 1) ? __fsub_rn(base, c_sc.r2) : MLP_LOG_ZERO;
            x = MLP_LOG_ZERO; y = MLP_LOG_ZERO;
        }
        nw[0] = m; nw[1] = x; nw[2] = y;
        F[slot] = m;
        if (j <= L2) RM[i * (L2 + 1) + j] = m;
    }
};

struct LocBwd {
    __device__ __forceinline__ void begin_block(int, int) const {}
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ void end_row() const {}
    typedef float T;
    enum { NS = 3, REV = 1, COLMASK = 0x3, NIN = 1, UNROLLC = 0 };   // keep B_M and X of row i+1; Y travels along the row
    const float* match; const float* ins; const LogAddLut* lut; unsigned lutb; float* F; float* VB; int L1, L2;
    float* stage; int Cmax, lane;
    __device__ __forceinline__ void prefetch(int slotbase, int C, int buf) const {
        for (int c = 0; c < C; ++c) cp_async4(stage + (buf * Cmax + c) * 32 + lane, F + slotbase + c * 32);
    }
    float ins1n; const float* mrown;   // residue i+1 (transition out of the cell)
    float ins1c; const float* mrowc;   // residue i   (the cell's own emission, for the Z term)
    const uint8_t* s1; const uint8_t* s2;
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = st[1] = st[2] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = MLP_LOG_ZERO; }
    __device__ __forceinline__ int row_residue_index(int i) const { return i + 1; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j + 1; }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        ins1n = ins[r1]; mrown = match + r1 * 26;
        const int rc = (i >= 1) ? s1[i - 1] : 0;
        ins1c = ins[rc]; mrowc = match + rc * 26;
    }
    __device__ __forceinline__ void cell(int i, int j, int c, int buf, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        if (j > L2) { nw[0] = nw[1] = nw[2] = MLP_LOG_ZERO; return; }
        // ProbabilisticModel.h:339-379 flag=false.  B_M starts at LOG_ONE in every cell.
        float bm = 0.0f, x = MLP_LOG_ZERO, y = MLP_LOG_ZERO;
        if (i < L1 && j < L2) {
            const float pxy = __fsub_rn(__fsub_rn(__fadd_rn(diag[0], mrown[r2]), ins1n), ins[r2]);
            bm = dev_log_add_lutb(bm, __fsub_rn(__fadd_rn(pxy, c_sc.lt00), c_sc.r2), lutb);
            x = __fsub_rn(__fadd_rn(pxy, c_sc.lt10), c_sc.r2);
            y = __fsub_rn(__fadd_rn(pxy, c_sc.lt20), c_sc.r2);
        }
        if (i < L1) {
            bm = dev_log_add_lutb(bm, __fsub_rn(__fadd_rn(old[1], c_sc.lt01), c_sc.r), lutb);
            x = dev_log_add_lutb(x, __fsub_rn(__fadd_rn(old[1], c_sc.lt11), c_sc.r), lutb);
        }
        if (j < L2) {
            bm = dev_log_add_lutb(bm, __fsub_rn(__fadd_rn(carry[2], c_sc.lt02), c_sc.r), lutb);
            y = dev_log_add_lutb(y, __fsub_rn(__fadd_rn(carry[2], c_sc.lt22), c_sc.r), lutb);
        }
        nw[0] = bm; nw[1] = x; nw[2] = y;
        // Z term of this cell, ProbabilisticModel.h:445-446: (((B_M + m) - a) - b) - 2r with the cell's own residues
        float vb = MLP_LOG_ZERO;
        if (i >= 1 && j >= 1) {
            const int rj = s2[j - 1];
            vb = __fsub_rn(__fsub_rn(__fsub_rn(__fadd_rn(bm, mrowc[rj]), ins1c), ins[rj]), c_sc.r2);
        }
        VB[i * (L2 + 1) + j] = vb;   // row-major: read only by the sequential Z replay
        F[slot] = __fadd_rn(stage[(buf * Cmax + c) * 32 + lane], bm);
    }
};

// Exact replay of the reference's sequential row-major LOG_PLUS_EQUALS chain over a dense slot-layout layer
// (ProbabilisticModel.h:434-451).  The running sum is monotone (LOOKUP(d) > d's loss), so a
// cell more than 7.5 below the sum can never change it: a warp tests 32 cells at once and applies only the
// cells that fire, in order -- bit-identical to the serial chain.
__device__ float replay_rowmajor(const float* __restrict__ rm /* row-major (L1+1)x(L2+1) */, const SweepCtx& cx) {
    float sum = MLP_LOG_ZERO;
    const int lane = cx.lane;
    const int W = cx.L2 + 1;
    // The chain is serial, the loads are not: the 32-column chunks of rows 1..L1 are visited in row-major order and the next
    // four chunks are always in flight (round 1 issued one dependent global load per chunk: ~3000 exposed latencies per pair).
    const int nch = (cx.L2 + 31) >> 5;
    const long long total = (long long)cx.L1 * nch;
    int pi = 1, pk = 0;                       // row / chunk of the next chunk to fetch
    auto fetch = [&]() -> float {
        float v = MLP_LOG_ZERO;
        if (pi <= cx.L1) {
            const int j = 1 + (pk << 5) + lane;
            if (j <= cx.L2) v = rm[(long long)pi * W + j];
            if (++pk == nch) { pk = 0; ++pi; }
        }
        return v;
    };
    float b0 = fetch(), b1 = fetch(), b2 = fetch(), b3 = fetch();
    for (long long q = 0; q < total; ++q) {
        const float v = b0;
        b0 = b1; b1 = b2; b2 = b3; b3 = fetch();
        int pos = 0;
        for (;;) {
            // the cell changes the sum unless sum >= v and (v == LOG_ZERO or sum - v >= 7.5); padding lanes hold LOG_ZERO and never fire
            // (the running sum starts at LOG_ZERO and only grows)
            const bool fires = (lane >= pos) && !(sum >= v && (v == MLP_LOG_ZERO || __fsub_rn(sum, v) >= 7.5f));
            const unsigned mask = __ballot_sync(MLP_FULL, fires);
            if (mask == 0) break;
            const int l0 = __ffs(mask) - 1;
            const float vv = __shfl_sync(MLP_FULL, v, l0);
            sum = dev_log_add(sum, vv);
            pos = l0 + 1;
        }
    }
    return sum;
}

__global__ void __launch_bounds__(MLP_BLOCK) k_loc_fwd(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* band; float* stage; uint8_t* colres; float* cap;
    warp_smem<float, 3, 0>(smem, MLP_HMM_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const unsigned lutb = log_add_lut_bias(lut);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        LocFwd m;
        m.match = match; m.ins = ins; m.lut = lut; m.lutb = lutb; m.F = a.layerSL + t.off; m.RM = a.layerVB + t.off; m.L1 = t.L1; m.L2 = t.L2;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        __syncwarp();
        __threadfence_block();
        const float tF = replay_rowmajor(a.layerVB + t.off, cx);
        if (lane == 0) a.pout[ti].tFL = tF;
    }
}

__global__ void __launch_bounds__(MLP_BLOCK) k_loc_bwd(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* band; float* stage; uint8_t* colres; float* cap;
    warp_smem<float, 2, 1>(smem, MLP_HMM_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const unsigned lutb = log_add_lut_bias(lut);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        LocBwd m;
        m.stage = stage; m.Cmax = a.Cmax; m.lane = lane; m.lut = lut; m.lutb = lutb;
        m.match = match; m.ins = ins; m.F = a.layerSL + t.off; m.VB = a.layerVB + t.off; m.L1 = t.L1; m.L2 = t.L2; m.s1 = cx.s1; m.s2 = cx.s2;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        __syncwarp();
        __threadfence_block();
        const float tB = replay_rowmajor(a.layerVB + t.off, cx);
        if (lane == 0) a.pout[ti].totalL = __fdiv_rn(__fadd_rn(a.pout[ti].tFL, tB), 2.0f);   // ProbabilisticModel.h:453
    }
}

// ------------------------------------------------------------------------------------------------ Viterbi (model selection)
// ProbabilisticModel.h:1043-1170: 3-state Viterbi with traceback; states 0 = M, 1 = X, 2 = Y.  The three traceback
// entries of a cell are packed into one byte (M: 2 bits holding tb+1, X: bit 2, Y: bit 3) stored in slot layout.
struct VitFwd {
    typedef float T;
    enum { NS = 3, REV = 0, COLMASK = 0x7, NIN = 0, UNROLLC = 0 };
    const float* match; const float* ins; unsigned char* TB; int L1, L2;
    float ins1; const float* mrow; float init0, init1;
    float fin[3]; bool has_fin;
    __device__ __forceinline__ void begin_block(int, int) const {}
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ void end_row() const {}
    __device__ __forceinline__ void prefetch(int, int, int) const {}
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = st[1] = st[2] = MLP_LOG_ZERO; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = MLP_LOG_ZERO; }
    __device__ __forceinline__ int row_residue_index(int i) const { return i; }
    __device__ __forceinline__ int col_residue_index(int j) const { return j; }
    __device__ __forceinline__ void begin_row(int, int r1) { ins1 = ins[r1]; mrow = match + r1 * 26; }
    __device__ __forceinline__ void cell(int i, int j, int, int, int r2, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        float m = MLP_LOG_ZERO, x = MLP_LOG_ZERO, y = MLP_LOG_ZERO;
        int tbm = -1, tbx = 0, tby = 0;
        if (i > 0 && j > 0) {   // :1091-1099, strict '<' keeps the first maximum
            const float mt = mrow[r2];
            const float c0 = __fadd_rn(__fadd_rn(diag[0], c_sc.lt00), mt);
            const float c1 = __fadd_rn(__fadd_rn(diag[1], c_sc.lt10), mt);
            const float c2 = __fadd_rn(__fadd_rn(diag[2], c_sc.lt20), mt);
            if (m < c0) { m = c0; tbm = 0; }
            if (m < c1) { m = c1; tbm = 1; }
            if (m < c2) { m = c2; tbm = 2; }
        }
        if (i > 0) {            // :1100-1113
            const float fm = __fadd_rn(__fadd_rn(ins1, old[0]), c_sc.lt01);
            const float fi = __fadd_rn(__fadd_rn(ins1, old[1]), c_sc.lt11);
            if (fm >= fi) { x = fm; tbx = 0; } else { x = fi; tbx = 1; }
        }
        if (j > 0) {            // :1114-1127
            const float ins2 = ins[r2];
            const float fm = __fadd_rn(__fadd_rn(ins2, carry[0]), c_sc.lt02);
            const float fi = __fadd_rn(__fadd_rn(ins2, carry[2]), c_sc.lt22);
            if (fm >= fi) { y = fm; tby = 0; } else { y = fi; tby = 1; }
        }
        if (i == 0 && j == 0) { m = init0; x = init1; y = init1; }   // :1070-1072
        nw[0] = m; nw[1] = x; nw[2] = y;
        if (j <= L2) TB[slot] = (unsigned char)((tbm + 1) | (tbx << 2) | (tby << 3));
        if (i == L1 && j == L2) { has_fin = true; fin[0] = m; fin[1] = x; fin[2] = y; }
    }
};

__global__ void __launch_bounds__(MLP_BLOCK) k_viterbi(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* match; float* ins; LogAddLut* lut;
    load_hmm_tables(smem, a, match, ins, lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* band; float* stage; uint8_t* colres; float* cap;
    warp_smem<float, 3, 0>(smem, MLP_HMM_TABLE_BYTES, a.Cmax, warp, band, stage, colres, cap);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    int* capi = reinterpret_cast<int*>(cap);
    // ProbabilisticModel.h:1070-1072: LOG() of double literals narrowed to float, evaluated on the host with glibc logf
    const float init0 = a.vit_init0, init1 = a.vit_init1;
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        VitFwd m;
        m.match = match; m.ins = ins; m.TB = a.layerTB8 + t.off; m.L1 = t.L1; m.L2 = t.L2; m.init0 = init0; m.init1 = init1; m.has_fin = false;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        if (m.has_fin) {   // best terminating state, :1137-1149 (strict '<', first maximum)
            float best = MLP_LOG_ZERO; int state = -1;
            const float iv[3] = {init0, init1, init1};
#pragma unroll
            for (int k = 0; k < 3; ++k) { const float p = __fadd_rn(m.fin[k], iv[k]); if (best < p) { best = p; state = k; } }
            capi[0] = state;
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) {   // traceback :1155-1163 and the identity count of MSA.cpp:819-836
            int state = capi[0], r = t.L1, c = t.L2, len = 0, same = 0;
            const int W = 32 * t.C;
            const unsigned char* tb = a.layerTB8 + t.off;
            char* aln = a.vit_aln ? a.vit_aln + a.vit_aln_off[t.pidx] : nullptr;
            while ((r != 0 || c != 0) && state >= 0) {
                const int cb = c / W, rem = c - cb * W, l = rem / t.C, cc = rem - l * t.C;
                const int byte = tb[((cb * cx.T + r + l) * t.C + cc) * 32 + l];
                int ns;
                if (aln) aln[len] = (state == 0) ? 'B' : ((state == 1) ? 'X' : 'Y');   // written back to front
                if (state == 0) { ns = (byte & 3) - 1; same += (cx.s1[r - 1] == cx.s2[c - 1]); --r; --c; }
                else if (state == 1) { ns = (byte >> 2) & 1; --r; }
                else { ns = ((byte >> 3) & 1) ? 2 : 0; --c; }
                ++len;
                state = ns;
            }
            a.vit_ident[t.pidx] = same;
            a.vit_len[t.pidx] = len;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ merge + MEA + sparsify
// states: 0 = MEA row score, 1 = number of kept cells so far in this row (exact small integer in a float)
// QPSPEC: compile-time constants for the QuickProbs default (flavour 0, models HMM5|PART, no traceback layer): the per-cell
// branches on mask / flavour / tb fold away
template <bool DENSE, bool QPSPEC>
struct FinalSweep {
    __device__ __forceinline__ void begin_block(int, int) const {}
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ void end_row() const {}
    typedef float T;
    enum { NS = 2, REV = 0, COLMASK = 0x1, NIN = 3, UNROLLC = 1 };
    const float* S5; const float* P; const float* SL;
    float* dstage; int Cmax, lane;
    __device__ __forceinline__ void prefetch(int slotbase, int C, int buf) const {
        const unsigned mask = QPSPEC ? 3u : this->mask;
        for (int c = 0; c < C; ++c) {
            const int g = slotbase + c * 32;
            float* d = dstage + ((buf * 3) * Cmax + c) * 32 + lane;
            if (mask & 1u) cp_async4(d, S5 + g);
            if (mask & 2u) cp_async4(d + Cmax * 32, P + g);
            if (mask & 4u) cp_async4(d + 2 * Cmax * 32, SL + g);
        }
    }
    float total5, totalL;
    int flavour; unsigned mask; float cutoff;
    int L1, L2;
    int* rowcnt;            // [L1+2] in the row-pointer pool (counts first, scanned later)
    int4* stage; int stage_cap; int* stage_n;   // per-warp staging of kept cells
    float* dense;           // optional dense dump (debug), row-major (L1+1)x(L2+1)
    float* dense5; float* denseP; float* denseL;
    float score; bool has_score;
    const ExpLut* elut;
    int* tb;                // traceback codes 0 = D, 1 = L, 2 = U (ProbabilisticModel.h:834-836), MLP_CPNP_P1 only
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = 0.0f; st[1] = 0.0f; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = 0.0f; e[1] = 0.0f; }
    __device__ __forceinline__ int row_residue_index(int) const { return 0; }
    __device__ __forceinline__ int col_residue_index(int) const { return 0; }
    __device__ __forceinline__ void begin_row(int, int) {}
    __device__ __forceinline__ void cell(int i, int j, int c, int buf, int, int slot, const T (&old)[NS], const T (&carry)[NS],
                                         const T (&diag)[NS], T (&nw)[NS]) {
        // Written branch-free except for the (rare) kept-cell staging: padding columns j > L2 compute on whatever the stage holds
        // and are masked out at the end, which is cheaper than a divergent early return per cell.
        const bool valid = (j <= L2);
        const bool inner = valid && i >= 1 && j >= 1;
        const unsigned mask = QPSPEC ? 3u : this->mask;
        const int flavour = QPSPEC ? 0 : this->flavour;
        int* const tb = QPSPEC ? nullptr : this->tb;
        float v5 = 0.0f, vp = 0.0f, vl = 0.0f, p;
        const float* sg = dstage + ((buf * 3) * Cmax + c) * 32 + lane;
        if (mask & 1u) v5 = dev_exp_lut(fminf(0.0f, __fsub_rn(sg[0], total5)), elut);
        if (mask & 2u) vp = sg[Cmax * 32];
        if (mask & 4u) vl = dev_exp_lut(fminf(0.0f, __fsub_rn(sg[2 * Cmax * 32], totalL)), elut);
        const bool origin = (i == 0 && j == 0);
        v5 = origin ? 0.0f : v5;                          // posterior[0] = 0, ProbabilisticModel.h:490
        vl = origin ? 0.0f : vl;
        if (flavour == 0) {
            // PosteriorStage.cpp:169-177: borders forced to 0, sqrt((v1^2+v2^2)*0.5)
            p = __fsqrt_rn(__fmul_rn(__fadd_rn(__fmul_rn(v5, v5), __fmul_rn(vp, vp)), 0.5f));
            p = (i == 0 || j == 0) ? 0.0f : p;
        } else if (mask == 7u) {
            // MSA.cpp:1001 ((dbl^2+glob^2)+loc^2)/3 ; MSA.cpp:1708 ((glob^2+loc^2)+dbl^2)/3
            const float q5 = __fmul_rn(v5, v5), qp = __fmul_rn(vp, vp), ql = __fmul_rn(vl, vl);
            const float s = (flavour == 2) ? __fadd_rn(__fadd_rn(qp, ql), q5) : __fadd_rn(__fadd_rn(q5, qp), ql);
            p = __fsqrt_rn(__fdiv_rn(s, 3.0f));
        } else {
            p = (mask & 1u) ? v5 : ((mask & 2u) ? vp : vl);
        }
        if (DENSE) {
            if (valid) {
                const long long d = (long long)i * (L2 + 1) + j;
                dense[d] = p;
                if (dense5) dense5[d] = v5;
                if (denseP) denseP[d] = vp;
                if (denseL) denseL[d] = vl;
            }
        }
        // MEA row DP: ProbabilisticModel.h:834-836 / PosteriorStage.cpp:177 (row 0 / column 0 stay 0)
        const float x1 = __fadd_rn(p, diag[0]), x2 = carry[0], x3 = old[0];
        float sc = inner ? fmaxf(fmaxf(x1, x2), x3) : 0.0f;
        if (tb) {   // ChooseBestOfThree tie order D >= L >= U (ScoreType.h:347-366); row 0 = 'L', column 0 = 'U'
            int code = (x1 >= x2) ? ((x1 >= x3) ? 0 : 2) : ((x2 >= x3) ? 1 : 2);
            code = (j >= 1) ? code : 2;
            code = (i >= 1) ? code : 1;
            if (valid) tb[slot] = code;
        }
        float cnt = (j == 0) ? 0.0f : carry[1];
        if (inner && p >= cutoff) {   // SparseMatrix.h:89 / PackedSparseMatrix.cpp:68
            const int k = atomicAdd(stage_n, 1);
            if (k < stage_cap) stage[k] = make_int4(i, (int)cnt, j, __float_as_int(p));
            cnt += 1.0f;
        }
        sc = valid ? sc : carry[0];
        cnt = valid ? cnt : carry[1];
        nw[0] = sc; nw[1] = cnt;
        if (j == L2) {
            rowcnt[i + 1] = (int)cnt;
            if (i == L1) { has_score = true; score = sc; }
        }
    }
};

template <bool DENSE, bool QPSPEC>
__global__ void __launch_bounds__(MLP_BLOCK) k_final_t(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    ExpLut* elut = reinterpret_cast<ExpLut*>(smem);
    exp_lut_fill(elut, threadIdx.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* band; float* stg; uint8_t* colres; float* cap;
    warp_smem<float, 1, 3>(smem, MLP_FINAL_TABLE_BYTES, a.Cmax, warp, band, stg, colres, cap);
    int* stage_n = reinterpret_cast<int*>(cap);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    float* edge = a.edge_f ? a.edge_f + gw * a.edge_stride : nullptr;
    int4* stage = a.stage + gw * a.stage_cap;
    for (;;) {
        const int ti = next_task(a.counter, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx cx = make_ctx(t, a, lane);
        const long long slotAB = (long long)t.a * a.n + t.b;
        int* rowptr = a.out.rp_pool + a.rp_off[slotAB];
        if (lane == 0) { *stage_n = 0; rowptr[0] = 0; rowptr[1] = 0; }
        __syncwarp();
        FinalSweep<DENSE, QPSPEC> m;
        m.elut = elut;
        m.S5 = a.layerS5 ? a.layerS5 + t.off : nullptr; m.P = a.layerP ? a.layerP + t.off : nullptr; m.SL = a.layerSL ? a.layerSL + t.off : nullptr; m.dstage = stg; m.Cmax = a.Cmax; m.lane = lane;
        m.total5 = a.pout[ti].total5; m.totalL = a.pout[ti].totalL;
        m.flavour = a.flavour; m.mask = a.mask; m.cutoff = a.cutoff;
        m.L1 = t.L1; m.L2 = t.L2; m.rowcnt = rowptr; m.stage = stage; m.stage_cap = a.stage_cap; m.stage_n = stage_n;
        m.dense = a.dense; m.dense5 = a.dense5; m.denseP = a.denseP; m.denseL = a.denseL;
        m.has_score = false; m.score = 0.0f;
        m.tb = (a.flavour == 2 && a.layerTB) ? a.layerTB + t.off : nullptr;
        run_sweep(m, cx, band, colres, a.Cmax, edge);
        if (m.has_score) a.pout[ti].mea = m.score;
        __syncwarp();
        __threadfence_block();
        if (lane == 0) {
            const float score = a.pout[ti].mea;
            float dist;
            if (a.flavour == 2) {
                // ArrangePosteriorProbs, MSA.cpp:1744-1752: distance = score / number of 'B' columns on the MEA traceback
                int r = t.L1, c = t.L2, nb = 0;
                const int W = 32 * t.C;
                while (r != 0 || c != 0) {
                    const int cb = c / W, rem = c - cb * W, l = rem / t.C, cc = rem - l * t.C;
                    const int code = m.tb[((cb * cx.T + r + l) * t.C + cc) * 32 + l];
                    if (code == 1) --c; else if (code == 2) --r; else { --r; --c; ++nb; }
                }
                dist = __fdiv_rn(score, (float)nb);
            } else {
                dist = __fsub_rn(1.0f, __fdiv_rn(score, (float)min(t.L1, t.L2)));   // MSA.cpp:1019 / PosteriorStage.cpp:194
            }
            a.dist[(long long)t.a * a.n + t.b] = dist;
            a.dist[(long long)t.b * a.n + t.a] = dist;
        }
        __syncwarp();
        __threadfence_block();
        // exclusive scan of the per-row counts -> row pointers (row i occupies rowptr[i]..rowptr[i+1])
        int run = 0;
        for (int base = 1; base <= t.L1; base += 32) {
            const int i = base + lane;
            int v = (i <= t.L1) ? rowptr[i + 1] : 0;
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
                a.out.cells[d] = make_int2(r.z, __float_as_int(a.flavour == 0 ? dev_quantize_u16(v) : v));
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ transpose
// Stable counting-sort transpose, one warp per pair.  Cells are visited in row-major order 32 at a time; lanes
// that hit the same column are ranked with __match_any_sync so the row order inside each transposed row is
// the reference's (PackedSparseMatrix.cpp:119-134 / SparseMatrix.h:226-244).
__global__ void __launch_bounds__(MLP_BLOCK) k_transpose(KArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    // A capacity miss in the producing kernel (k_final / k_relax_blk) leaves pairs whose cells were never written: their
    // columns are stale memory and must not be used as indices.  The host re-runs the batch after growing the pools, so
    // the whole transpose is skipped as soon as the error word is set.
    if (*reinterpret_cast<volatile int*>(a.err) != 0) return;
    for (long long ti = gw; ti < a.ntasks; ti += nw) {
        const PairTask t = a.tasks[ti];
        const long long sAB = (long long)t.a * a.n + t.b, sBA = (long long)t.b * a.n + t.a;
        const int* rp = a.out.rp_pool + a.rp_off[sAB];
        int* trp = a.out.rp_pool + a.rp_off[sBA];
        const int nnz = a.out.nz_cnt[sAB];
        const long long src = a.out.nz_off[sAB];
        long long dst = 0;
        if (lane == 0) {
            dst = (long long)atomicAdd(a.out.cursor, (unsigned long long)nnz);
            a.out.nz_off[sBA] = dst; a.out.nz_cnt[sBA] = nnz;
            if (dst + nnz > a.out.cap) atomicOr(a.err, 2);
        }
        dst = __shfl_sync(MLP_FULL, dst, 0);
        if (dst + nnz > a.out.cap || src + nnz > a.out.cap) continue;
        // 1. column histogram into trp[j+1]
        for (int j = lane; j <= t.L2 + 1; j += 32) trp[j] = 0;
        __syncwarp();
        for (int k = lane; k < nnz; k += 32) {
            const int col = a.out.cells[src + k].x;
            if ((unsigned)col > (unsigned)t.L2) { atomicOr(a.err, 8); continue; }   // never index with a column outside the matrix
            atomicAdd(&trp[col + 1], 1);
        }
        __syncwarp();
        // 2. inclusive scan over trp[2..L2+1] -> trp[j+1] = end of row j ; trp[j] = start of row j
        int run = 0;
        for (int base = 1; base <= t.L2; base += 32) {
            const int j = base + lane;
            int inc = (j <= t.L2) ? trp[j + 1] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(MLP_FULL, inc, d); if (lane >= d) inc += o; }
            if (j <= t.L2) trp[j + 1] = run + inc;
            run += __shfl_sync(MLP_FULL, inc, 31);
        }
        __syncwarp();
        // 3. stable placement; fill counters live in a.tfill (per warp, L2max+2 ints)
        int* fill = a.tfill + gw * a.tfill_stride;
        for (int j = lane; j <= t.L2 + 1; j += 32) fill[j] = 0;
        __syncwarp();
        // row of cell k: walk rows with a moving pointer
        int row = 1;
        for (int k0 = 0; k0 < nnz; k0 += 32) {
            const int k = k0 + lane;
            const bool ok = k < nnz;
            int col = 0, r = 0; int2 cf = make_int2(0, 0);
            if (ok) {
                cf = a.out.cells[src + k]; col = min(max(cf.x, 0), t.L2);
                // binary search the row containing k: largest r with rp[r] <= k
                int lo = row, hi = t.L1;
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rp[mid] <= k) lo = mid; else hi = mid - 1; }
                r = lo;
            }
            const unsigned peers = __match_any_sync(MLP_FULL, ok ? col : -1 - lane);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (ok) {
                const long long d = dst + trp[col] + fill[col] + rank;
                a.out.cells[d] = make_int2(r, cf.y);
            }
            __syncwarp();
            if (ok && (lane == 31 || (peers >> (lane + 1)) == 0)) fill[col] += __popc(peers);   // highest lane of each group
            __syncwarp();
            row = __shfl_sync(MLP_FULL, r, 31);
            if (row < 1) row = 1;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ launch helpers
size_t posterior_smem_bytes(int kernel, int Cmax, int warps) {
    size_t tables = 0, per = 0;   // per = band + staging, must mirror warp_smem<T, NS, NIN>
    switch (kernel) {
        case MLP_K_PART_FWD: tables = MLP_PART_TABLE_BYTES; per = (size_t)3 * Cmax * 32 * 8; break;
        case MLP_K_PART_REV: tables = MLP_PART_TABLE_BYTES; per = (size_t)(3 + 2) * Cmax * 32 * 8; break;
        case MLP_K_HMM_FWD: tables = MLP_HMM_TABLE_BYTES; per = (size_t)5 * Cmax * 32 * 4; break;
        case MLP_K_HMM_BWD: tables = MLP_HMM_TABLE_BYTES; per = (size_t)(3 + 2) * Cmax * 32 * 4; break;
        case MLP_K_LOCAL_FWD: case MLP_K_VITERBI: tables = MLP_HMM_TABLE_BYTES; per = (size_t)3 * Cmax * 32 * 4; break;
        case MLP_K_LOCAL_BWD: tables = MLP_HMM_TABLE_BYTES; per = (size_t)(2 + 2) * Cmax * 32 * 4; break;
        case MLP_K_FINAL: tables = MLP_FINAL_TABLE_BYTES; per = (size_t)(1 + 6) * Cmax * 32 * 4; break;
        default: return 0;
    }
    per += Cmax * 32 + 64;
    per = (per + 15) & ~(size_t)15;
    return tables + per * warps;
}

cudaError_t posterior_set_scalars(const DevScalars& s, cudaStream_t st) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_sc, &s, sizeof(DevScalars), 0, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = part_c_set_scalars(s, st);
    if (e == cudaSuccess) e = hmm_c_set_scalars(s, st);
    if (e == cudaSuccess) e = loc_c_set_scalars(s, st);
    if (e == cudaSuccess) e = part_sc_set_scalars(s, st);
    return e;
}

// ---- register-band kernels (one instantiation per columns-per-lane value)
// developer knob: MLP_OLD_SWEEP = bit mask of kernels that use the round-1 shared-memory-band version (1 part_fwd, 2 part_rev, 4 hmm_fwd,
// 8 hmm_bwd, 16 final, 32 the local model's sweeps + Z chain), read at every launch so that a test can A/B inside one process
// model mix of the C-specialised merge kernel (final_c.cu): 0 = QuickProbs default, a c_p_np_aln -p 0 model mask, -1 = general kernel
static int final_c_mode(const KArgs& a) {
    if (a.flavour == 0) return a.mask == 3u ? 0 : -1;
    if (a.flavour == 1 && (a.mask == 1u || a.mask == 2u || a.mask == 4u || a.mask == 7u)) return (int)a.mask;
    return -1;
}
static int old_sweeps() { const char* e = getenv("MLP_OLD_SWEEP"); return e ? atoi(e) : 0; }
bool posterior_c_available(int kernel, const KArgs& a) {
    if (a.dense) return false;
    const int old = old_sweeps();
    if ((kernel == MLP_K_PART_FWD && (old & 1)) || (kernel == MLP_K_PART_REV && (old & 2)) || (kernel == MLP_K_HMM_FWD && (old & 4)) ||
        (kernel == MLP_K_HMM_BWD && (old & 8)) || (kernel == MLP_K_FINAL && (old & 16))) return false;
    if ((kernel == MLP_K_LOCAL_FWD || kernel == MLP_K_LOCAL_BWD || kernel == MLP_K_LOCAL_CAND) && (old & 32)) return false;
    switch (kernel) {
        case MLP_K_PART_FWD: case MLP_K_PART_REV: return true;   // QuickProbs: part_c.cu (plain FP64); c_p_np_aln: part_sc.cu (rescaled FP64)
        case MLP_K_HMM_FWD: case MLP_K_HMM_BWD: return true;
        case MLP_K_LOCAL_FWD: case MLP_K_LOCAL_BWD: case MLP_K_LOCAL_CAND: return a.layerLC != nullptr;
        case MLP_K_FINAL: return final_c_mode(a) >= 0;
        default: return false;
    }
}
size_t posterior_c_smem(int kernel, const KArgs& a) {
    const int warps = MLP_BLOCK / 32;
    switch (kernel) {
        // tables | small per-warp words (backward caps / stage counters) | per-warp edge + residue ring (sweep_c.cuh)
        case MLP_K_PART_FWD: case MLP_K_PART_REV: return MLP_PART_TABLE_BYTES + warps * MLP_SWEEP_RING_BYTES(a.flavour == 0 ? 3 : 4, 8);   // cpnp: + the row's scale exponent
        case MLP_K_HMM_FWD: case MLP_K_HMM_BWD: return MLP_HMM_TABLE_BYTES + 128 + warps * MLP_SWEEP_RING_BYTES(5, 4);
        case MLP_K_FINAL: return MLP_FINAL_TABLE_BYTES + 64 + warps * MLP_SWEEP_RING_BYTES(2, 4);
        case MLP_K_LOCAL_FWD: case MLP_K_LOCAL_BWD: return MLP_HMM_TABLE_BYTES + 128 + warps * MLP_SWEEP_RING_BYTES(4, 4);
        case MLP_K_LOCAL_CAND: return 64 + warps * MLP_SWEEP_RING_BYTES(2, 4);
        default: return 0;
    }
}
static void (*c_kernel(int kernel, int C, const KArgs& a))(KArgs) {
    switch (kernel) {
        case MLP_K_PART_FWD: case MLP_K_PART_REV: return a.flavour == 0 ? part_c_kernel(kernel, C) : part_sc_kernel(kernel, C);
        case MLP_K_HMM_FWD: case MLP_K_HMM_BWD: return hmm_c_kernel(kernel, C);
        case MLP_K_FINAL: return final_c_kernel(C, final_c_mode(a));
        case MLP_K_LOCAL_FWD: case MLP_K_LOCAL_BWD: case MLP_K_LOCAL_CAND: return loc_c_kernel(kernel, C);
        default: return nullptr;
    }
}
int posterior_c_max_blocks_per_sm(int kernel, int C, const KArgs& a) {
    void (*fn)(KArgs) = c_kernel(kernel, C, a);
    int nb = 0;
    if (!fn || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, MLP_BLOCK, posterior_c_smem(kernel, a)) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
cudaError_t posterior_c_launch(int kernel, int C, const KArgs& a, int grid, cudaStream_t st) {
    void (*fn)(KArgs) = c_kernel(kernel, C, a);
    if (!fn) return cudaErrorInvalidValue;
    fn<<<grid, MLP_BLOCK, posterior_c_smem(kernel, a), st>>>(a);
    return cudaGetLastError();
}

cudaError_t posterior_launch(int kernel, const KArgs& a, int grid, size_t smem, cudaStream_t st) {
    void (*fn)(KArgs) = nullptr;
    const bool a_dense = a.dense != nullptr;
    const bool a_scaled = a.flavour != 0;
    switch (kernel) {
        case MLP_K_PART_FWD: fn = a_scaled ? k_part_fwd_t<true> : k_part_fwd_t<false>; break;
        case MLP_K_PART_REV: fn = a_scaled ? k_part_rev_t<true> : k_part_rev_t<false>; break;
        case MLP_K_HMM_FWD: fn = k_hmm_fwd; break;
        case MLP_K_HMM_BWD: fn = k_hmm_bwd; break;
        case MLP_K_LOCAL_FWD: fn = k_loc_fwd; break;
        case MLP_K_LOCAL_BWD: fn = k_loc_bwd; break;
        case MLP_K_FINAL: fn = a_dense ? k_final_t<true, false> : ((a.flavour == 0 && a.mask == 3u) ? k_final_t<false, true> : k_final_t<false, false>); break;
        case MLP_K_TRANSPOSE: fn = k_transpose; break;
        case MLP_K_VITERBI: fn = k_viterbi; break;
        default: return cudaErrorInvalidValue;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    fn<<<grid, MLP_BLOCK, smem, st>>>(a);
    return cudaGetLastError();
}

int posterior_max_blocks_per_sm(int kernel, size_t smem) {
    void (*fn)(KArgs) = nullptr;
    const bool a_dense = false, a_scaled = false;
    switch (kernel) {
        case MLP_K_PART_FWD: fn = a_scaled ? k_part_fwd_t<true> : k_part_fwd_t<false>; break;
        case MLP_K_PART_REV: fn = a_scaled ? k_part_rev_t<true> : k_part_rev_t<false>; break;
        case MLP_K_HMM_FWD: fn = k_hmm_fwd; break;
        case MLP_K_HMM_BWD: fn = k_hmm_bwd; break;
        case MLP_K_LOCAL_FWD: fn = k_loc_fwd; break;
        case MLP_K_LOCAL_BWD: fn = k_loc_bwd; break;
        case MLP_K_FINAL: fn = a_dense ? k_final_t<true, false> : k_final_t<false, true>; break;
        case MLP_K_TRANSPOSE: fn = k_transpose; break;
        case MLP_K_VITERBI: fn = k_viterbi; break;
        default: return 1;
    }
    if (smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, MLP_BLOCK, smem) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
