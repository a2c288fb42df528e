// c_p_np_aln's partition-function posterior (MSAPartProbs.cpp:78-660, 80-bit long double in the reference), rescaled FP64 on the
// register-band sweeps (sweep_c.cuh).  The arithmetic, the per-row power-of-two scale and the way a row's scale is chosen are those of
// k_part_fwd_t<true> / k_part_rev_t<true> (posterior.cu), which stay as the fallback (MLP_OLD_SWEEP bits 1 / 2);
// what changes is the skeleton: the previous row lives in registers, the column's residue is a per-column register, the dense
// forward layer of the next step is fetched one step ahead with plain loads.
//   k_part_fwd_s<C>   forward recursion, writes Zm * 2^-e_row (f64) in slot layout and every row's exponent (rowexp)
//   k_part_rev_s<C>   reverse recursion fused with the posterior  P = Zf * Zr / (score * Z)  -> f32, no threshold (MSAPartProbs.cpp:286-297)
// States 0 = Zm, 1 = H (gap run along the row), 2 = V (gap run down the column), 3 = the row's scale exponent (exact small integer in a
// double), which travels with the row.  Sum orders: forward (Zm+V)+H (MSAPartProbs.cpp:583,589), reverse (Zm+H)+V (:283).
#include "posterior.cuh"
#include "sweep_c.cuh"
#ifndef MLP_MINB_PARTS
#define MLP_MINB_PARTS 4
#endif

__constant__ DevScalars c_sc_parts;

cudaError_t part_sc_set_scalars(const DevScalars& s, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_sc_parts, &s, sizeof(DevScalars), 0, cudaMemcpyHostToDevice, st);
}

namespace {

__device__ __forceinline__ int next_task_c(const KArgs& a, int lane) {
    int ti = 0;
    if (lane == 0) ti = atomicAdd(a.counter, 1);
    return __shfl_sync(MLP_FULL, ti, 0) + a.task_begin;
}

__device__ __forceinline__ double pow2d(int k) { return __hiloint2double((1023 + k) << 20, 0); }   // |k| <= 1022
__device__ __forceinline__ int exp_of(double x) { return ((__double2hiint(x) >> 20) & 0x7ff) - 1023; }
__device__ __forceinline__ double pow2c(int k) { return pow2d(max(min(k, 1000), -1000)); }
#define MLP_EXP_NONE (-1000000)

template <int C>
struct PartFwdS {
    typedef double T;
    typedef double TIN;
    enum { NS = 4, NIN = 0, REV = 0, ROW_LO = 0, USES_S1 = 1 };
    __device__ __forceinline__ int row_residue(int i) const { return i - 1; }
    const double* sub; double* Z; int* rowexp; const uint8_t* s2; int L1, L2, W;
    int roff[C]; int jb;
    const double* srow; double zz; bool has_zz; int zexp; double o0, e0;
    // scale bookkeeping: seen_exp = largest true exponent in the row this lane finished last (this column block)
    int seen_exp, row_seen, gmax, e_prev, e_row, cb, cbi; double f;
    __device__ __forceinline__ void reset() { has_zz = false; zz = 0; zexp = 0; seen_exp = MLP_EXP_NONE; row_seen = MLP_EXP_NONE; gmax = MLP_EXP_NONE; e_prev = 0; e_row = 0; f = 1.0; cb = 0; cbi = 0; }
    __device__ __forceinline__ void step_sync() { gmax = __reduce_max_sync(MLP_FULL, seen_exp); }
    __device__ __forceinline__ double load_in(int, long long) const { return 0.0; }
    __device__ __forceinline__ void begin_block(int cb_, int cbi_, int jbase) {
        cb = cb_; cbi = cbi_; e_prev = 0; seen_exp = MLP_EXP_NONE; gmax = MLP_EXP_NONE; jb = jbase;
#pragma unroll
        for (int c = 0; c < C; ++c) { const int j = jbase + c; roff[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0; }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int) const { st[0] = st[1] = st[2] = st[3] = 0.0; }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = e[3] = 0.0; }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        srow = sub + r1 * 26; e_row = MLP_EXP_NONE; row_seen = MLP_EXP_NONE;
        o0 = (i == L1) ? 1.0 : c_sc_parts.go; e0 = (i == L1) ? 1.0 : c_sc_parts.ge;   // H-type gap is terminal (exp(0)) in the last row
    }
    __device__ __forceinline__ void cell(int c, int i, int j, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&)[1], T (&nw)[NS]) {
        double fz = 1.0, fc = 1.0;   // factors that bring the diagonal / left operands into this row's scale
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
            f = (d == 0) ? 1.0 : pow2c(d);
        }
        fz = f;
        if (origin && cbi > 0) {       // operands handed over by the previous column block carry that block's exponents
            fc = pow2c((int)carry[3] - e_row);
            fz = pow2c((int)diag[3] - e_row);
        }
        if (i == 0 || j == 0 || j > L2) {   // boundary: Zm(0,0)=1, H(0,j>=1)=1, V(i>=1,0)=1 (terminal gaps are exp(0))
            nw[0] = (i == 0 && j == 0) ? 1.0 : 0.0;
            nw[1] = (i == 0 && j >= 1 && j <= L2) ? 1.0 : 0.0;
            nw[2] = (j == 0 && i >= 1) ? pow2c(-e_row) : 0.0;
            Z[idx] = nw[0];
        } else {
            const double score = srow[roff[c]];
            double h = __dadd_rn(__dmul_rn(carry[0], o0), __dmul_rn(carry[1], e0));
            double v = __dadd_rn(__dmul_rn(old[0], c_sc_parts.go), __dmul_rn(old[2], c_sc_parts.ge));
            if (j == L2) v = __dadd_rn(old[0], old[2]);   // V-type gap is terminal in the last column: x * exp(0) == x
            double zm = __dmul_rn(__dadd_rn(__dadd_rn(diag[0], diag[2]), diag[1]), score);   // (Zm+V)+H, MSAPartProbs.cpp:583
            if (fc != 1.0) h = __dmul_rn(h, fc);          // power-of-two rescaling: exact
            if (f != 1.0) v = __dmul_rn(v, f);
            if (fz != 1.0) zm = __dmul_rn(zm, fz);
            nw[0] = zm; nw[1] = h; nw[2] = v;
            Z[idx] = zm;
            row_seen = max(row_seen, exp_of(zm) + e_row);
            if (i == L1 && j == L2) { has_zz = true; zz = __dadd_rn(__dadd_rn(zm, v), h); zexp = e_row; }   // :589
        }
    }
    __device__ __forceinline__ void end_row(int, int, const T (&)[C][NS], T (&)[NS]) { e_prev = e_row; if (row_seen != MLP_EXP_NONE) seen_exp = row_seen; }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_PARTS) k_part_fwd_s(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* sub = reinterpret_cast<double*>(smem);
    for (int k = threadIdx.x; k < 676; k += blockDim.x) sub[k] = a.sub[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    double* edge = a.edge_d ? a.edge_d + gw * a.edge_stride : nullptr;
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = a.residues + a.seq_off[t.a]; cx.s2 = a.residues + a.seq_off[t.b];
        cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        PartFwdS<C> m;
        m.reset();
        m.sub = sub; m.Z = a.layerZ + t.off; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2; m.W = 32 * C;
        m.rowexp = a.rowexp + (long long)ti * a.rowexp_stride;
        run_sweep_c<PartFwdS<C>, C>(m, cx, edge, smem + MLP_PART_TABLE_BYTES + warp * MLP_SWEEP_RING_BYTES(4, 8));
        if (m.has_zz) { a.pout[ti].Zpart = m.zz; a.pout[ti].zexp = m.zexp; }
    }
}

template <int C>
struct PartRevS {
    typedef double T;
    typedef double TIN;
    enum { NS = 4, NIN = 1, REV = 1, ROW_LO = 0, USES_S1 = 1 };
    __device__ __forceinline__ int row_residue(int i) const { return i - 1; }
    const double* sub; const double* Z; float* P; const int* rowexp; const uint8_t* s2; int L1, L2, W, nb; double Ztot; int zexp;
    int roff[C];
    const double* srow; double o0, e0;
    int seen_exp, row_seen, gmax, e_prev, e_row, fexp, cb, cbi; double f;
    __device__ __forceinline__ void reset() { seen_exp = MLP_EXP_NONE; row_seen = MLP_EXP_NONE; gmax = MLP_EXP_NONE; e_prev = 0; e_row = 0; f = 1.0; fexp = 0; cb = 0; cbi = 0; }
    __device__ __forceinline__ void step_sync() { gmax = __reduce_max_sync(MLP_FULL, seen_exp); }
    __device__ __forceinline__ double load_in(int, long long idx) const { return Z[idx]; }
    __device__ __forceinline__ void begin_block(int cb_, int cbi_, int jbase) {
        cb = cb_; cbi = cbi_; e_prev = 0; seen_exp = MLP_EXP_NONE; gmax = MLP_EXP_NONE;
#pragma unroll
        for (int c = 0; c < C; ++c) { const int j = jbase + c; roff[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0; }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int j) const {   // virtual row L1+1 (scale exponent 0)
        st[0] = (j == L2 + 1) ? 1.0 : 0.0;
        st[1] = (j >= 1 && j <= L2) ? 1.0 : 0.0;
        st[2] = 0.0; st[3] = 0.0;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int i) const {    // virtual column L2+1 when it lies outside the strips
        e[0] = 0.0; e[1] = 0.0; e[2] = (i >= 1 && i <= L1) ? 1.0 : 0.0; e[3] = 0.0;
    }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        srow = sub + r1 * 26; e_row = MLP_EXP_NONE; row_seen = MLP_EXP_NONE;
        o0 = (i == 1) ? 1.0 : c_sc_parts.go; e0 = (i == 1) ? 1.0 : c_sc_parts.ge;   // H-type terminal at the first row
        fexp = rowexp[cb * (L1 + 1) + i];
    }
    __device__ __forceinline__ void cell(int c, int i, int j, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&in)[1], T (&nw)[NS]) {
        double fz = 1.0, fc = 1.0;
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
            f = (d == 0) ? 1.0 : pow2c(d);
        }
        fz = f;
        if (origin && cbi > 0) {
            fc = pow2c((int)carry[3] - e_row);
            fz = pow2c((int)diag[3] - e_row);
        }
        if (j > L2) {
            nw[0] = 0.0; nw[1] = 0.0;
            nw[2] = (j == L2 + 1 && i >= 1) ? pow2c(-e_row) : 0.0;
            return;
        }
        if (i == 0 || j == 0) { nw[0] = 0.0; nw[1] = 0.0; nw[2] = 0.0; P[idx] = 0.0f; return; }
        const double score = srow[roff[c]];
        double v = __dadd_rn(__dmul_rn(old[0], c_sc_parts.go), __dmul_rn(old[2], c_sc_parts.ge));
        if (j == 1) v = __dadd_rn(old[0], old[2]);   // V-type terminal at the first column
        double h = __dadd_rn(__dmul_rn(carry[0], o0), __dmul_rn(carry[1], e0));
        double zm = __dmul_rn(__dadd_rn(__dadd_rn(diag[0], diag[1]), diag[2]), score);   // (Zm+H)+V, MSAPartProbs.cpp:283
        if (fc != 1.0) h = __dmul_rn(h, fc);
        if (f != 1.0) v = __dmul_rn(v, f);
        if (fz != 1.0) zm = __dmul_rn(zm, fz);
        nw[0] = zm; nw[1] = h; nw[2] = v;
        row_seen = max(row_seen, exp_of(zm) + e_row);
        // MSAPartProbs.cpp:286-297: posterior = Zf * Zr / (score * Z), no threshold (the 0.001 filter is commented out there)
        // The scale shift k is applied with one exact power-of-two multiply instead of scalbn: |k| > 1000 would mean a posterior
        // below 2^-800 (the scaled operands stay within 2^+-200 of one another), where the clamped factor gives the same +0 as a
        // float; a product in the double-denormal range rounds once either way.  (Skipping the division where the float result
        // underflows was measured: c_p_np_aln's unfiltered posterior is non-zero almost everywhere, the test cost more than it saved.)
        const double q = __ddiv_rn(__dmul_rn(in[0], zm), __dmul_rn(score, Ztot));
        P[idx] = (float)__dmul_rn(q, pow2c(fexp + e_row - zexp));
    }
    __device__ __forceinline__ void end_row(int, int, const T (&)[C][NS], T (&)[NS]) { e_prev = e_row; if (row_seen != MLP_EXP_NONE) seen_exp = row_seen; }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_PARTS) k_part_rev_s(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* sub = reinterpret_cast<double*>(smem);
    for (int k = threadIdx.x; k < 676; k += blockDim.x) sub[k] = a.sub[k];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    double* edge = a.edge_d ? a.edge_d + gw * a.edge_stride : nullptr;
    for (;;) {
        const int ti = next_task_c(a, lane);
        if (ti >= a.ntasks) break;
        const PairTask t = a.tasks[ti];
        SweepCtx2 cx;
        cx.s1 = a.residues + a.seq_off[t.a]; cx.s2 = a.residues + a.seq_off[t.b];
        cx.lane = lane; cx.L1 = t.L1; cx.L2 = t.L2; cx.nb = t.nb;
        PartRevS<C> m;
        m.reset();
        m.sub = sub; m.Z = a.layerZ + t.off; m.P = a.layerP + t.off; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2; m.W = 32 * C; m.nb = t.nb;
        m.Ztot = a.pout[ti].Zpart; m.zexp = a.pout[ti].zexp;
        m.rowexp = a.rowexp + (long long)ti * a.rowexp_stride;
        run_sweep_c<PartRevS<C>, C>(m, cx, edge, smem + MLP_PART_TABLE_BYTES + warp * MLP_SWEEP_RING_BYTES(4, 8));
    }
}

typedef void (*KFn)(KArgs);
template <int C> KFn pick(int kernel) { return kernel == MLP_K_PART_FWD ? (KFn)k_part_fwd_s<C> : (KFn)k_part_rev_s<C>; }

}  // namespace

void (*part_sc_kernel(int kernel, int C))(KArgs) {
    switch (C) {
        case 1: return pick<1>(kernel); case 2: return pick<2>(kernel); case 3: return pick<3>(kernel); case 4: return pick<4>(kernel);
        case 5: return pick<5>(kernel); case 6: return pick<6>(kernel); case 7: return pick<7>(kernel); case 8: return pick<8>(kernel);
    }
    return nullptr;
}
