// QuickProbs partition-function posterior (PartitionFunction.cpp:71-291), register-band sweeps (sweep_c.cuh).
//   k_part_fwd_c<C>   forward recursion, writes Zm (f64) of rows 1..L1 in slot layout
//   k_part_rev_c<C>   reverse recursion fused with the posterior  P = Zf * Zr / (score * Z)  (f32, thresholded at 0.001)
// The cpnp flavour (80-bit reference, rescaled FP64 here) stays on k_part_*_t<true> in posterior.cu.
//
// States 0 = Zm, 1 = H (gap run along the row, "Ze"), 2 = V (gap run down the column, "Zf").  What used to be per-cell
// boundary tests is data here:
//   * a terminal gap costs exp(0) = 1: the reference's `Zm + Zf` is Zm*1 + Zf*1 exactly, so the last/first column (and the
//     last/first row) just use open = extend = 1 -- per-column factors gov[c], gev[c] set once per column block, per-row
//     factors o0, e0 set once per row;
//   * column 0 of the forward pass (Zm = H = 0, V = 1) and the virtual column L2+1 of the reverse pass (same values) are
//     "terminal" columns fed by zeros: Zm = 0 * score and H = 0 come out of the ordinary arithmetic, V = 1 out of the
//     terminal factors (1*1 + 0*1, then 0*1 + 1*1, ...);
//   * row 0 / row L1+1 are virtual (band_init), never computed or stored; padding columns beyond L2(+1) stay zero in the
//     reverse pass (everything that flows into them is zero) and are garbage nobody reads in the forward pass.
#include "posterior.cuh"
#include "sweep_c.cuh"
#ifndef MLP_MINB_PART
#define MLP_MINB_PART 5   // minimum resident CTAs per SM the register allocation is held to (measured)
#endif

__constant__ DevScalars c_sc_part;

cudaError_t part_c_set_scalars(const DevScalars& s, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_sc_part, &s, sizeof(DevScalars), 0, cudaMemcpyHostToDevice, st);
}

namespace {

__device__ __forceinline__ int next_task_c(const KArgs& a, int lane) {
    int ti = 0;
    if (lane == 0) ti = atomicAdd(a.counter, 1);
    return __shfl_sync(MLP_FULL, ti, 0) + a.task_begin;
}

template <int C>
struct PartFwdQ {
    typedef double T;
    typedef double TIN;
    enum { NS = 3, NIN = 0, REV = 0, ROW_LO = 1, USES_S1 = 1 };
    __device__ __forceinline__ int row_residue(int i) const { return i - 1; }
    const double* sub; double* Z; const uint8_t* s1; const uint8_t* s2; int L1, L2;
    double go, ge;
    int roff[C]; double gov[C], gev[C];
    const double* srow; double o0, e0;
    double zz; bool has_zz;
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ double load_in(int, long long) const { return 0.0; }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            roff[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0;
            const bool term = (j == 0 || j == L2);
            gov[c] = term ? 1.0 : go; gev[c] = term ? 1.0 : ge;
        }
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int j) const {   // row 0: Zm(0,0) = 1, H(0,j>=1) = 1
        st[0] = (j == 0) ? 1.0 : 0.0; st[1] = (j >= 1 && j <= L2) ? 1.0 : 0.0; st[2] = 0.0;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const { e[0] = e[1] = e[2] = 0.0; }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        srow = sub + r1 * 26;
        o0 = (i == L1) ? 1.0 : go; e0 = (i == L1) ? 1.0 : ge;      // H-type gap is terminal in the last row
    }
    __device__ __forceinline__ void cell(int c, int, int, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&)[1], T (&nw)[NS]) {
        const double score = srow[roff[c]];
        const double h = __dadd_rn(__dmul_rn(carry[0], o0), __dmul_rn(carry[1], e0));
        const double v = __dadd_rn(__dmul_rn(old[0], gov[c]), __dmul_rn(old[2], gev[c]));
        const double zm = __dmul_rn(__dadd_rn(__dadd_rn(diag[0], diag[1]), diag[2]), score);   // (Zm+H)+V, PartitionFunction.cpp:137
        nw[0] = zm; nw[1] = h; nw[2] = v;
#ifndef MLP_EXP_NOSTORE
        Z[idx] = zm;
#else
        if (zm == -1.2345) Z[idx] = zm;   // timing experiment only
#endif
    }
    __device__ __forceinline__ void end_row(int i, int jbase, const T (&band)[C][NS], T (&)[NS]) {
        if (i == L1) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (jbase + c == L2) { has_zz = true; zz = __dadd_rn(__dadd_rn(band[c][0], band[c][1]), band[c][2]); }   // :139
        }
    }
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_PART) k_part_fwd_c(KArgs a) {
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
        PartFwdQ<C> m;
        m.sub = sub; m.Z = a.layerZ + t.off; m.s1 = cx.s1; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2;
        m.go = c_sc_part.go; m.ge = c_sc_part.ge; m.has_zz = false; m.zz = 0.0;
        run_sweep_c<PartFwdQ<C>, C>(m, cx, edge, smem + MLP_PART_TABLE_BYTES + warp * MLP_SWEEP_RING_BYTES(3, 8));
        if (m.has_zz) { a.pout[ti].Zpart = m.zz; a.pout[ti].zexp = 0; }
    }
}

template <int C>
struct PartRevQ {
    typedef double T;
    typedef double TIN;
    enum { NS = 3, NIN = 1, REV = 1, ROW_LO = 1, USES_S1 = 1 };
    __device__ __forceinline__ int row_residue(int i) const { return i - 1; }
    const double* sub; const double* Z; float* P; const uint8_t* s1; const uint8_t* s2; int L1, L2, ncols;
    double go, ge, Ztot;
    int roff[C]; double gov[C], gev[C];
    bool col0;                    // this lane's first column is column 0 (its posterior is forced to 0)
    const double* srow; double o0, e0;
    __device__ __forceinline__ void step_sync() const {}
    __device__ __forceinline__ double load_in(int, long long idx) const { return Z[idx]; }
    __device__ __forceinline__ void begin_block(int, int, int jbase) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c;
            roff[c] = (j >= 1 && j <= L2) ? s2[j - 1] : 0;
            const bool term = (j == 1 || j == L2 + 1);
            gov[c] = term ? 1.0 : go; gev[c] = term ? 1.0 : ge;
        }
        col0 = (jbase == 0);
    }
    __device__ __forceinline__ void band_init(T (&st)[NS], int j) const {   // virtual row L1+1
        st[0] = (j == L2 + 1) ? 1.0 : 0.0; st[1] = (j >= 1 && j <= L2) ? 1.0 : 0.0; st[2] = 0.0;
    }
    __device__ __forceinline__ void edge_init(T (&e)[NS], int) const {      // column 32*C*nb: the virtual column L2+1 only if it lies outside the strips
        e[0] = 0.0; e[1] = 0.0; e[2] = (ncols == L2 + 1) ? 1.0 : 0.0;
    }
    __device__ __forceinline__ void begin_row(int i, int r1) {
        srow = sub + r1 * 26;
        o0 = (i == 1) ? 1.0 : go; e0 = (i == 1) ? 1.0 : ge;        // H-type gap is terminal at the first row
    }
    __device__ __forceinline__ void cell(int c, int, int, long long idx, const T (&old)[NS], const T (&carry)[NS], const T (&diag)[NS],
                                         const TIN (&in)[1], T (&nw)[NS]) {
        const double score = srow[roff[c]];
        const double v = __dadd_rn(__dmul_rn(old[0], gov[c]), __dmul_rn(old[2], gev[c]));
        const double h = __dadd_rn(__dmul_rn(carry[0], o0), __dmul_rn(carry[1], e0));
        const double zm = __dmul_rn(__dadd_rn(__dadd_rn(diag[0], diag[2]), diag[1]), score);   // (Zm+V)+H, PartitionFunction.cpp:257
        nw[0] = zm; nw[1] = h; nw[2] = v;
        // PartitionFunction.cpp:259-270: posterior = Zf * Zr / (score * Z), kept only inside [0.001, 1].
        // `(double)p >= 0.001` on a float p is `p >= 0.001f` (0.001f is the smallest float not below the double 0.001).
        const double num = __dmul_rn(in[0], zm);
        const double den = __dmul_rn(score, Ztot);
        // Cells far below the threshold (num < 0.0009 * den: the quotient is < 0.00091 whatever the rounding) are 0 without the
        // division; the warp only runs the division when one of its lanes needs it.  NaN / inf operands: a NaN numerator fails
        // the test and gives 0, as the reference's `!(p <= 1 && p >= 0.001)` does; inf / finite passes and is divided.
        const bool need = num > __dmul_rn(den, 0.0009);
        float p = 0.0f;
        if (__any_sync(__activemask(), need)) {
            const float q = (float)__ddiv_rn(num, den);
            p = (need && q <= 1.0f && q >= 0.001f) ? q : 0.0f;
        }
        if (c == 0 && col0) p = 0.0f;
        P[idx] = p;
    }
    __device__ __forceinline__ void end_row(int, int, const T (&)[C][NS], T (&)[NS]) const {}
};

template <int C>
__global__ void __launch_bounds__(MLP_BLOCK, MLP_MINB_PART) k_part_rev_c(KArgs a) {
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
        PartRevQ<C> m;
        m.sub = sub; m.Z = a.layerZ + t.off; m.P = a.layerP + t.off; m.s1 = cx.s1; m.s2 = cx.s2; m.L1 = t.L1; m.L2 = t.L2;
        m.ncols = t.nb * 32 * C;
        m.go = c_sc_part.go; m.ge = c_sc_part.ge; m.Ztot = a.pout[ti].Zpart;
        // row 0 of the posterior is 0 (PartitionFunction.cpp:243) and is not part of the sweep: slot of (0, lane's columns) = lane
        for (int cb = 0; cb < t.nb; ++cb)
#pragma unroll
            for (int c = 0; c < C; ++c) m.P[((long long)(cb * (t.L1 + 32) + lane) * C + c) * 32 + lane] = 0.0f;
        run_sweep_c<PartRevQ<C>, C>(m, cx, edge, smem + MLP_PART_TABLE_BYTES + warp * MLP_SWEEP_RING_BYTES(3, 8));
    }
}

template <int C> void set_attr(size_t smem) {
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(k_part_fwd_c<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_part_rev_c<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
}

typedef void (*KFn)(KArgs);
template <int C> KFn pick(int kernel) { return kernel == MLP_K_PART_FWD ? (KFn)k_part_fwd_c<C> : (KFn)k_part_rev_c<C>; }

}  // namespace

void (*part_c_kernel(int kernel, int C))(KArgs) {
    switch (C) {
        case 1: return pick<1>(kernel); case 2: return pick<2>(kernel); case 3: return pick<3>(kernel); case 4: return pick<4>(kernel);
        case 5: return pick<5>(kernel); case 6: return pick<6>(kernel); case 7: return pick<7>(kernel); case 8: return pick<8>(kernel);
    }
    return nullptr;
}
