// Device-side arithmetic contract: FP32 log-space with the reference's piecewise
// polynomial log(1+e^x), the double-precision EXP polynomial, and NO fused multiply-add anywhere
// (the reference x86-64 binaries contain no FMA; this file is compiled with -fmad=false and uses the
// explicit _rn intrinsics where the order of roundings is part of the result).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MLP_LOG_ZERO (-2e20f)
#define MLP_FULL 0xffffffffu

// Transition/initial scalars live in constant memory: every lane reads the same address, so they become
// c[bank][off] operands of the FADDs (no load instruction).
struct DevScalars {
    float init[5];
    float t0q[5];   // trans[0][q]
    float tqq[5];   // trans[q][q]
    float tq0[5];   // trans[q][0]
    float lt00, lt01, lt02, lt10, lt11, lt20, lt22;   // local 3-state transitions
    float r, r2;                                       // random_transProb[1], 2*random_transProb[1]
    double go, ge;                                     // partition gap open / extend (terminal = 1)
};

// LOOKUP: cpnp ScoreType.h:198-216 == QP ScoreType.h:200-209. Coefficient selection by value, Horner with
// separate roundings.
__device__ __forceinline__ float dev_lookup(float x) {
    float a, b, c, d;
    if (x <= 1.00f) { a = -0.009350833524763f; b = 0.130659527668286f; c = 0.498799810682272f; d = 0.693203116424741f; }
    else if (x <= 2.50f) { a = -0.014532321752540f; b = 0.139942324101744f; c = 0.495635523139337f; d = 0.692140569840976f; }
    else if (x <= 4.50f) { a = -0.004605031767994f; b = 0.063427417320019f; c = 0.695956496475118f; d = 0.514272634594009f; }
    else { a = -0.000458661602210f; b = 0.009695946122598f; c = 0.930734667215156f; d = 0.168037164329057f; }
    float r = __fadd_rn(__fmul_rn(a, x), b);
    r = __fadd_rn(__fmul_rn(r, x), c);
    r = __fadd_rn(__fmul_rn(r, x), d);
    return r;
}

// LOG_ADD: ScoreType.h:279-285. Branch-free restatement: with d = max-min, the reference returns max when
// d >= 7.5 (or when the smaller operand is LOG_ZERO, which implies d >= 7.5 or both equal LOG_ZERO, where
// LOOKUP(0)+LOG_ZERO == LOG_ZERO), else LOOKUP(d) + min.
__device__ __forceinline__ float dev_log_add(float x, float y) {
    const float mx = fmaxf(x, y);
    const float mn = fminf(x, y);
    const float d = __fsub_rn(mx, mn);
    const float r = __fadd_rn(dev_lookup(d), mn);   // garbage (possibly inf) when d >= 7.5, discarded below
    return (d >= 7.5f) ? mx : r;
}

// Table-driven LOG_ADD.  The four cubic pieces of LOOKUP (break points 1, 2.5, 4.5, all multiples of 0.5) are laid out
// as 16 entries indexed by ceil(2d): entry t covers d in ((t-1)/2, t/2], so the reference's `x <= 1.0f`, `<= 2.5f`,
// `<= 4.5f` tests (upper bound inclusive) are reproduced exactly.  ceil(2d) comes out of one FFMA with round-up:
// 2d + 2^23 leaves ceil(2d) in the low mantissa bits.  Two conflict-free LDS.64 (a 16 x 8-byte table spans the 32
// banks exactly once) replace 3 compares + 12 selects on the half-rate ALU pipe.  Arithmetic on the coefficients is
// the same separate FMUL/FADD chain, so results are bit-identical to dev_log_add.
struct LogAddLut { float2 ab[16]; float2 cd[16]; };
__device__ __forceinline__ void log_add_lut_fill(LogAddLut* lut, int tid) {
    if (tid < 16) {
        float a, b, c, d;
        if (tid <= 2) { a = -0.009350833524763f; b = 0.130659527668286f; c = 0.498799810682272f; d = 0.693203116424741f; }
        else if (tid <= 5) { a = -0.014532321752540f; b = 0.139942324101744f; c = 0.495635523139337f; d = 0.692140569840976f; }
        else if (tid <= 9) { a = -0.004605031767994f; b = 0.063427417320019f; c = 0.695956496475118f; d = 0.514272634594009f; }
        else { a = -0.000458661602210f; b = 0.009695946122598f; c = 0.930734667215156f; d = 0.168037164329057f; }
        lut->ab[tid] = make_float2(a, b);
        lut->cd[tid] = make_float2(c, d);
    }
}
__device__ __forceinline__ float dev_log_add_lut(float x, float y, const LogAddLut* __restrict__ lut) {
    const float mx = fmaxf(x, y);
    const float mn = fminf(x, y);
    const float d = __fsub_rn(mx, mn);
    const unsigned idx = __float_as_uint(__fmaf_ru(d, 2.0f, 8388608.0f)) & 15u;
    const float2 ab = lut->ab[idx];
    const float2 cd = lut->cd[idx];
    float r = __fadd_rn(__fmul_rn(ab.x, d), ab.y);
    r = __fadd_rn(__fmul_rn(r, d), cd.x);
    r = __fadd_rn(__fmul_rn(r, d), cd.y);
    r = __fadd_rn(r, mn);
    return (d >= 7.5f) ? mx : r;
}

// Same function with the table addressed through a pre-biased 32-bit shared address (log_add_lut_bias): d is clamped to 7.5
// so ceil(2d) <= 15 needs no mask, and 8 * ceil(2d) is (u << 3) - (0x4B000000 << 3) in modular arithmetic -- the subtraction
// is folded into the loop-invariant base, leaving FMNMX + FFMA + LEA + 2 LDS.64 (one instruction less than mask + shift + add).
// For d >= 7.5 the polynomial value is discarded, as in dev_log_add.
__device__ __forceinline__ unsigned log_add_lut_bias(const LogAddLut* lut) {
    // The value travels through a shared-memory word: ptxas would otherwise fold "base - constant" back into every address
    // computation (one extra integer add per LOG_ADD) instead of keeping the biased base in a register.
    __shared__ unsigned s_bias;
    if (threadIdx.x == 0) s_bias = (unsigned)__cvta_generic_to_shared(lut) - 0x58000000u;
    __syncthreads();
    return *reinterpret_cast<volatile unsigned*>(&s_bias);
}
__device__ __forceinline__ float dev_log_add_lutb(float x, float y, unsigned lutb) {
    const float mx = fmaxf(x, y);
    const float mn = fminf(x, y);
    const float d = __fsub_rn(mx, mn);
    const unsigned u = __float_as_uint(__fmaf_ru(fminf(d, 7.5f), 2.0f, 8388608.0f));
    const unsigned sa = lutb + (u << 3);
    float a, b, c, e;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(sa));
    asm("ld.shared.v2.f32 {%0, %1}, [%2+128];" : "=f"(c), "=f"(e) : "r"(sa));
    float r = __fadd_rn(__fmul_rn(a, d), b);
    r = __fadd_rn(__fmul_rn(r, d), c);
    r = __fadd_rn(__fmul_rn(r, d), e);
    r = __fadd_rn(r, mn);
    return (d >= 7.5f) ? mx : r;
}

// cp.async (LDGSTS) helpers: stage the next wavefront slot of a dense layer into shared memory while the
// current row is being computed.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// EXP: ScoreType.h:36-68. Double Horner chain on the promoted float, rounded to float once. Callers clamp x <= 0.
__device__ __forceinline__ float dev_exp(float xf) {
    if (!(xf > -16.0f)) return 0.0f;
    const double x = (double)xf;
    double c4, c3, c2, c1, c0;
    if (xf > -2.0f) {
        if (xf > -0.5f) { c4 = 0.03254409303190190000; c3 = 0.16280432765779600000; c2 = 0.49929760485974900000; c1 = 0.99995149601363700000; c0 = 0.99999925508501600000; }
        else if (xf > -1.0f) { c4 = 0.01973899026052090000; c3 = 0.13822379685007000000; c2 = 0.48056651562365000000; c1 = 0.99326940370383500000; c0 = 0.99906756856399500000; }
        else { c4 = 0.00940528203591384000; c3 = 0.09414963667859410000; c2 = 0.40825793595877300000; c1 = 0.93933625499130400000; c0 = 0.98369508190545300000; }
    } else if (xf > -8.0f) {
        if (xf > -4.0f) { c4 = 0.00217245711583303000; c3 = 0.03484829428350620000; c2 = 0.22118199801337800000; c1 = 0.67049462206469500000; c0 = 0.83556950223398500000; }
        else { c4 = 0.00012398771025456900; c3 = 0.00349155785951272000; c2 = 0.03727721426017900000; c1 = 0.17974997741536900000; c0 = 0.33249299994217400000; }
    } else { c4 = 0.00000051741713416603; c3 = 0.00002721456879608080; c2 = 0.00053418601865636800; c1 = 0.00464101989351936000; c0 = 0.01507447981459420000; }
    double r = __dadd_rn(__dmul_rn(c4, x), c3);
    r = __dadd_rn(__dmul_rn(r, x), c2);
    r = __dadd_rn(__dmul_rn(r, x), c1);
    r = __dadd_rn(__dmul_rn(r, x), c0);
    return (float)r;
}

// Table-driven EXP: the six ranges of ScoreType.h:36-68 are (-2^k, -2^(k-1)] for k = -1..4 plus (-0.5, 0], so the piece index
// is the clamped binary exponent of |x| + 2 (x = -0.5 -> piece 1, x = -1 -> piece 2, ... exactly the reference's strict '>'
// tests). Coefficients (doubles) come from shared memory; same double Horner chain, one rounding to float.
struct ExpLut { double c[6][6]; };   // c[piece][0..4] = c4..c0, [5] pad
__device__ __forceinline__ void exp_lut_fill(ExpLut* lut, int tid) {
    if (tid < 6) {
        const double t[6][5] = {
            {0.03254409303190190000, 0.16280432765779600000, 0.49929760485974900000, 0.99995149601363700000, 0.99999925508501600000},
            {0.01973899026052090000, 0.13822379685007000000, 0.48056651562365000000, 0.99326940370383500000, 0.99906756856399500000},
            {0.00940528203591384000, 0.09414963667859410000, 0.40825793595877300000, 0.93933625499130400000, 0.98369508190545300000},
            {0.00217245711583303000, 0.03484829428350620000, 0.22118199801337800000, 0.67049462206469500000, 0.83556950223398500000},
            {0.00012398771025456900, 0.00349155785951272000, 0.03727721426017900000, 0.17974997741536900000, 0.33249299994217400000},
            {0.00000051741713416603, 0.00002721456879608080, 0.00053418601865636800, 0.00464101989351936000, 0.01507447981459420000}};
        for (int k = 0; k < 5; ++k) lut->c[tid][k] = t[tid][k];
        lut->c[tid][5] = 0.0;
    }
}
__device__ __forceinline__ float dev_exp_lut(float xf, const ExpLut* __restrict__ lut) {
    // callers guarantee xf <= 0
    const int e = (int)((__float_as_uint(xf) >> 23) & 0xffu) - 127;        // floor(log2|x|), -127 for 0/denormals
    const int piece = min(max(e + 2, 0), 5);
    const double* c = lut->c[piece];
    const double x = (double)xf;
    double r = __dadd_rn(__dmul_rn(c[0], x), c[1]);
    r = __dadd_rn(__dmul_rn(r, x), c[2]);
    r = __dadd_rn(__dmul_rn(r, x), c[3]);
    r = __dadd_rn(__dmul_rn(r, x), c[4]);
    return (xf > -16.0f) ? (float)r : 0.0f;
}

// posterior cell: EXP(min(LOG_ONE, F+B-total)) with s = F+B already rounded (ProbabilisticModel.h:483)
__device__ __forceinline__ float dev_posterior_from_sum(float s, float total) {
    return dev_exp(fminf(0.0f, __fsub_rn(s, total)));
}

// QuickProbs' uint16 fixed-point cell (SparseEntry.h:31-32): store (uint16)(v*65535) truncating, load (float)u/65535.
__device__ __forceinline__ float dev_quantize_u16(float v) {
    const unsigned code = __float2uint_rz(__fmul_rn(v, 65535.0f)) & 0xffffu;
    return __fdiv_rn((float)code, 65535.0f);
}

// One pair inside a batch. Dense DP layers use the "slot" layout [cb][t][c][lane]:
// column block cb, wavefront slot t = row + lane, column-in-strip c, lane -> one coalesced 128-byte line per c.
struct PairTask {
    int a, b;          // sequence indices, a < b
    int L1, L2;        // lengths (rows = seq a, columns = seq b)
    int C, nb;         // columns per lane, column blocks (32*C columns each)
    int pidx;          // global pair index (row-major a<b enumeration)
    int flags;
    long long off;     // element offset of this pair inside every dense layer
};

__device__ __forceinline__ long long task_layer_elems(const PairTask& t) {
    return (long long)t.nb * (t.L1 + 32) * t.C * 32;
}
