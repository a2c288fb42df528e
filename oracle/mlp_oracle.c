/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See mlp_oracle.h.
 * Plain-C restatement of the reference arithmetic (FP32 log-space pair-HMMs, FP64 / 80-bit
 * partition function, sparse threshold, uint16 quantisation, consistency relaxation).
 * Built with -ffp-contract=off: the reference binaries contain no FMA.
 */
#include "mlp_oracle.h"
#include "param_data.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG_ZERO (-2e20f)
#define LOG_ONE 0.0f
#define UNDERFLOW 7.5f

/* ---------------------------------------------------------------- math (ScoreType.h) */
/* cpnp ScoreType.h:198-216 == QP ScoreType.h:200-209 (LOOKUP_FLOAT) */
static inline float lookup(float x) {
    if (x <= 1.00f) return ((-0.009350833524763f * x + 0.130659527668286f) * x + 0.498799810682272f) * x + 0.693203116424741f;
    if (x <= 2.50f) return ((-0.014532321752540f * x + 0.139942324101744f) * x + 0.495635523139337f) * x + 0.692140569840976f;
    if (x <= 4.50f) return ((-0.004605031767994f * x + 0.063427417320019f) * x + 0.695956496475118f) * x + 0.514272634594009f;
    return ((-0.000458661602210f * x + 0.009695946122598f) * x + 0.930734667215156f) * x + 0.168037164329057f;
}
/* ScoreType.h:279-285 */
float orc_log_add(float x, float y) {
    if (x < y) return (x == LOG_ZERO || y - x >= UNDERFLOW) ? y : lookup(y - x) + x;
    return (y == LOG_ZERO || x - y >= UNDERFLOW) ? x : lookup(x - y) + y;
}
/* ScoreType.h:251-258 */
static inline void lpe(float* x, float y) { *x = orc_log_add(*x, y); }
/* ScoreType.h:36-68: double Horner chain on the promoted float, one rounding to float */
float orc_exp(float xf) {
    double x = xf;
    if (x > -2) {
        if (x > -0.5) {
            if (x > 0) return (float)exp(x);
            return (float)((((0.03254409303190190000 * x + 0.16280432765779600000) * x + 0.49929760485974900000) * x + 0.99995149601363700000) * x + 0.99999925508501600000);
        }
        if (x > -1) return (float)((((0.01973899026052090000 * x + 0.13822379685007000000) * x + 0.48056651562365000000) * x + 0.99326940370383500000) * x + 0.99906756856399500000);
        return (float)((((0.00940528203591384000 * x + 0.09414963667859410000) * x + 0.40825793595877300000) * x + 0.93933625499130400000) * x + 0.98369508190545300000);
    }
    if (x > -8) {
        if (x > -4) return (float)((((0.00217245711583303000 * x + 0.03484829428350620000) * x + 0.22118199801337800000) * x + 0.67049462206469500000) * x + 0.83556950223398500000);
        return (float)((((0.00012398771025456900 * x + 0.00349155785951272000) * x + 0.03727721426017900000) * x + 0.17974997741536900000) * x + 0.33249299994217400000);
    }
    if (x > -16) return (float)((((0.00000051741713416603 * x + 0.00002721456879608080) * x + 0.00053418601865636800) * x + 0.00464101989351936000) * x + 0.01507447981459420000);
    return 0;
}

/* ---------------------------------------------------------------- tables */
/* MSA.cpp:861-870 */
float orc_init_distrib2_for_identity(float identity) {
    if (identity <= 0.125) return 0.108854f;
    else if (identity <= 0.15) return 0.132548f;
    else if (identity <= 0.175) return 0.165248f;
    else if (identity <= 0.2) return 0.168284f;
    else if (identity <= 0.25) return 0.170705f;
    else if (identity <= 0.3) return 0.100675f;
    else if (identity <= 0.35) return 0.090755f;
    else if (identity <= 0.4) return 0.146188f;
    else if (identity <= 0.45) return 0.167858f;
    else if (identity <= 0.5) return 0.250769f;
    return 0.700645f;   /* Defaults.h:22-23 default */
}

/* ProbabilisticModel.h:58-135 (cpnp) ; PairHmm.cpp:4-33 + ProbabilisticModel.cpp:15-56 (QP) */
void orc_build_hmm(float init_distrib2, orc_hmm_tables* t) {
    /* Defaults.h:22-27 / ProteinHmm5.cpp:5-9 */
    const float initD[5] = {0.6814756989f, 8.615339902e-05f, 8.615339902e-05f, 0.1591759622f, 0.1591759622f};
    const float gapOpen[2] = {0.0119511066f, 0.008008334786f};
    const float gapExt[2] = {0.3965826333f, 0.8988758326f};
    const float lgo = 0.01993141696f, lge = 0.7943345308f;   /* gapOpen[1], gapExtend[1]: local model */
    float tm[5][5];
    memset(tm, 0, sizeof tm);
    tm[0][0] = 1;
    for (int i = 0; i < 2; i++) {
        tm[0][2 * i + 1] = gapOpen[i];
        tm[0][2 * i + 2] = gapOpen[i];
        tm[0][0] -= (gapOpen[i] + gapOpen[i]);
        tm[2 * i + 1][2 * i + 1] = gapExt[i];
        tm[2 * i + 2][2 * i + 2] = gapExt[i];
        tm[2 * i + 1][0] = 1 - gapExt[i];
        tm[2 * i + 2][0] = 1 - gapExt[i];
    }
    for (int i = 0; i < 5; i++) {
        t->init[i] = logf(initD[i]);
        for (int j = 0; j < 5; j++) t->trans[i][j] = logf(tm[i][j]);
    }
    const float e5 = 1e-5, e10 = 1e-10;
    for (int a = 0; a < 26; a++) {
        t->ins[a] = logf(e5);
        for (int b = 0; b < 26; b++) t->match[a][b] = logf(e10);
    }
    const char* al = MLP_HMM_ALPHABET;
    for (int i = 0; i < 20; i++) {
        t->ins[al[i] - 'A'] = logf(MLP_EMIT_SINGLE[i]);
        for (int j = 0; j <= i; j++) {
            float v = logf(MLP_EMIT_PAIRS_TRI[i * (i + 1) / 2 + j]);
            t->match[al[i] - 'A'][al[j] - 'A'] = v;
            t->match[al[j] - 'A'][al[i] - 'A'] = v;
        }
    }
    float lt[3][3];
    memset(lt, 0, sizeof lt);
    lt[0][0] = 1;
    lt[0][1] = lgo; lt[0][2] = lgo;
    lt[0][0] -= (lgo + lgo);
    lt[1][1] = lge; lt[2][2] = lge;
    lt[1][0] = 1 - lge; lt[2][0] = 1 - lge;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t->ltrans[i][j] = logf(lt[i][j]);
    t->rtrans[0] = logf(init_distrib2);
    t->rtrans[1] = logf(1 - init_distrib2);
}

/* MSAReadMatrix.cpp:85-116 (expf of a float product, widened), :158-210; MSAPartProbs.cpp:698-709 */
void orc_build_part_cpnp(orc_part_tables* t) {
    const float TEMPERATURE = 5;
    const float beta = (float)(1.0 / TEMPERATURE);
    for (int a = 0; a < 26; a++) for (int b = 0; b < 26; b++) t->sub[a][b] = NAN;
    const char* al = MLP_GONNET160_ALPHABET;
    int n = (int)strlen(al), pos = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double v = expf(beta * MLP_GONNET160_TRI[pos++]);
            t->sub[al[i] - 'A'][al[j] - 'A'] = v;
            t->sub[al[j] - 'A'][al[i] - 'A'] = v;
        }
    float gapopen = -22, gapext = -1;
    double b = beta;
    t->tgo = exp(b * 0.0);
    t->tge = exp(b * 0.0);
    t->go = exp(b * (double)gapopen);
    t->ge = exp(b * (double)gapext);
}

/* ExpPartitionFunctionParams.h:30-49 with Configuration.cpp:330-332 */
void orc_build_part_qp(orc_part_tables* t) {
    const double temperature = 5.6007, gi = -25.3549, ge = -1.30113;
    double beta = 1.0 / temperature;
    memset(t->sub, 0, sizeof t->sub);
    const char* al = MLP_VTML200_ALPHABET;
    int n = (int)strlen(al);
    for (int i = 0; i < n - 1; i++)
        for (int j = 0; j <= i; j++) {
            double v = exp(beta * MLP_VTML200[i * n + j]);
            t->sub[al[i] - 'A'][al[j] - 'A'] = v;
            t->sub[al[j] - 'A'][al[i] - 'A'] = v;
        }
    t->go = exp(beta * gi);
    t->ge = exp(beta * ge);
    t->tgo = exp(beta * 0);
    t->tge = exp(beta * 0);
}

/* ---------------------------------------------------------------- 5-state pair-HMM */
#define IX(i, j) ((size_t)(i) * W + (j))
#define R1(i) (s1[(i)-1] - 'A')   /* 1-based residue of seq1 */
#define R2(j) (s2[(j)-1] - 'A')

/* forward: ProbabilisticModel.h:153-274 (flag=true) == ParallelProbabilisticModel.cpp:40-131
 * backward: :292-395 == :152-234 ; total: :405-454 ; posterior: :464-493 == :240-269 */
void orc_hmm5_posterior(const orc_hmm_tables* t, const char* s1, int L1, const char* s2, int L2,
                        int qp_total_quirk, float* post, float* total_out) {
    const int W = L2 + 1;
    const size_t cells = (size_t)(L1 + 1) * W;
    float* F[5]; float* B[5];
    for (int k = 0; k < 5; k++) {
        F[k] = (float*)malloc(cells * sizeof(float));
        B[k] = (float*)malloc(cells * sizeof(float));
        for (size_t c = 0; c < cells; c++) { F[k][c] = LOG_ZERO; B[k][c] = LOG_ZERO; }
    }
    F[0][IX(1, 1)] = t->init[0] + t->match[R1(1)][R2(1)];
    for (int k = 0; k < 2; k++) {
        F[2 * k + 1][IX(1, 0)] = t->init[2 * k + 1] + t->ins[R1(1)];
        F[2 * k + 2][IX(0, 1)] = t->init[2 * k + 2] + t->ins[R2(1)];
    }
    for (int i = 0; i <= L1; i++)
        for (int j = 0; j <= L2; j++) {
            if (i > 1 || j > 1) {
                if (i > 0 && j > 0) {
                    float v = F[0][IX(i - 1, j - 1)] + t->trans[0][0];
                    for (int k = 1; k < 5; k++) lpe(&v, F[k][IX(i - 1, j - 1)] + t->trans[k][0]);
                    v += t->match[R1(i)][R2(j)];
                    F[0][IX(i, j)] = v;
                }
                if (i > 0)
                    for (int k = 0; k < 2; k++) {
                        int q = 2 * k + 1;
                        F[q][IX(i, j)] = t->ins[R1(i)] + orc_log_add(F[0][IX(i - 1, j)] + t->trans[0][q], F[q][IX(i - 1, j)] + t->trans[q][q]);
                    }
                if (j > 0)
                    for (int k = 0; k < 2; k++) {
                        int q = 2 * k + 2;
                        F[q][IX(i, j)] = t->ins[R2(j)] + orc_log_add(F[0][IX(i, j - 1)] + t->trans[0][q], F[q][IX(i, j - 1)] + t->trans[q][q]);
                    }
            }
        }
    for (int k = 0; k < 5; k++) B[k][IX(L1, L2)] = t->init[k];
    for (int i = L1; i >= 0; i--)
        for (int j = L2; j >= 0; j--) {
            if (i < L1 && j < L2) {
                const float pxy = B[0][IX(i + 1, j + 1)] + t->match[R1(i + 1)][R2(j + 1)];
                for (int k = 0; k < 5; k++) lpe(&B[k][IX(i, j)], pxy + t->trans[k][0]);
            }
            if (i < L1)
                for (int k = 0; k < 2; k++) {
                    int q = 2 * k + 1;
                    lpe(&B[0][IX(i, j)], B[q][IX(i + 1, j)] + t->ins[R1(i + 1)] + t->trans[0][q]);
                    lpe(&B[q][IX(i, j)], B[q][IX(i + 1, j)] + t->ins[R1(i + 1)] + t->trans[q][q]);
                }
            if (j < L2)
                for (int k = 0; k < 2; k++) {
                    int q = 2 * k + 2;
                    lpe(&B[0][IX(i, j)], B[q][IX(i, j + 1)] + t->ins[R2(j + 1)] + t->trans[0][q]);
                    lpe(&B[q][IX(i, j)], B[q][IX(i, j + 1)] + t->ins[R2(j + 1)] + t->trans[q][q]);
                }
        }
    float tF = LOG_ZERO;
    for (int k = 0; k < 5; k++) lpe(&tF, F[k][IX(L1, L2)] + B[k][IX(L1, L2)]);
    float tB = F[0][IX(1, 1)] + B[0][IX(1, 1)];
    for (int k = 0; k < 2; k++) {
        lpe(&tB, F[2 * k + 1][IX(1, 0)] + B[2 * k + 1][IX(1, 0)]);
        lpe(&tB, F[2 * k + 2][IX(0, 1)] + B[2 * k + 2][IX(0, 1)]);
    }
    float total = (tF + tB) / 2;
    if (qp_total_quirk && total == 0) total = 1.0f;   /* ParallelProbabilisticModel.cpp:252-254 */
    for (size_t c = 0; c < cells; c++) post[c] = orc_exp(fminf(LOG_ONE, F[0][c] + B[0][c] - total));
    post[0] = 0;
    if (total_out) *total_out = total;
    for (int k = 0; k < 5; k++) { free(F[k]); free(B[k]); }
}

/* ---------------------------------------------------------------- local 3-state pair-HMM (cpnp, flag=false) */
void orc_local_posterior(const orc_hmm_tables* t, const char* s1, int L1, const char* s2, int L2,
                         float* post, float* total_out) {
    const int W = L2 + 1;
    const size_t cells = (size_t)(L1 + 1) * W;
    float* F[3]; float* B[3];
    for (int k = 0; k < 3; k++) {
        F[k] = (float*)malloc(cells * sizeof(float));
        B[k] = (float*)malloc(cells * sizeof(float));
        for (size_t c = 0; c < cells; c++) { F[k][c] = LOG_ZERO; B[k][c] = LOG_ZERO; }
    }
    const float r = t->rtrans[1];
    const float r2 = 2 * r;
    /* forward ProbabilisticModel.h:205-262 */
    for (int i = 0; i <= L1; i++)
        for (int j = 0; j <= L2; j++) {
            if (i == 1 && j == 1) F[0][IX(i, j)] = t->match[R1(i)][R2(j)] - t->ins[R1(i)] - t->ins[R2(j)] - r2;
            if (i > 1 || j > 1) {
                if (i > 0 && j > 0) {
                    const float m = t->match[R1(i)][R2(j)], a = t->ins[R1(i)], b = t->ins[R2(j)];
                    float v = m - a - b - r2;
                    for (int k = 0; k < 3; k++) lpe(&v, m - a - b + F[k][IX(i - 1, j - 1)] + t->ltrans[k][0] - r2);
                    F[0][IX(i, j)] = v;
                }
                if (i > 0) F[1][IX(i, j)] = orc_log_add(F[0][IX(i - 1, j)] + t->ltrans[0][1] - r, F[1][IX(i - 1, j)] + t->ltrans[1][1] - r);
                if (j > 0) F[2][IX(i, j)] = orc_log_add(F[0][IX(i, j - 1)] + t->ltrans[0][2] - r, F[2][IX(i, j - 1)] + t->ltrans[2][2] - r);
            }
        }
    /* backward :335-381 */
    for (int i = L1; i >= 0; i--)
        for (int j = L2; j >= 0; j--) {
            B[0][IX(i, j)] = LOG_ONE;
            if (i < L1 && j < L2) {
                const float pxy = B[0][IX(i + 1, j + 1)] + t->match[R1(i + 1)][R2(j + 1)] - t->ins[R1(i + 1)] - t->ins[R2(j + 1)];
                for (int k = 0; k < 3; k++) lpe(&B[k][IX(i, j)], pxy + t->ltrans[k][0] - r2);
            }
            if (i < L1) {
                lpe(&B[0][IX(i, j)], B[1][IX(i + 1, j)] + t->ltrans[0][1] - r);
                lpe(&B[1][IX(i, j)], B[1][IX(i + 1, j)] + t->ltrans[1][1] - r);
            }
            if (j < L2) {
                lpe(&B[0][IX(i, j)], B[2][IX(i, j + 1)] + t->ltrans[0][2] - r);
                lpe(&B[2][IX(i, j)], B[2][IX(i, j + 1)] + t->ltrans[2][2] - r);
            }
        }
    /* total :434-453: sequential row-major chain */
    float tF = LOG_ZERO, tB = LOG_ZERO;
    for (int i = 1; i <= L1; i++)
        for (int j = 1; j <= L2; j++) {
            lpe(&tF, F[0][IX(i, j)]);
            lpe(&tB, B[0][IX(i, j)] + t->match[R1(i)][R2(j)] - t->ins[R1(i)] - t->ins[R2(j)] - r2);
        }
    float total = (tF + tB) / 2;
    for (size_t c = 0; c < cells; c++) post[c] = orc_exp(fminf(LOG_ONE, F[0][c] + B[0][c] - total));
    post[0] = 0;
    if (total_out) *total_out = total;
    for (int k = 0; k < 3; k++) { free(F[k]); free(B[k]); }
}

/* ---------------------------------------------------------------- Viterbi + model selection */
/* ProbabilisticModel.h:1043-1170 */
float orc_viterbi(const orc_hmm_tables* t, const char* s1, int L1, const char* s2, int L2, int* n_identical, int* aln_len, char* aln) {
    const int W = L2 + 1;
    const size_t cells = (size_t)(L1 + 1) * W;
    float* V[3]; int* TB[3];
    for (int k = 0; k < 3; k++) {
        V[k] = (float*)malloc(cells * sizeof(float)); TB[k] = (int*)malloc(cells * sizeof(int));
        for (size_t c = 0; c < cells; c++) { V[k][c] = LOG_ZERO; TB[k][c] = -1; }
    }
    const float init0 = logf((float)0.6080327034), init1 = logf((float)0.1959836632);
    V[0][0] = init0; V[1][0] = init1; V[2][0] = init1;
    for (int i = 0; i <= L1; i++)
        for (int j = 0; j <= L2; j++) {
            if (i > 0 && j > 0)
                for (int k = 0; k < 3; k++) {
                    float nv = V[k][IX(i - 1, j - 1)] + t->ltrans[k][0] + t->match[R1(i)][R2(j)];
                    if (V[0][IX(i, j)] < nv) { V[0][IX(i, j)] = nv; TB[0][IX(i, j)] = k; }
                }
            if (i > 0) {
                float fm = t->ins[R1(i)] + V[0][IX(i - 1, j)] + t->ltrans[0][1];
                float fi = t->ins[R1(i)] + V[1][IX(i - 1, j)] + t->ltrans[1][1];
                if (fm >= fi) { V[1][IX(i, j)] = fm; TB[1][IX(i, j)] = 0; } else { V[1][IX(i, j)] = fi; TB[1][IX(i, j)] = 1; }
            }
            if (j > 0) {
                float fm = t->ins[R2(j)] + V[0][IX(i, j - 1)] + t->ltrans[0][2];
                float fi = t->ins[R2(j)] + V[2][IX(i, j - 1)] + t->ltrans[2][2];
                if (fm >= fi) { V[2][IX(i, j)] = fm; TB[2][IX(i, j)] = 0; } else { V[2][IX(i, j)] = fi; TB[2][IX(i, j)] = 2; }
            }
        }
    float best = LOG_ZERO; int state = -1;
    const float iv[3] = {init0, init1, init1};
    for (int k = 0; k < 3; k++) {
        float p = V[k][IX(L1, L2)] + iv[k];
        if (best < p) { best = p; state = k; }
    }
    int r = L1, c = L2, len = 0, ident = 0;
    while ((r != 0 || c != 0) && state >= 0) {
        int ns = TB[state][IX(r, c)];
        if (state == 0) { if (s1[r - 1] == s2[c - 1]) ident++; if (aln) aln[len] = 'B'; r--; c--; }
        else if (state == 1) { if (aln) aln[len] = 'X'; r--; }
        else { if (aln) aln[len] = 'Y'; c--; }
        len++;
        state = ns;
    }
    if (aln) { for (int a = 0, b = len - 1; a < b; a++, b--) { char x = aln[a]; aln[a] = aln[b]; aln[b] = x; } }
    if (n_identical) *n_identical = ident;
    if (aln_len) *aln_len = len;
    for (int k = 0; k < 3; k++) { free(V[k]); free(TB[k]); }
    return best;
}

/* MSA.cpp:838-881 */
int orc_model_adjustment(int npairs, const int32_t* n_identical, const int32_t* aln_len, float* identity_out, float* sigma_out, float* init_distrib2) {
    float identity = 0;
    float* pid = (float*)malloc(sizeof(float) * (npairs + 1));
    for (int k = 0; k < npairs; k++) {
        float nc = (float)n_identical[k];
        pid[k] = nc / aln_len[k];
        identity += nc / aln_len[k];
    }
    identity /= npairs;
    float variance = 0;
    for (int k = 0; k < npairs; k++) variance += (pid[k] - identity) * (pid[k] - identity);
    variance /= npairs;
    variance = sqrtf(variance);
    free(pid);
    if (identity_out) *identity_out = identity;
    if (sigma_out) *sigma_out = variance;
    if (init_distrib2) *init_distrib2 = (identity <= 0.5) ? orc_init_distrib2_for_identity(identity) : 0.700645f;
    int vm = (variance > 0.115) ? 10 : 0;
    if (identity <= 0.18) return vm + 0;
    else if (identity <= 0.25) return vm + 1;
    else if (identity <= 0.4) return vm + 2;
    else if (identity <= 0.7) return vm + 3;
    return vm + 4;
}

/* MSA.cpp:646-762 */
int orc_g_features(const orc_hmm_tables* t, int n, const int32_t* len, const char* residues, const int64_t* res_off,
                   float theta, char* line, int cap) {
    const char* al = MLP_HMM_ALPHABET;
    int idx[26];
    for (int k = 0; k < 26; k++) idx[k] = -1;
    for (int k = 0; k < 20; k++) idx[al[k] - 'A'] = k;
    for (int i = 0; i < n; i++) for (int k = 0; k < len[i]; k++) if (idx[residues[res_off[i] + k] - 'A'] < 0) return 1;
    const int npairs = n * (n - 1) / 2;
    float identity = 0, tmp_sp = 0;
    int avg_length = 0, max_len = 0, tmp_sp_idx = 0;
    float* arr = (float*)calloc(20000 + 16, sizeof(float));
    float* pids = (float*)malloc(sizeof(float) * (npairs + 1));
    int p = 0;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++, p++) {
            const char* s1 = residues + res_off[a]; const char* s2 = residues + res_off[b];
            char* aln = (char*)malloc(len[a] + len[b] + 2);
            int same = 0, alen = 0;
            orc_viterbi(t, s1, len[a], s2, len[b], &same, &alen, aln);
            avg_length += alen;
            if (alen > max_len) max_len = alen;
            int i = 1, j = 1, num_idx = 0;
            float nc = 0;
            for (int k = 0; k < alen; k++) {
                if (aln[k] == 'B') {
                    char c1 = s1[i - 1], c2 = s2[j - 1]; i++; j++;
                    if (c1 == c2) nc += 1;
                    float bl = MLP_BLOSUM62[idx[c1 - 'A'] * 20 + idx[c2 - 'A']];
                    if (bl < 10) { arr[num_idx] += bl; tmp_sp += bl; } else { arr[num_idx] += 0; }
                } else if (aln[k] == 'X') i++;
                else j++;
                tmp_sp_idx += 1; num_idx++;
            }
            pids[p] = nc / alen;
            identity += nc / alen;
            free(aln);
        }
    tmp_sp /= tmp_sp_idx;
    identity /= npairs;
    avg_length /= npairs;
    float peak = 0;
    for (int k = 0; k < max_len; k++) { arr[k] /= npairs; if (theta <= arr[k]) peak += 1; }
    peak /= max_len;
    float variance = 0;
    for (int k = 0; k < npairs; k++) variance += (pids[k] - identity) * (pids[k] - identity);
    variance /= npairs;
    variance = sqrtf(variance);
    float factor = 2 * (float)n - (float)avg_length;
    snprintf(line, cap, "%f\t%f\t%d\t%d\t%f\t%f\t%f", identity, variance, n, avg_length, tmp_sp, peak, factor);
    free(arr); free(pids);
    return 0;
}

/* ---------------------------------------------------------------- partition function */
/* QP: PartitionFunction.cpp:71-157 (forward), :180-291 (reverse). FP64, no overflow check. */
int orc_part_posterior_qp(const orc_part_tables* t, const char* s1, int L1, const char* s2, int L2, float* post) {
    const int W = L2 + 1;
    const size_t cells = (size_t)(L1 + 1) * W;
    double* Zfm = (double*)calloc(cells, sizeof(double));
    double* buf = (double*)calloc((size_t)6 * W, sizeof(double));
    double *Ze = buf, *Zf = buf + 2 * W;
    double zz = 0;
    memset(post, 0, cells * sizeof(float));
    Zfm[0] = 1.0;
    Zf[1 * W + 0] = Zfm[0] * t->tgo;
    Ze[0 * W + 1] = Zfm[0] * t->tgo;
    for (int j = 2; j <= L2; j++) Ze[j] = Ze[j - 1] * t->tge;
    for (int i = 1; i <= L1; i++) {
        for (int j = 1; j <= L2; j++) {
            double score = t->sub[R1(i)][R2(j)];
            double open0 = t->go, open1 = t->go, ext0 = t->ge, ext1 = t->ge;
            if (i == L1) { open0 = t->tgo; ext0 = t->tge; }
            if (j == L2) { open1 = t->tgo; ext1 = t->tge; }
            Ze[W + j] = Zfm[IX(i, j - 1)] * open0 + Ze[W + j - 1] * ext0;
            Zf[W + j] = Zfm[IX(i - 1, j)] * open1 + Zf[j] * ext1;
            Zfm[IX(i, j)] = (Zfm[IX(i - 1, j - 1)] + Ze[j - 1] + Zf[j - 1]) * score;
            zz = Zfm[IX(i, j)] + Ze[W + j] + Zf[W + j];
        }
        for (int x = 0; x <= L2; x++) { Ze[x] = Ze[W + x]; Ze[W + x] = 0; Zf[x] = Zf[W + x]; Zf[W + x] = 0; }
        Zf[W + 0] = 1;
    }
    Zfm[0] = zz;
    /* reverse */
    memset(buf, 0, (size_t)6 * W * sizeof(double));
    double *Zm = buf; Ze = buf + 2 * W; Zf = buf + 4 * W;
    Zm[W + L2] = 1;
    Zf[W + L2] = Zm[W + L2] * t->tgo;
    if (L2 >= 1) Ze[L2 - 1] = Zm[W + L2] * t->tgo;
    for (int j = L2 - 2; j >= 0; j--) Ze[j] = Ze[j + 1] * t->tge;
    for (int i = L1 - 1; i >= 0; i--) {
        for (int j = L2 - 1; j >= 0; j--) {
            double scorez = t->sub[s1[i] - 'A'][s2[j] - 'A'];
            double open0 = t->go, open1 = t->go, ext0 = t->ge, ext1 = t->ge;
            if (i == 0) { open0 = t->tgo; ext0 = t->tge; }
            if (j == 0) { open1 = t->tgo; ext1 = t->tge; }
            Zf[W + j] = Zm[W + j] * open1 + Zf[j] * ext1;
            Ze[W + j] = Zm[j + 1] * open0 + Ze[W + j + 1] * ext0;
            Zm[j] = (Zm[W + j + 1] + Zf[j + 1] + Ze[j + 1]) * scorez;
            double tmp = Zfm[IX(i + 1, j + 1)] * Zm[j];
            tmp /= (scorez * Zfm[0]);
            float p = (float)tmp;
            if (p <= 1 && p >= 0.001) post[IX(i + 1, j + 1)] = p;
        }
        for (int x = 0; x <= L2; x++) {
            Ze[x] = Ze[W + x]; Ze[W + x] = 0;
            Zf[x] = Zf[W + x]; Zf[W + x] = 0;
            Zm[W + x] = Zm[x]; Zm[x] = 0;
        }
        Zf[L2] = 1;
    }
    post[0] = 0;
    free(Zfm); free(buf);
    return 0;
}

/* cpnp: MSAPartProbs.cpp:400-660 (partf), :78-394 (revers_partf), 80-bit long double.
 * The reference's outer loop runs over seq2 (sequences[1]), inner over seq1 (sequences[0]); the result is
 * written transposed into the (L1+1)x(L2+1) matrix (:297).  Returns 1 on the HUGE_VALL overflow exit. */
int orc_part_posterior_cpnp(const orc_part_tables* t, const char* s1, int L1, const char* s2, int L2, float* post) {
    const int len0 = L1, len1 = L2;              /* sequences[0]=seq1 (inner), sequences[1]=seq2 (outer) */
    const int Wd = len0 + 1;
    const size_t cells = (size_t)(len1 + 1) * Wd;
    long double* Zfm = (long double*)calloc(cells, sizeof(long double));
    long double* buf = (long double*)calloc((size_t)6 * Wd, sizeof(long double));
    long double *Ze = buf, *Zf = buf + 2 * Wd;
    long double zz = 0;
    int rc = 0;
    for (size_t c = 0; c < (size_t)(L1 + 1) * (L2 + 1); c++) post[c] = 0;
#define ZX(i, j) ((size_t)(i) * Wd + (j))
    Zfm[0] = 1.0;
    Zf[Wd + 0] = Zfm[0] * t->tgo;
    Ze[1] = Zfm[0] * t->tgo;
    for (int j = 2; j <= len0; j++) Ze[j] = Ze[j - 1] * t->tge;
    for (int i = 1; i <= len1 && !rc; i++) {
        for (int j = 1; j <= len0; j++) {
            double score = t->sub[s2[i - 1] - 'A'][s1[j - 1] - 'A'];
            double open0 = t->go, open1 = t->go, ext0 = t->ge, ext1 = t->ge;
            if (i == len1) { open0 = t->tgo; ext0 = t->tge; }
            if (j == len0) { open1 = t->tgo; ext1 = t->tge; }
            Ze[Wd + j] = Zfm[ZX(i, j - 1)] * open0 + Ze[Wd + j - 1] * ext0;
            Zf[Wd + j] = Zfm[ZX(i - 1, j)] * open1 + Zf[j] * ext1;
            Zfm[ZX(i, j)] = (Zfm[ZX(i - 1, j - 1)] + Ze[j - 1] + Zf[j - 1]) * score;
            if (Ze[Wd + j] >= HUGE_VALL || Zf[Wd + j] >= HUGE_VALL || Zfm[ZX(i, j)] >= HUGE_VALL) { rc = 1; break; }
            zz = Zfm[ZX(i, j)] + Ze[Wd + j] + Zf[Wd + j];
        }
        for (int x = 0; x <= len0; x++) { Ze[x] = Ze[Wd + x]; Ze[Wd + x] = 0; Zf[x] = Zf[Wd + x]; Zf[Wd + x] = 0; }
        Zf[Wd + 0] = 1;
    }
    if (!rc) {
        Zfm[0] = zz;
        memset(buf, 0, (size_t)6 * Wd * sizeof(long double));
        long double *Zm = buf; Ze = buf + 2 * Wd; Zf = buf + 4 * Wd;
        Zm[Wd + len0] = 1;
        Zf[Wd + len0] = Zm[Wd + len0] * t->tgo;
        if (len0 >= 1) Ze[len0 - 1] = Zm[Wd + len0] * t->tgo;
        for (int j = len0 - 2; j >= 0; j--) Ze[j] = Ze[j + 1] * t->tge;
        for (int i = len1 - 1; i >= 0; i--) {
            for (int j = len0 - 1; j >= 0; j--) {
                double scorez = t->sub[s2[i] - 'A'][s1[j] - 'A'];
                double open0 = t->go, open1 = t->go, ext0 = t->ge, ext1 = t->ge;
                if (i == 0) { open0 = t->tgo; ext0 = t->tge; }
                if (j == 0) { open1 = t->tgo; ext1 = t->tge; }
                Zf[Wd + j] = Zm[Wd + j] * open1 + Zf[j] * ext1;
                Ze[Wd + j] = Zm[j + 1] * open0 + Ze[Wd + j + 1] * ext0;
                Zm[j] = (Zm[Wd + j + 1] + Zf[j + 1] + Ze[j + 1]) * scorez;
                long double tmp = Zfm[ZX(i + 1, j + 1)] * Zm[j];
                tmp /= (scorez * Zfm[0]);
                post[(size_t)(j + 1) * (len1 + 1) + (i + 1)] = (float)tmp;
            }
            for (int x = 0; x <= len0; x++) {
                Ze[x] = Ze[Wd + x]; Ze[Wd + x] = 0;
                Zf[x] = Zf[Wd + x]; Zf[Wd + x] = 0;
                Zm[Wd + x] = Zm[x]; Zm[x] = 0;
            }
            Zf[len0] = 1;
        }
    }
#undef ZX
    free(Zfm); free(buf);
    return rc;
}

/* ---------------------------------------------------------------- merge + MEA */
/* PosteriorStage.cpp:156-196 */
float orc_combine_qp(int L1, int L2, const float* in1, const float* in2, float* out) {
    const int W = L2 + 1;
    float* two = (float*)malloc(sizeof(float) * 2 * W);
    float *oldRow = two, *newRow = two + W;
    for (int j = 0; j < 2 * W; j++) two[j] = 0;   /* reference leaves oldRow uninitialised; row 0 overwrites newRow before any read of oldRow matters */
    for (int i = 0; i <= L1; i++) {
        for (int j = 0; j <= L2; j++) {
            if (i == 0 || j == 0) { out[IX(i, j)] = 0; newRow[j] = 0; }
            else {
                float v1 = in1[IX(i, j)], v2 = in2[IX(i, j)];
                float o = sqrtf((v1 * v1 + v2 * v2) * 0.5f);
                out[IX(i, j)] = o;
                float a = o + oldRow[j - 1], b = newRow[j - 1], c = oldRow[j];
                float m = a > b ? a : b;
                newRow[j] = m > c ? m : c;
            }
        }
        float* tmp = oldRow; oldRow = newRow; newRow = tmp;
    }
    float total = oldRow[L2];
    free(two);
    return 1.0f - total / (L1 < L2 ? L1 : L2);
}

/* MSA.cpp:1001 (-p 0: (dbl^2+glob^2)+loc^2) ; MSA.cpp:1708 (-p 1: (glob^2+loc^2)+dbl^2) */
void orc_merge3_cpnp(int n, const float* p5, const float* pp, const float* pl, int p1_order, float* out) {
    for (int k = 0; k < n; k++) {
        float v1 = p5[k], v2 = pp[k], v3 = pl[k];
        out[k] = p1_order ? sqrtf((v2 * v2 + v3 * v3 + v1 * v1) / 3) : sqrtf((v1 * v1 + v2 * v2 + v3 * v3) / 3);
    }
}

/* ProbabilisticModel.h:804-864; tie order D >= L >= U (ScoreType.h:347-366). n_match = number of 'B' in the traceback. */
float orc_mea_score(int L1, int L2, const float* post, int* n_match) {
    const int W = L2 + 1;
    float* two = (float*)calloc(2 * W, sizeof(float));
    float *oldRow = two, *newRow = two + W;
    char* tb = (char*)malloc((size_t)(L1 + 1) * W);
    for (int j = 0; j <= L2; j++) tb[j] = 'L';
    for (int i = 1; i <= L1; i++) {
        newRow[0] = 0; tb[IX(i, 0)] = 'U';
        for (int j = 1; j <= L2; j++) {
            float x1 = post[IX(i, j)] + oldRow[j - 1], x2 = newRow[j - 1], x3 = oldRow[j];
            float x; char b;
            if (x1 >= x2) { if (x1 >= x3) { x = x1; b = 'D'; } else { x = x3; b = 'U'; } }
            else if (x2 >= x3) { x = x2; b = 'L'; }
            else { x = x3; b = 'U'; }
            newRow[j] = x; tb[IX(i, j)] = b;
        }
        float* tmp = oldRow; oldRow = newRow; newRow = tmp;
    }
    float total = oldRow[L2];
    if (n_match) {
        int r = L1, c = L2, nb = 0;
        while (r != 0 || c != 0) {
            char ch = tb[IX(r, c)];
            if (ch == 'L') c--; else if (ch == 'U') r--; else { r--; c--; nb++; }
        }
        *n_match = nb;
    }
    free(two); free(tb);
    return total;
}

/* pdoAlign per-pair body MSA.cpp:935-1023 ; ArrangePosteriorProbs MSA.cpp:1660-1756 ; PosteriorStage::computePairwise PosteriorStage.cpp:123-154 */
int orc_pair_posterior(int flavour, int model_mask, const orc_hmm_tables* ht, const orc_part_tables* pt,
                       const char* s1, int L1, const char* s2, int L2, float* post, float* dist) {
    const size_t cells = (size_t)(L1 + 1) * (L2 + 1);
    int rc = 0;
    if (flavour == ORC_QP) {
        float* ph = (float*)malloc(cells * sizeof(float));
        float* pp = (float*)malloc(cells * sizeof(float));
        orc_part_posterior_qp(pt, s1, L1, s2, L2, pp);
        orc_hmm5_posterior(ht, s1, L1, s2, L2, 1, ph, NULL);
        *dist = orc_combine_qp(L1, L2, ph, pp, post);
        free(ph); free(pp);
        return 0;
    }
    float *p5 = NULL, *pp = NULL, *pl = NULL;
    if (model_mask & ORC_M_HMM5) { p5 = (float*)malloc(cells * sizeof(float)); orc_hmm5_posterior(ht, s1, L1, s2, L2, 0, p5, NULL); }
    if (model_mask & ORC_M_PART) { pp = (float*)malloc(cells * sizeof(float)); rc = orc_part_posterior_cpnp(pt, s1, L1, s2, L2, pp); }
    if (model_mask & ORC_M_LOCAL) { pl = (float*)malloc(cells * sizeof(float)); orc_local_posterior(ht, s1, L1, s2, L2, pl, NULL); }
    if (p5 && pp && pl) orc_merge3_cpnp((int)cells, p5, pp, pl, flavour == ORC_CPNP_P1, post);
    else memcpy(post, p5 ? p5 : (pp ? pp : pl), cells * sizeof(float));
    int nm = 0;
    float score = orc_mea_score(L1, L2, post, flavour == ORC_CPNP_P1 ? &nm : NULL);
    if (flavour == ORC_CPNP_P1) *dist = score / nm;
    else *dist = 1.0f - score / (L1 < L2 ? L1 : L2);
    free(p5); free(pp); free(pl);
    return rc;
}

/* ---------------------------------------------------------------- sparse */
/* SparseEntry.h:31-32 */
static inline uint16_t q_store(float v) { return (uint16_t)(v * 65535); }
static inline float q_load(uint16_t u) { return (float)u / 65535; }

int64_t orc_sparsify(int L1, int L2, const float* post, float cutoff, int quantize_u16,
                     int32_t* rowptr, int32_t* col, float* val, int64_t cap) {
    const int W = L2 + 1;
    int64_t n = 0;
    rowptr[0] = 0; rowptr[1] = 0;
    for (int i = 1; i <= L1; i++) {
        for (int j = 1; j <= L2; j++) {
            float v = post[IX(i, j)];
            if (v >= cutoff) {
                if (n < cap) { col[n] = j; val[n] = quantize_u16 ? q_load(q_store(v)) : v; }
                n++;
            }
        }
        rowptr[i + 1] = (int32_t)n;
    }
    return n;
}

void orc_transpose(int L1, int L2, const int32_t* rowptr, const int32_t* col, const float* val,
                   int32_t* t_rowptr, int32_t* t_col, float* t_val) {
    int32_t* cnt = (int32_t*)calloc(L2 + 2, sizeof(int32_t));
    for (int i = 1; i <= L1; i++) for (int k = rowptr[i]; k < rowptr[i + 1]; k++) cnt[col[k]]++;
    t_rowptr[0] = 0; t_rowptr[1] = 0;
    for (int j = 1; j <= L2; j++) t_rowptr[j + 1] = t_rowptr[j] + cnt[j];
    memset(cnt, 0, (L2 + 2) * sizeof(int32_t));
    for (int i = 1; i <= L1; i++)
        for (int k = rowptr[i]; k < rowptr[i + 1]; k++) {
            int j = col[k];
            int d = t_rowptr[j] + cnt[j]++;
            t_col[d] = i; t_val[d] = val[k];
        }
    free(cnt);
}

static int set_append(orc_csr_set* s, int a, int b, int La, const int32_t* rowptr, const int32_t* col, const float* val) {
    int64_t nz = rowptr[La + 1];
    int64_t rp, nzo;
#pragma omp critical(orc_set_append)
    {
        rp = s->rp_used; nzo = s->nz_used;
        s->rp_used += La + 2; s->nz_used += nz;
    }
    if (rp + La + 2 > s->rp_cap || nzo + nz > s->nz_cap) return -1;
    s->rp_off[(size_t)a * s->n + b] = rp;
    s->nz_off[(size_t)a * s->n + b] = nzo;
    memcpy(s->rowptr + rp, rowptr, (La + 2) * sizeof(int32_t));
    memcpy(s->col + nzo, col, nz * sizeof(int32_t));
    memcpy(s->val + nzo, val, nz * sizeof(float));
    return 0;
}

static int store_pair(orc_csr_set* out, int a, int b, int La, int Lb, const float* post, float cutoff, int quant) {
    int64_t cap = (int64_t)La * Lb;
    int32_t* rp = (int32_t*)malloc((La + 2) * sizeof(int32_t));
    int32_t* trp = (int32_t*)malloc((Lb + 2) * sizeof(int32_t));
    int32_t* c = (int32_t*)malloc(cap * sizeof(int32_t));
    float* v = (float*)malloc(cap * sizeof(float));
    int64_t nz = orc_sparsify(La, Lb, post, cutoff, quant, rp, c, v, cap);
    int32_t* tc = (int32_t*)malloc((nz + 1) * sizeof(int32_t));
    float* tv = (float*)malloc((nz + 1) * sizeof(float));
    orc_transpose(La, Lb, rp, c, v, trp, tc, tv);
    int rc = set_append(out, a, b, La, rp, c, v);
    rc |= set_append(out, b, a, Lb, trp, tc, tv);
    free(rp); free(trp); free(c); free(v); free(tc); free(tv);
    return rc;
}

int orc_posterior_stage(int flavour, int model_mask, const orc_hmm_tables* ht, const orc_part_tables* pt,
                        int n, const int32_t* len, const char* residues, const int64_t* res_off,
                        float cutoff, float* dist, orc_csr_set* out, int threads) {
    int npairs = n * (n - 1) / 2, rc = 0;
    int* pa = (int*)malloc(sizeof(int) * (npairs + 1));
    int* pb = (int*)malloc(sizeof(int) * (npairs + 1));
    { int p = 0; for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) { pa[p] = a; pb[p] = b; p++; } }
    out->n = n; out->len = len; out->rp_used = 0; out->nz_used = 0;
    for (int i = 0; i < n; i++) dist[(size_t)i * n + i] = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic) reduction(| : rc)
    for (int p = 0; p < npairs; p++) {
        int a = pa[p], b = pb[p];
        int La = len[a], Lb = len[b];
        float* post = (float*)malloc((size_t)(La + 1) * (Lb + 1) * sizeof(float));
        float d = 0;
        rc |= orc_pair_posterior(flavour, model_mask, ht, pt, residues + res_off[a], La, residues + res_off[b], Lb, post, &d);
        dist[(size_t)a * n + b] = dist[(size_t)b * n + a] = d;
        rc |= store_pair(out, a, b, La, Lb, post, cutoff, flavour == ORC_QP) ? 2 : 0;
        free(post);
    }
    free(pa); free(pb);
    return rc;
}

/* ---------------------------------------------------------------- consistency */
#define SET_RP(s, a, b) ((s)->rowptr + (s)->rp_off[(size_t)(a) * (s)->n + (b)])
#define SET_COL(s, a, b) ((s)->col + (s)->nz_off[(size_t)(a) * (s)->n + (b)])
#define SET_VAL(s, a, b) ((s)->val + (s)->nz_off[(size_t)(a) * (s)->n + (b)])

static void densify(const orc_csr_set* s, int a, int b, float* P) {
    int La = s->len[a], W = s->len[b] + 1;
    memset(P, 0, (size_t)(La + 1) * W * sizeof(float));
    const int32_t* rp = SET_RP(s, a, b); const int32_t* c = SET_COL(s, a, b); const float* v = SET_VAL(s, a, b);
    for (int i = 1; i <= La; i++) for (int k = rp[i]; k < rp[i + 1]; k++) P[IX(i, c[k])] = v[k];
}
static void mask_to(const orc_csr_set* s, int a, int b, float* P) {
    int La = s->len[a], Lb = s->len[b], W = Lb + 1;
    const int32_t* rp = SET_RP(s, a, b); const int32_t* c = SET_COL(s, a, b);
    for (int y = 0; y <= Lb; y++) P[y] = 0;
    for (int x = 1; x <= La; x++) {
        int curr = 0;
        for (int k = rp[x]; k < rp[x + 1]; k++) { while (curr < c[k]) P[IX(x, curr++)] = 0; curr++; }
        while (curr <= Lb) P[IX(x, curr++)] = 0;
    }
}
/* MSA.cpp:1290-1322 / ConsistencyStage.cpp:269-300; weight<0 means unweighted (cpnp) */
static void relax_xz_zy(int Lx, int W, const int32_t* xz_rp, const int32_t* xz_c, const float* xz_v,
                        const int32_t* zy_rp, const int32_t* zy_c, const float* zy_v, float weight, int weighted, float* P) {
    for (int i = 1; i <= Lx; i++)
        for (int k = xz_rp[i]; k < xz_rp[i + 1]; k++) {
            int z = xz_c[k]; float xv = xz_v[k];
            for (int m = zy_rp[z]; m < zy_rp[z + 1]; m++) {
                if (weighted) P[IX(i, zy_c[m])] += weight * xv * zy_v[m];
                else P[IX(i, zy_c[m])] += xv * zy_v[m];
            }
        }
}
/* MSA.cpp:1331-1360 */
static void relax1_zx_zy(int Lz, int W, const int32_t* zx_rp, const int32_t* zx_c, const float* zx_v,
                         const int32_t* zy_rp, const int32_t* zy_c, const float* zy_v, float* P) {
    for (int k = 1; k <= Lz; k++)
        for (int a = zx_rp[k]; a < zx_rp[k + 1]; a++) {
            float xv = zx_v[a]; int x = zx_c[a];
            for (int m = zy_rp[k]; m < zy_rp[k + 1]; m++) P[IX(x, zy_c[m])] += xv * zy_v[m];
        }
}

/* MSA.cpp:1172-1281: only the a<b orientation of the INPUT set is read (as the reference stores it);
 * k>j uses an explicit transpose of S_jk (MSA.cpp:1226). */
int orc_relax_cpnp(const orc_csr_set* in, float cutoff, orc_csr_set* out, int threads) {
    const int n = in->n; int rc = 0;
    out->n = n; out->len = in->len; out->rp_used = 0; out->nz_used = 0;
    int npairs = n * (n - 1) / 2;
    int* pa = (int*)malloc(sizeof(int) * (npairs + 1)); int* pb = (int*)malloc(sizeof(int) * (npairs + 1));
    { int p = 0; for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) { pa[p] = a; pb[p] = b; p++; } }
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic) reduction(| : rc)
    for (int p = 0; p < npairs; p++) {
        int i = pa[p], j = pb[p];
        int Li = in->len[i], Lj = in->len[j], W = Lj + 1;
        size_t cells = (size_t)(Li + 1) * W;
        float* P = (float*)malloc(cells * sizeof(float));
        densify(in, i, j, P);
        for (size_t c = 0; c < cells; c++) P[c] += P[c];
        for (int k = 0; k < n; k++) {
            if (k == i || k == j) continue;
            if (k < i)
                relax1_zx_zy(in->len[k], W, SET_RP(in, k, i), SET_COL(in, k, i), SET_VAL(in, k, i), SET_RP(in, k, j), SET_COL(in, k, j), SET_VAL(in, k, j), P);
            else if (k < j)
                relax_xz_zy(Li, W, SET_RP(in, i, k), SET_COL(in, i, k), SET_VAL(in, i, k), SET_RP(in, k, j), SET_COL(in, k, j), SET_VAL(in, k, j), 0, 0, P);
            else {
                int Lk = in->len[k];
                int64_t nz = SET_RP(in, j, k)[Lj + 1];
                int32_t* trp = (int32_t*)malloc((Lk + 2) * sizeof(int32_t));
                int32_t* tc = (int32_t*)malloc((nz + 1) * sizeof(int32_t));
                float* tv = (float*)malloc((nz + 1) * sizeof(float));
                orc_transpose(Lj, Lk, SET_RP(in, j, k), SET_COL(in, j, k), SET_VAL(in, j, k), trp, tc, tv);
                relax_xz_zy(Li, W, SET_RP(in, i, k), SET_COL(in, i, k), SET_VAL(in, i, k), trp, tc, tv, 0, 0, P);
                free(trp); free(tc); free(tv);
            }
        }
        for (size_t c = 0; c < cells; c++) P[c] /= n;
        mask_to(in, i, j, P);
        rc |= store_pair(out, i, j, Li, Lj, P, cutoff, 0) ? 2 : 0;
        free(P);
    }
    free(pa); free(pb);
    return rc;
}

/* ConsistencyStage.cpp:133-266 with the default Deterministic filter / Max function
 * (Configuration.cpp:100-112): z is accepted iff max(d_iz, d_jz) <= selectivity. */
int orc_relax_qp(const orc_csr_set* in, const float* weights, const float* seldist, float selectivity,
                 float selfweight, float cutoff, orc_csr_set* out, int threads) {
    const int n = in->n; int rc = 0;
    out->n = n; out->len = in->len; out->rp_used = 0; out->nz_used = 0;
    int npairs = n * (n - 1) / 2;
    int* pa = (int*)malloc(sizeof(int) * (npairs + 1)); int* pb = (int*)malloc(sizeof(int) * (npairs + 1));
    { int p = 0; for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) { pa[p] = a; pb[p] = b; p++; } }
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic) reduction(| : rc)
    for (int p = 0; p < npairs; p++) {
        int i = pa[p], j = pb[p];
        int Li = in->len[i], Lj = in->len[j], W = Lj + 1;
        size_t cells = (size_t)(Li + 1) * W;
        float* P = (float*)malloc(cells * sizeof(float));
        densify(in, i, j, P);
        int accepted = 0;
        for (int k = 0; k < n; k++) {
            if (k == i || k == j) continue;
            float x = fmaxf(seldist[(size_t)i * n + k], seldist[(size_t)j * n + k]);
            if (x <= selectivity) accepted++;
        }
        float wi_wj = 1.0f + (selfweight - 1.0f) * (float)accepted / selectivity;
        wi_wj *= weights[i] + weights[j];
        float sumW = 1.0f;
        for (int k = 0; k < n; k++) {
            if (k == i || k == j) continue;
            float x = fmaxf(seldist[(size_t)i * n + k], seldist[(size_t)j * n + k]);
            if (!(x <= selectivity)) continue;
            float wk = weights[k] / wi_wj;
            sumW += wk;
            relax_xz_zy(Li, W, SET_RP(in, i, k), SET_COL(in, i, k), SET_VAL(in, i, k), SET_RP(in, k, j), SET_COL(in, k, j), SET_VAL(in, k, j), wk, 1, P);
        }
        for (size_t c = 0; c < cells; c++) P[c] /= sumW;
        mask_to(in, i, j, P);
        rc |= store_pair(out, i, j, Li, Lj, P, cutoff, 1) ? 2 : 0;
        free(P);
    }
    free(pa); free(pb);
    return rc;
}
