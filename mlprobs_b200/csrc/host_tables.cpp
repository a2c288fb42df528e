// Host-side construction of the reference's default model tables (product code; the oracle has its own copy).
// Built with the same glibc calls the reference uses so the float tables are bit-identical
// (logf for the HMM tables, expf for cpnp's substitution table, double exp for gap terms).
#include "../../include/mlprobs_b200.h"
#include "param_data.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

// cpnp ProbabilisticModel.h:58-135 ; QP PairHmm.cpp:4-33 + ProbabilisticModel.cpp:15-56
void build_hmm(float init_distrib2, mlp_hmm_tables* t) {
    const float initD[5] = {0.6814756989f, 8.615339902e-05f, 8.615339902e-05f, 0.1591759622f, 0.1591759622f};
    const float gapOpen[2] = {0.0119511066f, 0.008008334786f};
    const float gapExt[2] = {0.3965826333f, 0.8988758326f};
    const float lgo = 0.01993141696f, lge = 0.7943345308f;
    float tm[5][5];
    std::memset(tm, 0, sizeof tm);
    tm[0][0] = 1;
    for (int i = 0; i < 2; i++) {
        const int x = 2 * i + 1, y = 2 * i + 2;
        tm[0][x] = gapOpen[i];
        tm[0][y] = gapOpen[i];
        tm[0][0] -= (gapOpen[i] + gapOpen[i]);
        tm[x][x] = gapExt[i];
        tm[y][y] = gapExt[i];
        tm[x][0] = 1 - gapExt[i];
        tm[y][0] = 1 - gapExt[i];
    }
    for (int i = 0; i < 5; i++) {
        t->init[i] = logf(initD[i]);
        for (int j = 0; j < 5; j++) t->trans[i][j] = logf(tm[i][j]);
    }
    const float unknown_single = 1e-5, unknown_pair = 1e-10;   // MSA.cpp:46-47 / ProbabilisticModel.cpp:36-39
    for (int a = 0; a < 26; a++) {
        t->ins[a] = logf(unknown_single);
        for (int b = 0; b < 26; b++) t->match[a][b] = logf(unknown_pair);
    }
    const char* al = MLP_HMM_ALPHABET;
    for (int i = 0; i < 20; i++) {
        t->ins[al[i] - 'A'] = logf(MLP_EMIT_SINGLE[i]);
        for (int j = 0; j <= i; j++) {
            const float v = logf(MLP_EMIT_PAIRS_TRI[i * (i + 1) / 2 + j]);
            t->match[al[i] - 'A'][al[j] - 'A'] = v;
            t->match[al[j] - 'A'][al[i] - 'A'] = v;
        }
    }
    float lt[3][3];
    std::memset(lt, 0, sizeof lt);
    lt[0][0] = 1;
    lt[0][1] = lgo;
    lt[0][2] = lgo;
    lt[0][0] -= (lgo + lgo);
    lt[1][1] = lge;
    lt[2][2] = lge;
    lt[1][0] = 1 - lge;
    lt[2][0] = 1 - lge;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t->ltrans[i][j] = logf(lt[i][j]);
    t->rtrans[0] = logf(init_distrib2);
    t->rtrans[1] = logf(1 - init_distrib2);
}

// MSAReadMatrix.cpp:85-116,158-210 ; MSAPartProbs.cpp:698-709
void build_part_cpnp(mlp_part_tables* t) {
    const float temperature = 5;
    const float beta = (float)(1.0 / temperature);
    for (int a = 0; a < 26; a++)
        for (int b = 0; b < 26; b++) t->sub[a][b] = NAN;
    const char* al = MLP_GONNET160_ALPHABET;
    const int n = (int)std::strlen(al);
    int pos = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            const double v = expf(beta * MLP_GONNET160_TRI[pos++]);
            t->sub[al[i] - 'A'][al[j] - 'A'] = v;
            t->sub[al[j] - 'A'][al[i] - 'A'] = v;
        }
    const float gapopen = -22, gapext = -1;
    const double b = beta;
    t->tgo = std::exp(b * 0.0);
    t->tge = std::exp(b * 0.0);
    t->go = std::exp(b * (double)gapopen);
    t->ge = std::exp(b * (double)gapext);
}

// ExpPartitionFunctionParams.h:30-49 ; Configuration.cpp:330-332
void build_part_qp(mlp_part_tables* t) {
    const double temperature = 5.6007, gi = -25.3549, ge = -1.30113;
    const double beta = 1.0 / temperature;
    std::memset(t->sub, 0, sizeof t->sub);
    const char* al = MLP_VTML200_ALPHABET;
    const int n = (int)std::strlen(al);
    for (int i = 0; i < n - 1; i++)
        for (int j = 0; j <= i; j++) {
            const double v = std::exp(beta * MLP_VTML200[i * n + j]);
            t->sub[al[i] - 'A'][al[j] - 'A'] = v;
            t->sub[al[j] - 'A'][al[i] - 'A'] = v;
        }
    t->go = std::exp(beta * gi);
    t->ge = std::exp(beta * ge);
    t->tgo = std::exp(beta * 0);
    t->tge = std::exp(beta * 0);
}

}  // namespace

extern "C" int mlp_default_tables(int flavour, float init_distrib2, mlp_hmm_tables* hmm, mlp_part_tables* part) {
    if (flavour < MLP_QP || flavour > MLP_CPNP_P1) return MLP_E_ARG;
    if (hmm) build_hmm(flavour == MLP_QP ? 0.700645f : init_distrib2, hmm);
    if (part) {
        if (flavour == MLP_QP) build_part_qp(part);
        else build_part_cpnp(part);
    }
    return MLP_OK;
}

// MSA.cpp:838-881: average identity, its standard deviation, the initDistrib[2] override and the model class.
extern "C" int mlp_cpnp_model_adjustment(int64_t npairs, const int32_t* n_identical, const int32_t* align_len,
                                         float* identity_out, float* sigma_out, float* init_distrib2) {
    if (npairs < 1 || !n_identical || !align_len) return MLP_E_ARG;
    float identity = 0;
    for (int64_t k = 0; k < npairs; ++k) identity += (float)n_identical[k] / align_len[k];
    identity /= (int)npairs;
    float variance = 0;
    for (int64_t k = 0; k < npairs; ++k) {
        const float pid = (float)n_identical[k] / align_len[k];
        variance += (pid - identity) * (pid - identity);
    }
    variance /= (int)npairs;
    variance = sqrtf(variance);
    float i2 = 0.700645f;   // Defaults.h:22-23
    if (identity <= 0.125) i2 = 0.108854f;
    else if (identity <= 0.15) i2 = 0.132548f;
    else if (identity <= 0.175) i2 = 0.165248f;
    else if (identity <= 0.2) i2 = 0.168284f;
    else if (identity <= 0.25) i2 = 0.170705f;
    else if (identity <= 0.3) i2 = 0.100675f;
    else if (identity <= 0.35) i2 = 0.090755f;
    else if (identity <= 0.4) i2 = 0.146188f;
    else if (identity <= 0.45) i2 = 0.167858f;
    else if (identity <= 0.5) i2 = 0.250769f;
    if (identity_out) *identity_out = identity;
    if (sigma_out) *sigma_out = variance;
    if (init_distrib2) *init_distrib2 = i2;
    const int vm = (variance > 0.115) ? 10 : 0;
    if (identity <= 0.18) return vm + 0;
    if (identity <= 0.25) return vm + 1;
    if (identity <= 0.4) return vm + 2;
    if (identity <= 0.7) return vm + 3;
    return vm + 4;
}

// MSA::Alter_ModelAdjustmentTest, MSA.cpp:646-762 (the -G line MLProbs' first classifier reads).
extern "C" int mlp_cpnp_g_features(int n, const int32_t* len, const uint8_t* residues, const char* aln, const int64_t* aln_off,
                                   float theta, char* line, int line_cap) {
    if (n < 2 || !len || !residues || !aln || !aln_off || !line || line_cap < 64) return MLP_E_ARG;
    const char* al = MLP_HMM_ALPHABET;
    int idx[26];
    for (int k = 0; k < 26; k++) idx[k] = -1;
    for (int k = 0; k < 20; k++) idx[al[k] - 'A'] = k;
    std::vector<long long> off(n);
    long long tot = 0;
    for (int i = 0; i < n; i++) { off[i] = tot; tot += len[i]; }
    for (long long k = 0; k < tot; k++) if (residues[k] < 'A' || residues[k] > 'Z') return MLP_E_ARG;
    // Letters outside the 20-letter alphabet (B, J, O, U, X, Z): the reference indexes BLOSUM62[find(c1)][find(c2)] with
    // string::npos and reads whatever the linker placed in front of that array (in a build of the unmodified sources here:
    // a few unrelated int/bool globals, i.e. denormals or small numbers, all < 10 and therefore added).  That has no
    // defined value to reproduce; a column pair with such a letter contributes 0 here.  Fields 1-4 and 7 of the line do not
    // depend on it and stay exact; fields 5 and 6 agree with the compiled reference to about 1e-3 relative on the bundled
    // families that hold such letters (tests/test_host.py).
    auto blosum_as_compiled = [&](int i1, int i2) -> float { return (i1 < 0 || i2 < 0) ? 0.0f : MLP_BLOSUM62[i1 * 20 + i2]; };
    const int npairs = n * (n - 1) / 2;
    float identity = 0, tmp_sp = 0;
    int avg_length = 0, max_len = 0, tmp_sp_idx = 0;
    std::vector<float> arr(10000, 0.0f), pids(npairs);          // MAX_ARR, MSA.cpp:17
    int p = 0;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++, p++) {
            const uint8_t* s1 = residues + off[a]; const uint8_t* s2 = residues + off[b];
            const char* s = aln + aln_off[p];
            const int alen = (int)(aln_off[p + 1] - aln_off[p]);
            if (alen > 10000) return MLP_E_UNSUPPORTED;         // the reference overruns its fixed array here
            avg_length += alen;
            if (alen > max_len) max_len = alen;
            int i = 1, j = 1, num_idx = 0;
            float nc = 0;
            for (int k = 0; k < alen; k++) {
                if (s[k] == 'B') {
                    const uint8_t c1 = s1[i - 1], c2 = s2[j - 1]; i++; j++;
                    if (c1 == c2) nc += 1;
                    const float bl = blosum_as_compiled(idx[c1 - 'A'], idx[c2 - 'A']);
                    if (bl < 10) { arr[num_idx] += bl; tmp_sp += bl; }
                } else if (s[k] == 'X') i++;
                else j++;
                tmp_sp_idx += 1; num_idx++;
            }
            pids[p] = nc / alen;
            identity += nc / alen;
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
    const float factor = 2 * (float)n - (float)avg_length;
    std::snprintf(line, line_cap, "%f\t%f\t%d\t%d\t%f\t%f\t%f", identity, variance, n, avg_length, tmp_sp, peak, factor);
    return MLP_OK;
}

// MLProbs' Python-side column scoring (utils/calculate_column_scores.py:37-82,123-137): per column the
// BLOSUM62 sum over all sequence pairs (letters outside the 20 standard ones and gaps score 0) divided by N(N-1)/2, then the
// mean over columns, the standard deviation around it and the fraction of columns scoring >= 1.  The Python loop is
// O(columns * N^2); the pair sum only depends on the letter counts of the column, so this is O(columns * (N + 400)) with
// the same value: every partial sum there is an integer, exactly representable in a double, and the divisions, the running
// double sums (column order) and pow(x, 2) are the ones CPython performs.
extern "C" int mlp_column_scores(int n, int columns, const char* rows, double* col_score, double* mean_out, double* sd_out, double* peak_ratio_out) {
    if (n < 2 || columns < 0 || !rows) return MLP_E_ARG;
    const char* al = MLP_HMM_ALPHABET;                 // "ARNDCQEGHILKMFPSTWYV", the order of tmp_str
    int idx[256];
    for (int k = 0; k < 256; k++) idx[k] = -1;
    for (int k = 0; k < 20; k++) idx[(unsigned char)al[k]] = k;
    const double pairs = ((double)n * (double)(n - 1)) / 2;
    std::vector<double> local;
    if (!col_score) { local.resize((size_t)columns); col_score = local.data(); }
    double sum = 0.0;
    for (int c = 0; c < columns; c++) {
        long long cnt[20] = {0};
        for (int i = 0; i < n; i++) { const int k = idx[(unsigned char)rows[(size_t)i * columns + c]]; if (k >= 0) cnt[k]++; }
        long long s = 0;
        for (int a = 0; a < 20; a++) {
            if (!cnt[a]) continue;
            s += (long long)MLP_BLOSUM62[a * 20 + a] * (cnt[a] * (cnt[a] - 1) / 2);
            for (int b = a + 1; b < 20; b++) s += (long long)MLP_BLOSUM62[a * 20 + b] * cnt[a] * cnt[b];
        }
        double v = (double)s;
        v /= pairs;
        sum += v;
        col_score[c] = v;
    }
    double mean = 0.0, sd = 0.0, ratio = 0.0;
    if (columns != 0) {
        mean = sum / columns;
        for (int c = 0; c < columns; c++) sd += std::pow(col_score[c] - mean, 2.0);
        sd /= columns;
        sd = std::sqrt(sd);
        for (int c = 0; c < columns; c++) if (col_score[c] >= 1.0) ratio += 1;
        ratio = ratio / columns;
    }
    if (mean_out) *mean_out = mean;
    if (sd_out) *sd_out = sd;
    if (peak_ratio_out) *peak_ratio_out = ratio;
    return MLP_OK;
}
