/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 * CPU restatement of the reference's all-pairs posterior + consistency path, used ONLY as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Pinned against the compiled reference (oracle/_ref/ref_cpnp, oracle/_ref/ref_qp) through the
 * golden vectors in tests/golden/ (tests/test_oracle_golden.py).
 * Every function cites the reference file:line it follows (cpnp = baseMSA/C_P_NP_Aln,
 * QP = realign/QuickProbs/src/Alignment).  Sequences are passed as upper-case letters, 0-based.
 */
#ifndef MLP_ORACLE_H
#define MLP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float init[5];        /* log initial distribution                         */
    float trans[5][5];    /* log transitions (only [0][q],[q][q],[k][0] used)  */
    float match[26][26];  /* log match emissions, index letter-'A'             */
    float ins[26];        /* log insert emissions                              */
    float ltrans[3][3];   /* local 3-state transitions (cpnp only)             */
    float rtrans[2];      /* flanking random-state transitions (cpnp only)     */
} orc_hmm_tables;

typedef struct {
    double sub[26][26];   /* exp(beta*score), index letter-'A'; NaN = letter unsupported */
    double go, ge, tgo, tge;
} orc_part_tables;

/* flavours */
enum { ORC_QP = 0, ORC_CPNP_P0 = 1, ORC_CPNP_P1 = 2 };
/* model masks */
enum { ORC_M_HMM5 = 1, ORC_M_PART = 2, ORC_M_LOCAL = 4 };

void orc_build_hmm(float init_distrib2, orc_hmm_tables* t);          /* ProbabilisticModel.h:58-135, PairHmm.cpp:4-33 */
void orc_build_part_cpnp(orc_part_tables* t);                        /* MSAReadMatrix.cpp:85-116,158-210; MSAPartProbs.cpp:698-709 */
void orc_build_part_qp(orc_part_tables* t);                          /* ExpPartitionFunctionParams.h:30-49 */
float orc_init_distrib2_for_identity(float identity);                /* MSA.cpp:861-870 */

float orc_log_add(float x, float y);                                 /* ScoreType.h:279-285 */
float orc_exp(float x);                                              /* ScoreType.h:36-68  */

/* dense posteriors, row-major (L1+1)x(L2+1) */
void orc_hmm5_posterior(const orc_hmm_tables* t, const char* s1, int L1, const char* s2, int L2,
                        int qp_total_quirk, float* post, float* total);
void orc_local_posterior(const orc_hmm_tables* t, const char* s1, int L1, const char* s2, int L2,
                         float* post, float* total);
int  orc_part_posterior_qp(const orc_part_tables* t, const char* s1, int L1, const char* s2, int L2, float* post);
int  orc_part_posterior_cpnp(const orc_part_tables* t, const char* s1, int L1, const char* s2, int L2, float* post);

/* merge + MEA score */
float orc_combine_qp(int L1, int L2, const float* hmm, const float* part, float* out);     /* PosteriorStage.cpp:156-196 */
void  orc_merge3_cpnp(int n, const float* p5, const float* pp, const float* pl, int p1_order, float* out); /* MSA.cpp:992-1007 / 1699-1714 */
float orc_mea_score(int L1, int L2, const float* post, int* n_match);                      /* ProbabilisticModel.h:804-864 */

/* 3-state Viterbi alignment used for model selection (ProbabilisticModel.h:1043-1170): returns the log probability, the
 * alignment length and the number of 'B' columns with identical residues; aln (may be NULL) receives the B/X/Y string. */
float orc_viterbi(const orc_hmm_tables* t, const char* s1, int L1, const char* s2, int L2, int* n_identical, int* aln_len, char* aln);
/* ModelAdjustmentTest (MSA.cpp:775-882) from the per-pair counts in pair order (one-core summation order):
 * returns variance_mean (pid + 10 if sigma > 0.115) and writes identity, sigma, the overridden initDistrib[2]. */
int orc_model_adjustment(int npairs, const int32_t* n_identical, const int32_t* aln_len, float* identity, float* sigma, float* init_distrib2);

/* the `c_p_np_aln -G` feature line (MSA::Alter_ModelAdjustmentTest, MSA.cpp:646-762), one-core summation order.
 * Returns 0, or 1 if a sequence holds a letter outside the 20 standard ones (the reference indexes out of bounds there). */
int orc_g_features(const orc_hmm_tables* t, int n, const int32_t* len, const char* residues, const int64_t* res_off,
                   float theta, char* line, int cap);

/* one pair end to end: returns dense posterior and distance */
int orc_pair_posterior(int flavour, int model_mask, const orc_hmm_tables* ht, const orc_part_tables* pt,
                       const char* s1, int L1, const char* s2, int L2, float* post, float* dist);

/* sparse: rowptr has L1+2 entries (rowptr[i]..rowptr[i+1] is row i, row 0 empty) */
int64_t orc_sparsify(int L1, int L2, const float* post, float cutoff, int quantize_u16,
                     int32_t* rowptr, int32_t* col, float* val, int64_t cap);              /* SparseMatrix.h:55-98; PackedSparseMatrix.cpp:40-83 */
void orc_transpose(int L1, int L2, const int32_t* rowptr, const int32_t* col, const float* val,
                   int32_t* t_rowptr, int32_t* t_col, float* t_val);                       /* SparseMatrix.h:205-248; PackedSparseMatrix.cpp:93-140 */

/* whole family. CSR set layout: for ordered pair (a,b), a!=b, slot = a*n+b; off[slot] is the start of its
 * rowptr block inside rowptr_pool (len[a]+2 ints) and nz_off[slot] the start of its cells in col/val pools.
 * Both orientations are stored (transposes are value copies). */
typedef struct {
    int n;
    const int32_t* len;
    int64_t* rp_off;   /* n*n */
    int64_t* nz_off;   /* n*n */
    int32_t* rowptr;   /* pool */
    int32_t* col;      /* pool */
    float* val;        /* pool */
    int64_t rp_cap, nz_cap, rp_used, nz_used;
} orc_csr_set;

/* all-pairs posterior stage: fills dist (n*n) and the set; threads = OpenMP threads */
int orc_posterior_stage(int flavour, int model_mask, const orc_hmm_tables* ht, const orc_part_tables* pt,
                        int n, const int32_t* len, const char* residues, const int64_t* res_off,
                        float cutoff, float* dist, orc_csr_set* out, int threads);
/* one consistency repetition, cpnp (MSA.cpp:1172-1281) or QP (ConsistencyStage.cpp:133-266) */
int orc_relax_cpnp(const orc_csr_set* in, float cutoff, orc_csr_set* out, int threads);
int orc_relax_qp(const orc_csr_set* in, const float* weights, const float* seldist, float selectivity,
                 float selfweight, float cutoff, orc_csr_set* out, int threads);

#ifdef __cplusplus
}
#endif
#endif
