/* mlprobs_b200 -- C ABI of the B200-native all-pairs posterior + consistency engine.
 *
 * This is the drop-in boundary for the ONE hot path of kuangmeng/MLProbs' two aligners: every entry point
 * below replaces a reference call site (cited per function; cpnp = baseMSA/C_P_NP_Aln,
 * QP = realign/QuickProbs/src/Alignment).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions: every function returns 0 on success or a negative MLP_E_* code and never throws;
 * nothing is printed to stdout/stderr (the callers capture both streams); the caller owns
 * host buffers, the library owns device memory; one context per process/GPU, calls serialised by the caller.
 * There is NO CPU fallback: without a CUDA device mlp_create fails with MLP_E_NO_DEVICE.
 */
#ifndef MLPROBS_B200_H
#define MLPROBS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlp_ctx mlp_ctx;

enum { MLP_OK = 0, MLP_E_NO_DEVICE = -1, MLP_E_CUDA = -2, MLP_E_ARG = -3, MLP_E_STATE = -4,
       MLP_E_CAPACITY = -5, MLP_E_OVERFLOW = -6, MLP_E_UNSUPPORTED = -7, MLP_E_NCCL = -8, MLP_E_NOMEM = -9 };

/* flavour = whose arithmetic is reproduced */
enum { MLP_QP = 0,        /* quickprobs: PosteriorStage.cpp:58-196, ConsistencyStage.cpp:133-300            */
       MLP_CPNP_P0 = 1,   /* c_p_np_aln -p 0: MSA.cpp:895-1051 (pdoAlign pair loop + DoRelaxation)          */
       MLP_CPNP_P1 = 2 }; /* c_p_np_aln -p 1: MSA.cpp:1636-1765 (ArrangePosteriorProbs) + DoRelaxation      */
/* model mask (cpnp: pid<=1 -> all three, pid==2 -> LOCAL, pid>=3 -> PART; QP: always HMM5|PART) */
enum { MLP_M_HMM5 = 1, MLP_M_PART = 2, MLP_M_LOCAL = 4 };

/* Log-space pair-HMM tables, index = letter - 'A'.
 * cpnp ProbabilisticModel.h:42-47,58-135 ; QP ProbabilisticModel.h:43-46, ProbabilisticModel.cpp:15-56 */
typedef struct {
    float init[5];
    float trans[5][5];
    float match[26][26];
    float ins[26];
    float ltrans[3][3];   /* cpnp local model, ProbabilisticModel.h:46,103-127 */
    float rtrans[2];      /* cpnp flanking random states, ProbabilisticModel.h:47,130-131 */
} mlp_hmm_tables;

/* Partition-function tables, index = letter - 'A' (NaN = letter the reference cannot score).
 * cpnp MSAReadMatrix.cpp:85-116 + MSAPartProbs.cpp:698-709 ; QP ExpPartitionFunctionParams.h:17-49 */
typedef struct {
    double sub[26][26];
    double go, ge, tgo, tge;
} mlp_part_tables;

/* Build the reference's default tables on the host with glibc logf/expf/exp.
 * init_distrib2 is cpnp's identity-dependent initDistrib[2] (MSA.cpp:861-870); ignored for MLP_QP. */
int mlp_default_tables(int flavour, float init_distrib2, mlp_hmm_tables* hmm, mlp_part_tables* part);

/* Context on one CUDA device (replaces the OpenMP thread team of MSA.cpp:146-152 / PosteriorStage.cpp:67). */
int mlp_create(int device, mlp_ctx** out);
void mlp_destroy(mlp_ctx* ctx);
const char* mlp_last_error(const mlp_ctx* ctx);
/* scratch budget for dense DP layers in bytes (0 = 60 % of free device memory) and sparse pool capacity in cells (0 = auto) */
int mlp_configure(mlp_ctx* ctx, int64_t scratch_bytes, int64_t cell_capacity);

int mlp_set_tables(mlp_ctx* ctx, const mlp_hmm_tables* hmm, const mlp_part_tables* part);
/* n sequences, upper-case letters 'A'..'Z', concatenated (what MultiSequence::LoadMFA leaves: cpnp
 * MultiSequence.h:98-111 / Sequence.h:96-112; QP SequenceIO.cpp:98-155). */
int mlp_set_sequences(mlp_ctx* ctx, int n, const int32_t* len, const uint8_t* residues);
/* Restrict the posterior / relax stages of THIS context to pairs p with p % world == rank of the
 * cost-sorted pair list (multi-GPU sharding). Default rank 0 of 1. */
int mlp_set_shard(mlp_ctx* ctx, int rank, int world);

/* Host-only helper (no GPU): the pairs a shard owns, in the order the device processes them. pairs_out receives
 * 2*count ints (a0,b0,a1,b1,...); pass NULL to query *count. Same rule as mlp_set_shard. */
int mlp_shard_pairs(int n, const int32_t* len, int rank, int world, int32_t* pairs_out, int64_t* count);

/* All-pairs 3-state Viterbi alignment (model selection): for every pair in row-major a<b order
 * the alignment length and the number of match columns with identical residues.
 * Replaces the pair loop of MSA::ModelAdjustmentTest MSA.cpp:801-837 (ProbabilisticModel::ComputeViterbiAlignment
 * ProbabilisticModel.h:1043-1170).  Uses the HMM tables of mlp_set_tables (local transitions + emissions). */
int mlp_viterbi_all_pairs(mlp_ctx* ctx, int32_t* n_identical, int32_t* align_len);
/* Same, additionally returning every pair's Viterbi alignment as a B/X/Y string (what ComputeViterbiAlignment returns):
 * aln receives the strings back to back, aln_off[p]..aln_off[p+1] is pair p (aln_off has npairs+1 entries); aln must hold
 * sum over pairs of len[a]+len[b] bytes. */
int mlp_viterbi_all_pairs_ex(mlp_ctx* ctx, int32_t* n_identical, int32_t* align_len, char* aln, int64_t* aln_off);
/* The `c_p_np_aln -G` feature line (MSA::Alter_ModelAdjustmentTest MSA.cpp:646-762) from the Viterbi alignments, host only,
 * one-core summation order: "identity\tsigma\tN\tavgLen\tavgSP\tpeakRatio\tfactor" with std::to_string formatting.
 * A residue outside the 20 standard letters makes the reference index BLOSUM62 with string::npos and add whatever lies in
 * front of that array (observed in a build of its sources: anything from denormals to -4.9e13 in field 5); such a column
 * pair contributes 0 here, which leaves fields 1-4 and 7 exact and fields 5-6 within about 1e-3 of the reference where the
 * reference's own value is sane.  MLP_E_UNSUPPORTED for an alignment longer than the reference's fixed 10,000-entry array. */
int mlp_cpnp_g_features(int n, const int32_t* len, const uint8_t* residues, const char* aln, const int64_t* aln_off,
                        float theta, char* line, int line_cap);
/* Host part of ModelAdjustmentTest (MSA.cpp:838-881): sequential one-core float sums in pair order.
 * Returns variance_mean = pid + (sigma > 0.115 ? 10 : 0), pid in 0..4; init_distrib2 = the overridden initDistrib[2]. */
int mlp_cpnp_model_adjustment(int64_t npairs, const int32_t* n_identical, const int32_t* align_len,
                              float* identity, float* sigma, float* init_distrib2);

/* All-pairs posterior stage: for every owned pair a<b -> dense posteriors of the selected models, merge,
 * MEA score -> distance, threshold to CSR (both orientations).
 * Replaces cpnp MSA.cpp:927-1031 (and :1652-1765 for -p 1) / QP PosteriorStage::run PosteriorStage.cpp:58-121. */
int mlp_posterior_all_pairs(mlp_ctx* ctx, int flavour, uint32_t model_mask, float cutoff);

/* n*n float distances (row-major, symmetric, 0 diagonal): cpnp MSA.cpp:1019 / QP PosteriorStage.cpp:107,113 */
int mlp_get_distances(mlp_ctx* ctx, float* nxn);

/* One consistency repetition over all owned pairs.
 * cpnp: MSA::DoRelaxation MSA.cpp:1172-1281 (weights/seldist NULL, unweighted, /N).
 * QP:   ConsistencyStage::doRelaxation ConsistencyStage.cpp:133-266 with the default Max/Deterministic
 *       selectivity (accept z iff max(seldist[i][z], seldist[j][z]) <= selectivity), weights from the guide tree;
 *       weights = seldist = NULL takes both from the preceding mlp_qp_guide_tree_device. */
int mlp_relax(mlp_ctx* ctx, int flavour, const float* weights, const float* seldist_nxn,
              float selectivity, float selfweight, float cutoff);

/* Streamed posterior stage for families whose sparse set does not fit HBM (BASELINE config #5: 4,000 x 500, 16 M matrices; the
 * reference's own GPU path streams relaxation sectors for the same reason, KernelAlignment/Multiple/QuickConsistencyStage.cpp:163-190).
 * Between mlp_stream_begin and mlp_stream_end, mlp_posterior_all_pairs(MLP_QP) finishes every batch of pairs on the spot: the cells get
 * the re-quantisation a consistency repetition applies to a matrix whose pair accepts no third sequence (ConsistencyStage.cpp:
 * 213-258 with sumW = 1; reps must be 1, QuickProbs' own count above 50 sequences), the per-matrix digest (format of mlp_set_digest) is accumulated, and the cell pool is
 * reused by the next batch.  Afterwards the distances are complete and mlp_qp_guide_tree_device can run; the pairs that DO accept third
 * sequences (subtree distance <= selectivity: a few per cent of a large family) are then recomputed by an ordinary stage restricted
 * to them with mlp_restrict_pairs, exchanged (mlp_exchange_needed) and relaxed as usual; mlp_set_shard restores the full shard.
 * mlp_stream_end copies the digests of the streamed matrices to per_matrix_nn (n*n, may be NULL). */
int mlp_stream_begin(mlp_ctx* ctx, int reps);
int mlp_stream_end(mlp_ctx* ctx, uint64_t* per_matrix_nn);
int mlp_restrict_pairs(mlp_ctx* ctx, const float* seldist_nxn, float selectivity);
/* Host utility (no GPU work): the pairs of mlp_shard_pairs(rank, world) that mlp_restrict_pairs keeps; pairs_out may be NULL (count only). */
int mlp_shard_pairs_within(int n, const int32_t* len, int rank, int world, const float* seldist_nxn, float selectivity,
                           int32_t* pairs_out, int64_t* count);

/* QuickProbs' guide tree ON THE DEVICE, from the distance matrix the posterior stage left in HBM (which stays untouched):
 * UPGMA clustering (ClusterTree::build ClusterTree.cpp:17-124), normalised sequence weights (GuideTree::calculateSeqsWeights
 * GuideTree.cpp:114-154) raised to at least min_weight (ExtendedMSA.cpp:237-238 saturates them at 1e-6; pass 0 for the raw
 * weights) and the subtree-size selectivity distances (GuideTree::calculateSubtreeDistances GuideTree.cpp:189-221).
 * Bit-identical to mlp_qp_guide_tree_ex on the same matrix.  The weights and the selectivity distances stay resident: a following
 * mlp_relax(ctx, MLP_QP, NULL, NULL, ...) uses them without a host round trip.  weights_out: n floats; parent / left / right:
 * 2n-1 ints (leaves 0..n-1, inner nodes in merge order, root last), optional; seldist_out: n*n floats, optional (mlp_exchange_needed
 * wants them on the host).  After a sharded stage call mlp_exchange_distances first.  MLP_E_UNSUPPORTED beyond ~4,800 sequences
 * (the row minima and tree arrays of the single-CTA clustering live in shared memory): use the host function there. */
int mlp_qp_guide_tree_device(mlp_ctx* ctx, float min_weight, float* weights_out, int32_t* parent_out, int32_t* left_out,
                             int32_t* right_out, float* seldist_out);

/* Host utility of the QuickProbs flavour (no GPU work): UPGMA guide tree on the n*n distances (updated IN PLACE, as the
 * reference does), normalised sequence weights and subtree-size selectivity distances for mlp_relax.
 * Replaces ClusterTree::build ClusterTree.cpp:17-124, GuideTree::calculateSeqsWeights GuideTree.cpp:114-154 and
 * GuideTree::calculateSubtreeDistances GuideTree.cpp:189-221.  parent_out (2n-1 ints, -1 = root) may be NULL. */
int mlp_qp_guide_tree(int n, float* dist_nxn_inout, float* weights_out, float* subtree_dist_nxn_out, int32_t* parent_out);
/* Same, additionally returning the children of every internal node v (n <= v < 2n-1; entries below n are -1): left = the
 * cluster found first by the reference's scan (connectNodes' leftChild, ClusterTree.cpp:88-89), right = the other one.
 * The root is node 2n-2.  This is the tree mlp_qp_finish_alignment* walks. */
int mlp_qp_guide_tree_ex(int n, float* dist_nxn_inout, float* weights_out, float* subtree_dist_nxn_out, int32_t* parent_out,
                         int32_t* left_out, int32_t* right_out);

/* QuickProbs flavour, the stages after consistency: progressive construction along the guide tree and column-based
 * iterative refinement, giving the final multiple alignment.
 * Replaces ConstructionStage::operator() ConstructionStage.cpp:11-127 and RefinementBase::operator() RefinementBase.cpp:13-63
 * with ColumnRefinement.cpp (called from ExtendedMSA::doAlign ExtendedMSA.cpp:235-252), default configuration
 * (finalSelectivity = FLT_MAX, acceptance by length, no recursion).  weights = tree weights saturated at 1e-6;
 * ref_iters 0 or -1 selects the reference's default pass count (as `-r 0` does there), -2 skips refinement (construction
 * only; an extension used by tests), ref_seed 0 keeps its default-seeded std::mt19937.
 * rows_out receives a malloc'ed n x aln_len byte matrix (input order, '-' for gaps) to release with mlp_free_host.
 *   _host: the sparse set is a host copy in the pooled layout of mlp_csr_layout/mlp_get_csr_raw (no GPU work);
 *   the ctx variant sums the pair matrices with a CUDA kernel over the set resident in HBM (no read-back of the set). */
int mlp_qp_finish_alignment_host(int n, const int32_t* len, const uint8_t* residues, const float* weights,
                                 const int32_t* left, const int32_t* right, const int64_t* rp_off, const int64_t* nz_off,
                                 const int32_t* rp_pool, const void* cells, int ref_iters, uint32_t ref_seed,
                                 char** rows_out, int32_t* aln_len);
int mlp_qp_finish_alignment(mlp_ctx* ctx, const float* weights, const int32_t* left, const int32_t* right,
                            int ref_iters, uint32_t ref_seed, char** rows_out, int32_t* aln_len);
void mlp_free_host(void* p);

/* c_p_np_aln flavour, the stages after consistency for `-p 0` (progressive): guide tree, progressive alignment with
 * weighted profile posteriors, iterative refinement by random bipartition.
 * mlp_cpnp_guide_tree replaces MSAClusterTree::create(vpid) (MSAClusterTree.cpp:30-39,170-296: UPGMA, plain average of the
 * two joined rows when variance_id == 0, size-weighted otherwise; distances updated IN PLACE) and
 * MSAGuideTree::getSeqsWeights (MSAGuideTree.cpp:274-322: integer weights normalised to 1000, at least 1).
 * mlp_cpnp_finish_alignment* replace MSA::ComputeFinalAlignment (MSA.cpp:1481-1534) with ProcessTree / AlignAlignments
 * (:1369-1471) and DoIterativeRefinement (:1537-1623).  refine_reps = -ir (reference default 100), pid = model class
 * (variance_mean % 10).  The reference draws its bipartitions from an unseeded glibc rand(); the library carries a private
 * replica of that generator (seed 1).  With more than one OpenMP thread the reference's refinement races on the shared
 * posterior and its output changes from run to run; this is its one-thread result.
 * rows_out: malloc'ed n x aln_len matrix in the reference's OUTPUT order (refinement never re-sorts the rows);
 * order_out[k] = input index of row k (may be NULL).  Release rows_out with mlp_free_host. */
int mlp_cpnp_guide_tree(int n, float* dist_nxn_inout, int variance_id, int32_t* weights_out, int32_t* left_out, int32_t* right_out);
int mlp_cpnp_finish_alignment_host(int n, const int32_t* len, const uint8_t* residues, const int32_t* iweights,
                                   const int32_t* left, const int32_t* right, const int64_t* rp_off, const int64_t* nz_off,
                                   const int32_t* rp_pool, const void* cells, int refine_reps, int pid,
                                   char** rows_out, int32_t* aln_len, int32_t* order_out);
int mlp_cpnp_finish_alignment(mlp_ctx* ctx, const int32_t* iweights, const int32_t* left, const int32_t* right,
                              int refine_reps, int pid, char** rows_out, int32_t* aln_len, int32_t* order_out);
/* c_p_np_aln -p 1 (the non-progressive strategy), everything after the relaxation of MSA::npdoAlign (MSA.cpp:1084-1160):
 * MSA::ComputeGraph (MSA.cpp:1777-1845) with the alignment graph of AlignGraph.h (cells of the a<b matrices sorted by its own
 * quicksort, greedy insertion strongest first with its cycle tests, Graph2Align) and MSA::DoRefinement (MSA.cpp:1852-1980:
 * FindSimilar's two-means sets, each sequence re-aligned to its similar set and the set to the rest, unweighted profile
 * posteriors; skipped above 150 sequences).  The sparse set must come from mlp_posterior_all_pairs(MLP_CPNP_P1, ...) +
 * mlp_relax; distances = its distance matrix (read only).  refine_reps = -ir (reference default 100).
 * The reference calls srand(time(0)) before every refinement sweep, so its output depends on the wall clock; seed < 0 does
 * the same, seed >= 0 uses that value instead of the clock (what `oracle/_ref/ref_cpnp msa --p1 --fixtime T` pins).
 * rows_out: malloc'ed n x aln_len matrix, rows in input order (AlignAlignments re-sorts by label); mlp_free_host releases it.
 *   _host: sparse set given as a host copy in the pooled layout of mlp_csr_layout / mlp_get_csr_raw (no GPU work);
 *   device: the graph is built on the host from a read-back of the set, the refinement runs on the resident set. */
int mlp_cpnp_np_finish_alignment_host(int n, const int32_t* len, const uint8_t* residues, const float* distances,
                                      const int64_t* rp_off, const int64_t* nz_off, const int32_t* rp_pool, const void* cells,
                                      int refine_reps, int64_t seed, char** rows_out, int32_t* aln_len);
int mlp_cpnp_np_finish_alignment(mlp_ctx* ctx, int refine_reps, int64_t seed, char** rows_out, int32_t* aln_len);
/* MLProbs' Python-side column scores of an alignment (utils/calculate_column_scores.py:37-82 calculateColScore, :123-137
 * getSD / getPeakLengthRatio): rows = n x columns bytes.  col_score (may be NULL) receives the per-column values; the mean
 * over columns, the standard deviation and the fraction of columns >= 1 go to the three outputs.  Host only; same doubles
 * as the Python code, computed from letter counts instead of the O(N^2) pair loop. */
int mlp_column_scores(int n, int columns, const char* rows, double* col_score, double* mean, double* sd, double* peak_ratio);
/* test hook: first `count` outputs of the private glibc rand() replica */
int mlp_debug_glibc_rand(int count, int32_t* out);
/* test hook: permutation produced by the alignment graph's sort (AlignGraph.h:62-113 restated; tasks != 0: OpenMP tasks) */
int mlp_debug_reference_sort(int64_t n, const float* keys, int32_t* idx_out, int tasks);
/* the same after srand(seed) */
int mlp_debug_glibc_rand_seeded(uint32_t seed, int count, int32_t* out);

/* Sparse posterior read-back. Ordered pair (a,b), a != b; rows 1..len[a]; row_ptr has len[a]+2 entries
 * (row_ptr[i]..row_ptr[i+1] = row i, row 0 empty).  val is the dequantised value for MLP_QP
 * (SparseEntry.h:31-32).  Pass NULL col/val to query *nnz only. */
int mlp_get_csr(mlp_ctx* ctx, int a, int b, int32_t* row_ptr, int32_t* col, float* val, int64_t* nnz);
/* total cells currently stored over all ordered pairs */
int mlp_total_cells(mlp_ctx* ctx, int64_t* cells);
/* bulk read-back of the a<b orientation in pair order (p = row-major index of a<b): nnz[p], then concatenated
 * col/val; row_ptr concatenated with len[a]+2 entries per pair. Any pointer may be NULL. */
int mlp_get_csr_bulk(mlp_ctx* ctx, int64_t* nnz_per_pair, int32_t* row_ptr, int32_t* col, float* val);

/* Zero-reshuffle bulk read-back for callers that keep the library's own pooled layout (fastest path: four plain
 * device->host copies, use pinned buffers from mlp_alloc_pinned for full PCIe speed).
 *   mlp_csr_layout : rp_off[n*n] (offset of the len[a]+2 row pointers of ordered pair (a,b) inside the row-pointer pool),
 *                    *rp_total (ints in that pool), *cells_used (cells currently in the cell pool).
 *   mlp_get_csr_raw: nz_off[n*n], nz_cnt[n*n], rp_pool[rp_total], cells[cells_used] as {int32 column, float value}.
 * Matrix (a,b): rows i=1..len[a] -> cells[nz_off[a*n+b] + rp_pool[rp_off[a*n+b]+i] .. + rp_pool[rp_off[a*n+b]+i+1]). */
int mlp_csr_layout(mlp_ctx* ctx, int64_t* rp_off, int64_t* rp_total, int64_t* cells_used);
int mlp_get_csr_raw(mlp_ctx* ctx, int64_t* nz_off, int32_t* nz_cnt, int32_t* rp_pool, void* cells);
/* QuickProbs flavour only: the same pooled layout in the reference's own storage format (PackedSparseMatrix.h:9-80,
 * SparseEntry.h: uint16 column + uint16 fixed-point value per cell, row sizes next to row indices) -- half the bytes
 * of mlp_get_csr_raw over PCIe.  cells[k] = column | (code << 16), i.e. the bytes of SparseEntry<uint16_t, uint16_t>{first = column, second = code} on a
 * little-endian host, value = code / 65535.0f; row_sizes has the
 * rp_total entries of the row-pointer pool, entry rp_off[a*n+b] + i = number of cells in row i (row 0 and the entry
 * after the last row are 0), so rowIndices are their running sum.  Any pointer may be NULL.  Packing runs on the device
 * into the idle half of the double-buffered cell pool. */
int mlp_get_csr_packed(mlp_ctx* ctx, int64_t* nz_off, int32_t* nz_cnt, uint16_t* row_sizes, uint32_t* cells);
/* The same read-back in two halves: _begin packs the current set and enqueues the copies on a stream of its own, _end waits
 * for them.  In between the caller may submit the next family (mlp_set_sequences of the same shape) and run its
 * mlp_posterior_all_pairs: the PCIe transfer of one result hides behind the next posterior stage.  Calls that would overwrite
 * the set being read (mlp_relax, a gathering exchange, a family of another shape, ...) end the read-back themselves.
 * The host buffers (page-locked for a truly asynchronous copy, mlp_alloc_pinned) must not be touched before _end. */
int mlp_get_csr_packed_begin(mlp_ctx* ctx, int64_t* nz_off, int32_t* nz_cnt, uint16_t* row_sizes, uint32_t* cells);
int mlp_get_csr_packed_end(mlp_ctx* ctx);
int mlp_alloc_pinned(int64_t bytes, void** out);
void mlp_free_pinned(void* p);

/* Developer hook (tools/loc_ab.py with MLP_LOC_SPLIT=1): candidates / firing cells of the local model's forward and backward Z
 * chains (loc_c.cu) since the last call, out4 = {fwd candidates, fwd firing, bwd candidates, bwd firing}. */
int mlp_debug_loc_counters(mlp_ctx* ctx, unsigned long long* out4);

/* Test hook: overwrite the resident n*n distance matrix (the device guide tree is checked on tie-heavy matrices with it). */
int mlp_debug_set_distances(mlp_ctx* ctx, const float* nxn);

/* Dense per-pair debug read-back (tests): runs one pair and returns the merged dense posterior
 * (len[a]+1 x len[b]+1) and, if non-NULL, each model's posterior. */
int mlp_debug_pair_dense(mlp_ctx* ctx, int flavour, uint32_t model_mask, int a, int b,
                         float* merged, float* p_hmm5, float* p_part, float* p_local, float* distance);

/* Multi-GPU exchange (one process per GPU): all-gather the sparse posteriors + distances of every rank's
 * shard so that each context holds the complete set (needed before mlp_relax and by the host tail).
 * unique_id is the 128-byte ncclUniqueId from mlp_nccl_unique_id on rank 0, distributed by the caller. */
int mlp_nccl_unique_id(uint8_t id128[128]);
int mlp_comm_init(mlp_ctx* ctx, const uint8_t id128[128], int rank, int world);
int mlp_exchange(mlp_ctx* ctx);
/* The same exchange in two halves: _begin enqueues everything and returns; mlp_get_distances then only waits for the distance
 * all-reduce, so the host guide tree overlaps the cell broadcasts; _end (or any later stage call) waits for the rest. */
int mlp_exchange_begin(mlp_ctx* ctx);
int mlp_exchange_end(mlp_ctx* ctx);
/* Selective exchange for the QuickProbs flavour.  QuickProbs' consistency reads S_xz only when the subtree distance d[x][z] is
 * within the selectivity (ConsistencyStage.cpp:181-216), so once the guide tree exists each rank ships only those matrices:
 *   mlp_exchange_distances   all-reduce of the distance matrix (the tree needs all of it; no-op when already complete)
 *   mlp_exchange_needed      imports, on every rank, the matrices with seldist[a*n+b] <= selectivity that other ranks own
 * The following mlp_relax works on the rank's own pairs; its result stays sharded (mlp_get_csr* return the owned pairs,
 * mlp_exchange gathers everything where a tail needs the whole set). */
/* Device-side digest of the current set: per ordered matrix (a,b) of the pairs this rank owns, a 64-bit position-weighted
 * hash of its row pointers and cells (0 for matrices of other ranks), n*n values.  Summed over the ranks of a sharded run it
 * equals the single-GPU vector: bench.py and tools/multigpu_check.py use it as the N-GPU parity check. */
int mlp_set_digest(mlp_ctx* ctx, uint64_t* per_matrix_nn);
int mlp_exchange_distances(mlp_ctx* ctx);
int mlp_exchange_needed(mlp_ctx* ctx, const float* seldist_nxn, float selectivity);

/* timing / accounting of the last stage call, measured with CUDA events on the library's stream */
typedef struct {
    double ms_total;        /* whole stage on the device                         */
    double ms_kernel[8];    /* per kernel family, see MLP_K_*                    */
    int64_t launches;       /* kernels launched                                  */
    int64_t cells;          /* DP cells (L1+1)*(L2+1) summed over pairs processed */
    int64_t pairs;
    int64_t nnz;            /* sparse cells produced (a<b orientation)           */
    int64_t h2d_bytes, d2h_bytes;
} mlp_stage_stats;
enum { MLP_K_PART_FWD = 0, MLP_K_PART_REV = 1, MLP_K_HMM_FWD = 2, MLP_K_HMM_BWD = 3, MLP_K_LOCAL_FWD = 4,
       MLP_K_LOCAL_BWD = 5, MLP_K_FINAL = 6, MLP_K_RELAX = 7 };
int mlp_last_stats(mlp_ctx* ctx, mlp_stage_stats* out);

#ifdef __cplusplus
}
#endif
#endif
