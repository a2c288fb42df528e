// Internal: the QuickProbs flavour's tail (progressive construction + column refinement) on top of the pooled sparse set.
// Reference path: ExtendedMSA::doAlign ExtendedMSA.cpp:235-252 -> ConstructionStage.cpp:51-127 -> ColumnRefinement.cpp.
// The control flow (tree walk, MEA traceback, gap insertion, column scoring, RNG) is host code shared by both posterior
// providers; only "sum the pair matrices of two profiles into one dense matrix" differs: HostProfilePosterior walks a
// host copy of the pooled CSR, DeviceProfilePosterior (qp_tail_dev.cu) runs a kernel over the set resident in HBM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace qptail {

struct Profile {
    std::vector<int> ids;             // sequence labels (input order indices)
    std::vector<std::string> rows;    // gapped rows, 0-based columns, all of equal length
    int length() const { return rows.empty() ? 0 : (int)rows[0].size(); }
    int count() const { return (int)ids.size(); }
};

// How a pair (a in A, b in B) is weighted inside a profile-profile posterior.
//   QP_DOUBLE     : w = (float)((double)wf[a] * wf[b] / total), total = double sum over all pairs   (ParallelProbabilisticModel.cpp:318-360)
//   CPNP_INT      : w = (float)(wi[a] * wi[b]) / total, total = FLOAT running sum of the int products (cpnp ProbabilisticModel.h:1304-1331)
//   UNWEIGHTED    : w = 1                                                                             (cpnp ProbabilisticModel.h:1197-1285)
struct WeightSpec {
    enum Mode { QP_DOUBLE = 0, CPNP_INT = 1, UNWEIGHTED = 2 };
    int mode = QP_DOUBLE;
    const float* wf = nullptr;
    const int32_t* wi = nullptr;
    double weight_of(int id) const { return mode == QP_DOUBLE ? (double)wf[id] : (mode == CPNP_INT ? (double)wi[id] : 1.0); }
    // the reference's normaliser, accumulated in its order and precision (returned as double; exact for the float case)
    double total(const Profile& A, const Profile& B) const;
    float pair_weight(double wa, double wb, double tot) const {
        if (mode == QP_DOUBLE) return (float)((wa * wb) / tot);
        if (mode == CPNP_INT) return (float)(wa * wb) / (float)tot;      // int product (exact in float) / float total
        return 1.0f;
    }
};

// Produces the weighted sum of the pairwise sparse posteriors of every (a in A, b in B), scattered through the
// profiles' column mappings, as a dense (lenA+1) x (lenB+1) row-major float matrix (row/column 0 unused, zero).
struct ProfilePosterior {
    virtual ~ProfilePosterior() {}
    // returns 0 or a negative MLP_E_* code; *dense points at a provider-owned buffer valid until the next call
    virtual int build(const Profile& A, const Profile& B, const WeightSpec& ws, const float** dense) = 0;
    // optional fused path: MEA traceback string directly (returns 1 = not supported, caller runs the host DP on `dense`)
    virtual int build_and_align(const Profile& A, const Profile& B, const WeightSpec& ws, std::string& path) { (void)A; (void)B; (void)ws; path.clear(); return 1; }
    // optional fused path for cpnp's refinement: MEA path, its score (last cell of the DP) and the dense values at the given
    // element offsets (the posterior mass on the old alignment's columns); returns 1 = not supported
    virtual int build_align_score(const Profile& A, const Profile& B, const WeightSpec& ws, const std::vector<long long>& offsets,
                                  std::string& path, float* score, std::vector<float>& values) {
        (void)A; (void)B; (void)ws; (void)offsets; (void)score; (void)values; path.clear(); return 1;
    }
};

struct HostCsrView {
    int n;
    const int32_t* len;
    const int64_t* rp_off;
    const int64_t* nz_off;
    const int32_t* rp_pool;
    const void* cells;                // {int32 column, float value}
};

ProfilePosterior* make_host_provider(const HostCsrView& v);

struct TailOptions {
    int ref_iters = -1;               // -r / --ref-count; 0/-1: 30 passes up to 200 sequences, 200 above (RefinementBase.cpp:33-36); -2: none
    uint32_t ref_seed = 0;            // --ref-seed; 0 keeps std::mt19937's default seed (ColumnRefinement.h:13-15)
};

// residues: concatenated upper-case letters. left/right: children of node v (n <= v < 2n-1), root = 2n-2.
int run_tail(int n, const int32_t* len, const uint8_t* residues, const float* weights, const int32_t* left, const int32_t* right,
             ProfilePosterior& prov, const TailOptions& opt, Profile& out, std::string& err);

// c_p_np_aln -p 0 (cpnp flavour): ProcessTree with weighted profile posteriors, then DoIterativeRefinement's random
// bipartitions (MSA.cpp:1369-1623).  iweights: MSAGuideTree's integer weights; refine_reps: -ir (default 100); pid: model
// class (variance_mean % 10).  `out` keeps the reference's row order (refinement never re-sorts), ids = input indices.
int run_cpnp_tail(int n, const int32_t* len, const uint8_t* residues, const int32_t* iweights, const int32_t* left, const int32_t* right,
                  ProfilePosterior& prov, int refine_reps, int pid, Profile& out, std::string& err);

// c_p_np_aln -p 1 (non-progressive, MSA::npdoAlign MSA.cpp:1084-1160 after the relaxation): the alignment graph over the
// a<b matrices of `graph_set` (cpnp_graph.cpp), then MSA::DoRefinement's similar-set re-alignments through `prov`.
// distances: n x n (needed when refine_reps > 0 and n <= 150); seed < 0: time(0) before every sweep, as the reference does;
// seed >= 0: that value instead (what the tests and `ref_cpnp --fixtime` use).  Rows of `out` are in input order.
int build_graph_alignment(const HostCsrView& graph_set, const uint8_t* residues, Profile& out, std::string& err);
int run_cpnp_np_tail(const HostCsrView& graph_set, const uint8_t* residues, const float* distances, ProfilePosterior& prov,
                     int refine_reps, long long seed, Profile& out, std::string& err);

// MEA traceback over a dense profile posterior (ProbabilisticModel::computeAlignment, ProbabilisticModel.cpp:345-421);
// *score (may be NULL) receives the maximum sum, i.e. the last cell of the last row
std::string mea_path(int l1, int l2, const float* dense, float* score = nullptr);

}  // namespace qptail
