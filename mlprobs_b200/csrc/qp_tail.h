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

// Produces the weighted sum of the pairwise sparse posteriors of every (a in A, b in B), scattered through the
// profiles' column mappings, as a dense (lenA+1) x (lenB+1) row-major float matrix (row/column 0 unused, zero).
struct ProfilePosterior {
    virtual ~ProfilePosterior() {}
    // returns 0 or a negative MLP_E_* code; *dense points at a provider-owned buffer valid until the next call
    virtual int build(const Profile& A, const Profile& B, const float* weights, const float** dense) = 0;
    // optional fused path: MEA traceback string directly (returns 1 = not supported, caller runs the host DP on `dense`)
    virtual int build_and_align(const Profile& A, const Profile& B, const float* weights, std::string& path) { (void)A; (void)B; (void)weights; path.clear(); return 1; }
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

// MEA traceback over a dense profile posterior (ProbabilisticModel::computeAlignment, ProbabilisticModel.cpp:345-421)
std::string mea_path(int l1, int l2, const float* dense);

}  // namespace qptail
