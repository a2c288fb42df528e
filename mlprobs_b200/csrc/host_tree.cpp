// Host-side guide tree of the QuickProbs flavour: UPGMA clustering, sequence weights and subtree-size
// "selectivity" distances that ConsistencyStage consumes (ExtendedMSA.cpp:86-108,169-170).
// Restates ClusterTree::build (ClusterTree.cpp:17-124), GuideTree::calculateSeqsWeights (GuideTree.cpp:114-154)
// and GuideTree::calculateSubtreeDistances (GuideTree.cpp:189-221) with flat arrays instead of linked nodes:
// same scan order, strict '<' minimum, first minimum wins, in-place update of the distance matrix.
#include "../../include/mlprobs_b200.h"
#include <vector>

extern "C" int mlp_qp_guide_tree(int n, float* dist, float* weights, float* subtree_dist, int32_t* parent_out) {
    if (n < 2 || !dist || !weights) return MLP_E_ARG;
    const int total = 2 * n - 1;
    std::vector<int> parent(total, -1), leaves(total, 0), slot_node(n), alive;
    std::vector<float> branch(total, 0.0f), joins(n);
    alive.reserve(n);
    for (int i = 0; i < n; ++i) { alive.push_back(i); slot_node[i] = i; leaves[i] = 1; }
    for (int node = n; node < total; ++node) {
        float best = 2.0f;
        int bi = -1, bj = -1;
        for (size_t x = 0; x < alive.size(); ++x) {
            const int mi = alive[x];
            const float* row = dist + (size_t)mi * n;
            for (size_t y = 0; y < x; ++y) {          // alive is kept ascending, so alive[y] < mi
                const float d = row[alive[y]];
                if (d < 0) return MLP_E_ARG;
                if (d < best) { best = d; bi = (int)x; bj = (int)y; }
            }
        }
        if (bi < 0) return MLP_E_ARG;
        const int si = alive[bi], sj = alive[bj];
        const int ni = slot_node[si], nj = slot_node[sj];
        const float half = best * 0.5f;
        parent[ni] = node; parent[nj] = node; branch[ni] = half; branch[nj] = half;
        leaves[node] = leaves[ni] + leaves[nj];
        alive.erase(alive.begin() + bj);
        const unsigned isize = (unsigned)leaves[ni], jsize = (unsigned)leaves[nj];
        for (int idx : alive) {
            const float idist = dist[(size_t)si * n + idx], jdist = dist[(size_t)sj * n + idx];
            joins[idx] = (idist * isize + jdist * jsize) / (isize + jsize);
        }
        slot_node[si] = node;
        for (int idx : alive) { dist[(size_t)si * n + idx] = joins[idx]; dist[(size_t)idx * n + si] = joins[idx]; }
    }
    // weights: sum over the path to the root of branch / leaves-below, float accumulation from the leaf upwards
    float wsum = 0.0f;
    for (int i = 0; i < n; ++i) {
        float w = 0;
        for (int c = i; parent[c] >= 0; c = parent[c]) w += branch[c] / leaves[c];
        weights[i] = w;
    }
    for (int i = 0; i < n; ++i) wsum += weights[i];
    if (wsum == 0) { for (int i = 0; i < n; ++i) weights[i] = 1.0f; wsum = (float)n; }
    for (int i = 0; i < n; ++i) weights[i] = weights[i] / wsum;
    if (subtree_dist) {
        // distance(i,j) = number of leaves under the lowest common ancestor = leaves[child_i] + leaves[child_j]
        std::vector<int> depth(total, 0);
        for (int v = total - 2; v >= 0; --v) depth[v] = depth[parent[v]] + 1;   // parents have larger indices
        for (int i = 0; i < n; ++i) {
            subtree_dist[(size_t)i * n + i] = 0.0f;
            for (int j = i + 1; j < n; ++j) {
                int a = i, b = j;
                while (depth[a] > depth[b]) a = parent[a];
                while (depth[b] > depth[a]) b = parent[b];
                while (parent[a] != parent[b]) { a = parent[a]; b = parent[b]; }
                const float d = (float)(size_t)(leaves[a] + leaves[b]);
                subtree_dist[(size_t)i * n + j] = d; subtree_dist[(size_t)j * n + i] = d;
            }
        }
    }
    if (parent_out) for (int v = 0; v < total; ++v) parent_out[v] = parent[v];
    return MLP_OK;
}
