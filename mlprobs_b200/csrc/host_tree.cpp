// Host-side guide tree of the QuickProbs flavour: UPGMA clustering, sequence weights and subtree-size
// "selectivity" distances that ConsistencyStage consumes (ExtendedMSA.cpp:86-108,169-170).
// Restates ClusterTree::build (ClusterTree.cpp:17-124), GuideTree::calculateSeqsWeights (GuideTree.cpp:114-154)
// and GuideTree::calculateSubtreeDistances (GuideTree.cpp:189-221) with flat arrays instead of linked nodes:
// same scan order, strict '<' minimum, first minimum wins, in-place update of the distance matrix.
#include "../../include/mlprobs_b200.h"
#include <algorithm>
#include <vector>

// first_best: the reference's initial minDist (2.0 QuickProbs, 1.1 cpnp); plain_average: cpnp's varianceid == 0 join
// (idist + jdist) / 2 instead of the size-weighted mean (MSAClusterTree.cpp:274-275); iweights: cpnp's integer weights.
static int upgma_core(int n, float* dist, float first_best, bool plain_average, float* weights, int32_t* iweights, float* subtree_dist,
                      int32_t* parent_out, int32_t* left_out, int32_t* right_out) {
    if (n < 2 || !dist || (!weights && !iweights)) return MLP_E_ARG;
    const int total = 2 * n - 1;
    std::vector<int> parent(total, -1), leaves(total, 0), slot_node(n), alive, lch(total, -1), rch(total, -1);
    std::vector<float> branch(total, 0.0f), joins(n);
    alive.reserve(n);
    for (int i = 0; i < n; ++i) { alive.push_back(i); slot_node[i] = i; leaves[i] = 1; }
    // Cached row minima: rowmin[i] / rowarg[i] = smallest d[i][j] over alive j < i and the smallest such j.  The reference scans
    // (i ascending, j ascending, strict '<'), i.e. it takes the lexicographically first minimal pair; the cache reproduces
    // exactly that choice while turning the O(N^3) scan into ~O(N^2).
    std::vector<float> rowmin(n, 3.0f);
    std::vector<int> rowarg(n, -1);
    std::vector<char> is_alive(n, 1);
    auto rescan = [&](int i) {
        float best = 3.0f; int arg = -1;
        const float* row = dist + (size_t)i * n;
        for (int j : alive) { if (j >= i) break; if (row[j] < best) { best = row[j]; arg = j; } }
        rowmin[i] = best; rowarg[i] = arg;
    };
    for (int i : alive) rescan(i);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) if (dist[(size_t)i * n + j] < 0) return MLP_E_ARG;
    for (int node = n; node < total; ++node) {
        float best = first_best;
        int si = -1;
        for (int i : alive) if (rowarg[i] >= 0 && rowmin[i] < best) { best = rowmin[i]; si = i; }
        if (si < 0) return MLP_E_ARG;
        const int sj = rowarg[si];
        const int ni = slot_node[si], nj = slot_node[sj];
        const float half = best * 0.5f;
        parent[ni] = node; parent[nj] = node; branch[ni] = half; branch[nj] = half;
        lch[node] = ni; rch[node] = nj;
        leaves[node] = leaves[ni] + leaves[nj];
        alive.erase(std::lower_bound(alive.begin(), alive.end(), sj));
        is_alive[sj] = 0;
        const unsigned isize = (unsigned)leaves[ni], jsize = (unsigned)leaves[nj];
        for (int idx : alive) {
            const float idist = dist[(size_t)si * n + idx], jdist = dist[(size_t)sj * n + idx];
            joins[idx] = plain_average ? (idist + jdist) / 2 : (idist * isize + jdist * jsize) / (isize + jsize);
        }
        slot_node[si] = node;
        for (int idx : alive) { dist[(size_t)si * n + idx] = joins[idx]; dist[(size_t)idx * n + si] = joins[idx]; }
        // repair the cache
        rescan(si);
        for (int i : alive) {
            if (i <= sj || i == si) continue;
            if (rowarg[i] == sj || rowarg[i] == si) { rescan(i); continue; }
            if (i > si) {
                const float d = dist[(size_t)i * n + si];
                if (d < rowmin[i] || (d == rowmin[i] && si < rowarg[i])) { rowmin[i] = d; rowarg[i] = si; }
            }
        }
    }
    // weights: sum over the path to the root of branch / leaves-below, float accumulation from the leaf upwards
    if (iweights) {
        // MSAGuideTree::getSeqsWeights MSAGuideTree.cpp:274-322: integer weights, (int)(100 w) normalised to INT_MULTIPLY = 1000, at least 1
        int wsum = 0;
        for (int i = 0; i < n; ++i) {
            float w = 0;
            for (int c = i; parent[c] >= 0; c = parent[c]) w += branch[c] / leaves[c];
            iweights[i] = (int)(100 * w);
            wsum += iweights[i];
        }
        if (wsum == 0) { for (int i = 0; i < n; ++i) iweights[i] = 1; wsum = n; }
        for (int i = 0; i < n; ++i) { iweights[i] = (iweights[i] * 1000) / wsum; if (iweights[i] < 1) iweights[i] = 1; }
    }
    if (weights) {
        float wsum = 0.0f;
        for (int i = 0; i < n; ++i) {
            float w = 0;
            for (int c = i; parent[c] >= 0; c = parent[c]) w += branch[c] / leaves[c];
            weights[i] = w;
        }
        for (int i = 0; i < n; ++i) wsum += weights[i];
        if (wsum == 0) { for (int i = 0; i < n; ++i) weights[i] = 1.0f; wsum = (float)n; }
        for (int i = 0; i < n; ++i) weights[i] = weights[i] / wsum;
    }
    if (subtree_dist) {
        // distance(i,j) = number of leaves under the lowest common ancestor = leaves[left child] + leaves[right child].
        // Every internal node is the LCA of exactly (leaves under left) x (leaves under right) pairs, so one pass over the nodes
        // with the leaves laid out in depth-first order writes each of the N(N-1)/2 entries once: O(N^2) instead of an
        // ancestor walk per pair.
        std::vector<int> lo(total, 0), order;
        order.reserve(n);
        {
            std::vector<int> stack;
            stack.push_back(total - 1);
            std::vector<int> pre;                     // preorder, left before right
            while (!stack.empty()) {
                const int v = stack.back(); stack.pop_back();
                if (v < n) { lo[v] = (int)order.size(); order.push_back(v); }
                else { stack.push_back(rch[v]); stack.push_back(lch[v]); pre.push_back(v); }
            }
            for (size_t k = pre.size(); k-- > 0;) { const int v = pre[k]; lo[v] = lo[lch[v]]; }   // children appear after their parent in preorder
        }
        for (int i = 0; i < n; ++i) subtree_dist[(size_t)i * n + i] = 0.0f;
        for (int v = n; v < total; ++v) {
            const int l = lch[v], r = rch[v];
            const float d = (float)(size_t)(leaves[l] + leaves[r]);
            const int* L = order.data() + lo[l]; const int nl = leaves[l];
            const int* R = order.data() + lo[r]; const int nr = leaves[r];
            for (int x = 0; x < nl; ++x) {
                const int a = L[x];
                for (int y = 0; y < nr; ++y) { const int b = R[y]; subtree_dist[(size_t)a * n + b] = d; subtree_dist[(size_t)b * n + a] = d; }
            }
        }
    }
    if (parent_out) for (int v = 0; v < total; ++v) parent_out[v] = parent[v];
    if (left_out) for (int v = 0; v < total; ++v) left_out[v] = lch[v];
    if (right_out) for (int v = 0; v < total; ++v) right_out[v] = rch[v];
    return MLP_OK;
}

extern "C" int mlp_qp_guide_tree_ex(int n, float* dist, float* weights, float* subtree_dist, int32_t* parent_out,
                                    int32_t* left_out, int32_t* right_out) {
    if (!weights) return MLP_E_ARG;
    return upgma_core(n, dist, 2.0f, false, weights, nullptr, subtree_dist, parent_out, left_out, right_out);
}

// cpnp: MSAClusterTree::generateClusterTree(varianceid) MSAClusterTree.cpp:170-296 + MSAGuideTree::getSeqsWeights
extern "C" int mlp_cpnp_guide_tree(int n, float* dist, int variance_id, int32_t* weights_out, int32_t* left_out, int32_t* right_out) {
    if (!weights_out) return MLP_E_ARG;
    return upgma_core(n, dist, 1.1f, variance_id == 0, nullptr, weights_out, nullptr, nullptr, left_out, right_out);
}

extern "C" int mlp_qp_guide_tree(int n, float* dist, float* weights, float* subtree_dist, int32_t* parent_out) {
    return mlp_qp_guide_tree_ex(n, dist, weights, subtree_dist, parent_out, nullptr, nullptr);
}
