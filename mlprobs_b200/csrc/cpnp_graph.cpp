// c_p_np_aln -p 1 (non-progressive strategy), host side: the alignment graph that turns the relaxed sparse posteriors
// into a multiple alignment without a guide tree.
// Restates, on flat arrays and bit rows instead of VVI / SafeVector<bool> copies:
//   MSA::ComputeGraph                                 MSA.cpp:1777-1845   (entry list: pairs a<b, rows, cells in row order)
//   AlignGraph::AlignGraph (sort + greedy insertion)  AlignGraph.h:863-1064
//   AlignGraph::Partition / Quick_sort                AlignGraph.h:62-113  (the unstable order of ties is observable)
//   FindCloseNodes                                    AlignGraph.h:181-218
//   CheckAddNewNode / CheckAddColumnEx / CheckAddColumnMrg   AlignGraph.h:383-494 / 503-619 / 628-756
//   Graph2Align / AddtoPath / FindPath / Path2Align   AlignGraph.h:1071-1126 / 765-803 / 810-861
// The reference copies the whole child table for every candidate edit and throws it away when the cycle test fails; the
// cycle tests only read the ancestor/descendant relations, so here the test runs first and the table is edited in place.
// Its relations are not always the exact transitive closure (the tests look at one parent / one child only); they are kept
// with the same update rules, so the same candidates are accepted and rejected.
#include "qp_tail.h"
#include "../../include/mlprobs_b200.h"
#include <algorithm>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace qptail {

namespace {

struct Cell { int32_t col; float val; };
struct Residue { int seq, pos; };
typedef std::vector<uint64_t> Bits;

inline bool test(const Bits& b, int i) { return (b[(size_t)i >> 6] >> (i & 63)) & 1; }
inline void set(Bits& b, int i) { b[(size_t)i >> 6] |= 1ull << (i & 63); }
inline void or_into(Bits& dst, const Bits& src) { for (size_t k = 0; k < dst.size(); ++k) dst[k] |= src[k]; }
inline bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }
inline void drop(std::vector<int>& v, int x) { v.erase(std::remove(v.begin(), v.end(), x), v.end()); }
template <class F> inline void for_each_bit(const Bits& b, F f) {
    for (size_t k = 0; k < b.size(); ++k)
        for (uint64_t w = b[k]; w; w &= w - 1) f((int)(k * 64 + __builtin_ctzll(w)));
}
// remove position `pos` from the row, everything above it moves down by one
inline void delete_bit(Bits& b, int pos) {
    const size_t w = (size_t)pos >> 6;
    const uint64_t below = (1ull << (pos & 63)) - 1;
    const size_t n = b.size();
    b[w] = (b[w] & below) | ((b[w] >> 1) & ~below) | (w + 1 < n ? (b[w + 1] & 1) << 63 : 0);
    for (size_t k = w + 1; k < n; ++k) b[k] = (b[k] >> 1) | (k + 1 < n ? (b[k + 1] & 1) << 63 : 0);
}

// AlignGraph::Partition: hole-moving partition around the first element of [low, high]; returns the pivot's final position
long long reference_partition(float* key, int* idx, long long low, long long high) {
    const float pivot = key[low];
    const int pivot_idx = idx[low];
    while (high > low) {
        while (pivot <= key[high] && high > low) --high;
        key[low] = key[high]; idx[low] = idx[high];
        while (pivot >= key[low] && high > low) ++low;
        key[high] = key[low]; idx[high] = idx[low];
    }
    key[low] = pivot; idx[low] = pivot_idx;
    return low;
}

// AlignGraph::Quick_sort.  The two sub-ranges of a partition are disjoint, so neither the order in which they are finished
// nor the thread that finishes them changes the result: an explicit stack replaces the recursion, and large sub-ranges
// become OpenMP tasks.
void reference_sort_serial(float* key, int* idx, long long lo, long long hi) {
    std::vector<std::pair<long long, long long>> todo;
    todo.emplace_back(lo, hi);
    while (!todo.empty()) {
        const long long low = todo.back().first, high = todo.back().second;
        todo.pop_back();
        if (low >= high) continue;
        const long long p = reference_partition(key, idx, low, high);
        todo.emplace_back(low, p - 1);
        todo.emplace_back(p + 1, high);
    }
}

void reference_sort_tasks(float* key, int* idx, long long lo, long long hi) {
    const long long grain = 1 << 16;
    while (hi - lo > grain) {
        const long long p = reference_partition(key, idx, lo, hi);
#pragma omp task default(none) firstprivate(key, idx, lo, p)
        reference_sort_tasks(key, idx, lo, p - 1);
        lo = p + 1;
    }
    reference_sort_serial(key, idx, lo, hi);
}

void reference_sort(std::vector<float>& key, std::vector<int>& idx) {
    if (key.empty()) return;
    float* k = key.data();
    int* i = idx.data();
    const long long last = (long long)key.size() - 1;
#pragma omp parallel
#pragma omp single nowait
    reference_sort_tasks(k, i, 0, last);
}

class Graph {
public:
    Graph(int n, const int32_t* len) : n_(n), len_(len) {
        maxlen_ = 0;
        for (int i = 0; i < n; ++i) maxlen_ = std::max(maxlen_, (int)len[i]);
        node_of_.assign(n, std::vector<int>(maxlen_, -1));
        words_ = 4;
    }

    // one residue pair, strongest first (the constructor's main loop, AlignGraph.h:928-1041)
    void offer(Residue x, Residue y) {
        int cx = node_of_[x.seq][x.pos], cy = node_of_[y.seq][y.pos];
        const bool fx = cx != -1, fy = cy != -1;
        if (!fx && !fy) add_node(x, y);
        else if (fx != fy) {
            if (fy) { std::swap(x, y); std::swap(cx, cy); }
            if (!test(members_[cx], y.seq)) extend_column(y, cx);           // unless cx already holds a residue of y's sequence
        } else if (cx != cy) {
            if (!test(members_[cx], y.seq) && !test(members_[cy], x.seq)) {
                if (cx > cy) std::swap(cx, cy);
                merge_columns(cx, cy);
            }
        }
        reserve_bits();
    }

    // Graph2Align + Path2Align: rows in input order, 0-based columns
    void alignment(const uint8_t* residues, const std::vector<long long>& off, Profile& out) const {
        const int nodes = (int)child_.size();
        std::vector<char> has_parent(nodes, 0), marked(nodes, 0);
        for (const std::vector<int>& c : child_) for (int v : c) has_parent[v] = 1;
        std::vector<int> path;
        path.reserve(nodes);
        for (int r = 0; r < nodes; ++r) {
            if (has_parent[r]) continue;
            path.insert(path.begin(), r);                    // AddtoPath(Path, -1, root): every further root goes to the front
            walk(r, marked, path);
        }
        std::vector<int> where(nodes, 0);
        for (int i = 0; i < (int)path.size(); ++i) where[path[i]] = i;
        // residues that never joined a column: one column each, right after the column of the nearest aligned residue to
        // their left in the same sequence (or before everything)
        std::vector<std::vector<Residue>> after(path.size());
        std::vector<Residue> front;
        for (int s = 0; s < n_; ++s)
            for (int p = 0; p < len_[s]; ++p) {
                if (node_of_[s][p] != -1) continue;
                int q = p - 1;
                while (q >= 0 && node_of_[s][q] == -1) --q;
                if (q >= 0) after[where[node_of_[s][q]]].push_back(Residue{s, p});
                else front.push_back(Residue{s, p});
            }
        std::vector<std::vector<Residue>> column(nodes);    // members in sequence-major order
        for (int s = 0; s < n_; ++s)
            for (int p = 0; p < len_[s]; ++p)
                if (node_of_[s][p] != -1) column[node_of_[s][p]].push_back(Residue{s, p});
        size_t total = front.size() + path.size();
        for (const std::vector<Residue>& a : after) total += a.size();
        out.ids.resize(n_);
        out.rows.assign(n_, std::string(total, '-'));
        for (int s = 0; s < n_; ++s) out.ids[s] = s;
        size_t c = 0;
        auto put = [&](const Residue& r) { out.rows[r.seq][c] = (char)residues[off[r.seq] + r.pos]; };
        for (const Residue& r : front) { put(r); ++c; }
        for (size_t i = 0; i < path.size(); ++i) {
            for (const Residue& r : column[path[i]]) put(r);
            ++c;
            for (const Residue& r : after[i]) { put(r); ++c; }
        }
    }

private:
    // nearest aligned residues of x's sequence to the left / right of x: their nodes, or -1 (FindCloseNodes)
    void neighbours(const Residue& x, int& parent, int& child) const {
        const std::vector<int>& row = node_of_[x.seq];
        parent = child = -1;
        for (int i = x.pos - 1; i >= 0; --i) if (row[i] != -1) { parent = row[i]; break; }
        for (int i = x.pos + 1; i < (int)row.size(); ++i)
            if (row[i] != -1) { if (i != 10000) child = row[i]; break; }      // the reference's "infinity" is the literal 10000
    }

    void add_node(const Residue& x, const Residue& y) {
        int px, chx, py, chy;
        neighbours(x, px, chx);
        neighbours(y, py, chy);
        if (px != -1 && chy != -1 && (test(desc_[chy], px) || px == chy)) return;
        if (py != -1 && chx != -1 && (test(desc_[chx], py) || py == chx)) return;
        std::vector<int> parents, children;
        if (px != -1) parents.push_back(px);
        if (py != -1 && py != px) parents.push_back(py);
        if (chx != -1) children.push_back(chx);
        if (chy != -1 && chy != chx) children.push_back(chy);
        const int g = (int)child_.size();
        child_.push_back(children);
        for (int p : parents) child_[p].push_back(g);
        // edges made redundant by the new node
        if (px != -1 && py != -1) {
            if (test(desc_[px], py)) drop(child_[px], g);
            if (test(desc_[py], px)) drop(child_[py], g);
        }
        if (chx != -1 && chy != -1) {
            if (test(desc_[chx], chy)) drop(child_[g], chy);
            if (test(desc_[chy], chx)) drop(child_[g], chx);
        }
        for (int p : parents) for (int c : children) drop(child_[p], c);
        node_of_[x.seq][x.pos] = g;
        node_of_[y.seq][y.pos] = g;
        members_.push_back(Bits(((size_t)n_ + 63) / 64, 0));
        set(members_[g], x.seq);
        set(members_[g], y.seq);
        Bits a(words_, 0), d(words_, 0);
        if (!parents.empty()) a = anc_[parents[0]];
        if (parents.size() == 2) or_into(a, anc_[parents[1]]);
        for (int p : parents) set(a, p);
        if (!children.empty()) d = desc_[children[0]];
        if (children.size() == 2) or_into(d, desc_[children[1]]);
        for (int c : children) set(d, c);
        anc_.push_back(a);
        desc_.push_back(d);
        close_over(g, a, d);
    }

    void extend_column(const Residue& y, int cx) {
        int parent, child;
        neighbours(y, parent, child);
        bool ok = true;
        if (child != -1) ok = !test(desc_[child], cx) && child != cx;
        if (parent != -1) ok = ok && !test(desc_[cx], parent) && parent != cx;
        if (!ok) return;
        const bool parent_had = parent != -1 && contains(child_[parent], cx);
        const bool cx_had = child != -1 && contains(child_[cx], child);
        if (parent != -1 && !parent_had) child_[parent].push_back(cx);
        if (child != -1 && !cx_had) child_[cx].push_back(child);
        if (parent != -1 && test(desc_[parent], cx) && !parent_had) drop(child_[parent], cx);
        if (child != -1 && test(desc_[cx], child) && !cx_had) drop(child_[cx], child);
        if (parent != -1 && child != -1) drop(child_[parent], child);
        node_of_[y.seq][y.pos] = cx;
        set(members_[cx], y.seq);
        if (parent != -1) { or_into(anc_[cx], anc_[parent]); set(anc_[cx], parent); }
        if (child != -1) { or_into(desc_[cx], desc_[child]); set(desc_[cx], child); }
        const Bits a = anc_[cx], d = desc_[cx];
        close_over(cx, a, d);
    }

    void merge_columns(int cx, int cy) {
        if (test(desc_[cx], cy) || test(desc_[cy], cx)) return;
        const int nodes = (int)child_.size();
        auto renum = [cx, cy](int v) { return v < cy ? v : (v == cy ? cx : v - 1); };
        std::vector<std::vector<int>> next;
        next.reserve(nodes - 1);
        for (int j = 0; j < nodes; ++j) {
            if (j == cy) continue;
            std::vector<int> list;
            if (j == cx) {
                list = child_[cx];
                for (int v : child_[cy]) if (!contains(list, v)) list.push_back(v);
                for (int& v : list) v = renum(v);
            } else {
                bool seen = false;
                for (int v : child_[j]) {
                    if (v == cx || v == cy) { if (!seen) { list.push_back(cx); seen = true; } }
                    else list.push_back(v < cy ? v : v - 1);
                }
            }
            next.push_back(std::move(list));
        }
        // edges made redundant by the merge; all membership tests look at the table and the relations BEFORE the merge
        const Bits &ax = anc_[cx], &ay = anc_[cy], &dx = desc_[cx], &dy = desc_[cy];
        for_each_bit(ax, [&](int a) { for (int v : child_[a]) if (test(dy, v)) drop(next[renum(a)], renum(v)); });
        for_each_bit(ax, [&](int a) { if (contains(child_[a], cy) && !contains(child_[a], cx)) drop(next[renum(a)], cx); });
        for_each_bit(ay, [&](int a) { for (int v : child_[a]) if (test(dx, v)) drop(next[renum(a)], renum(v)); });
        for_each_bit(ay, [&](int a) { if (contains(child_[a], cx) && !contains(child_[a], cy)) drop(next[renum(a)], cx); });
        for (int p = 0; p < nodes; ++p) {
            const bool of_x = contains(child_[p], cx), of_y = contains(child_[p], cy);
            if (of_x && test(ay, p) && !of_y) drop(next[renum(p)], cx);
            if (of_y && test(ax, p) && !of_x) drop(next[renum(p)], cx);
        }
        for (int c : child_[cx]) if (test(dy, c) && !contains(child_[cy], c)) drop(next[cx], renum(c));
        for (int c : child_[cy]) if (test(dx, c) && !contains(child_[cx], c)) drop(next[cx], renum(c));
        for (std::vector<int>& row : node_of_) for (int& v : row) v = renum(v);
        child_.swap(next);
        or_into(members_[cx], members_[cy]);
        members_.erase(members_.begin() + cy);
        Bits a = anc_[cx], d = desc_[cx];
        or_into(a, anc_[cy]);
        or_into(d, desc_[cy]);
        anc_.erase(anc_.begin() + cy);
        desc_.erase(desc_.begin() + cy);
        anc_[cx] = a;
        desc_[cx] = d;
        for (Bits& r : anc_) delete_bit(r, cy);
        for (Bits& r : desc_) delete_bit(r, cy);
        const Bits a2 = anc_[cx], d2 = desc_[cx];
        close_over(cx, a2, d2);
    }

    // every descendant in `d` gains node v and all of `a` as ancestors, every ancestor in `a` gains v and (when there is at
    // least one descendant) all of `d` as descendants
    void close_over(int v, const Bits& a, const Bits& d) {
        for_each_bit(d, [&](int k) { set(anc_[k], v); or_into(anc_[k], a); });
        for_each_bit(a, [&](int k) { or_into(desc_[k], d); set(desc_[k], v); });
    }

    void reserve_bits() {
        if (child_.size() + 16 <= words_ * 64) return;
        words_ += 4;
        for (Bits& r : anc_) r.resize(words_, 0);
        for (Bits& r : desc_) r.resize(words_, 0);
    }

    // FindPath: depth first, every newly reached child is put right behind its parent
    void walk(int from, std::vector<char>& marked, std::vector<int>& path) const {
        std::vector<std::pair<int, size_t>> stack;
        stack.emplace_back(from, 0);
        while (!stack.empty()) {
            const int node = stack.back().first;
            const size_t k = stack.back().second;
            if (k >= child_[node].size()) { stack.pop_back(); continue; }
            ++stack.back().second;
            const int c = child_[node][k];
            if (marked[c]) continue;
            marked[c] = 1;
            path.insert(std::find(path.begin(), path.end(), node) + 1, c);
            stack.emplace_back(c, 0);
        }
    }

    int n_;
    const int32_t* len_;
    int maxlen_;
    size_t words_;
    std::vector<std::vector<int>> child_;          // G
    std::vector<std::vector<int>> node_of_;        // IsPresent
    std::vector<Bits> anc_, desc_;                 // Ancs, Descs
    std::vector<Bits> members_;                    // per node: which sequences have a residue in it (the reference scans IsPresent rows)
};

}  // namespace

int build_graph_alignment(const HostCsrView& v, const uint8_t* residues, Profile& out, std::string& err) {
    const int n = v.n;
    if (n < 1) { err = "no sequences"; return MLP_E_ARG; }
    std::vector<long long> off(n);
    long long tot = 0;
    for (int i = 0; i < n; ++i) { off[i] = tot; tot += v.len[i]; }
    // MSA::ComputeGraph: all cells of the a<b matrices, pair-major, row-major
    std::vector<float> key;
    std::vector<Residue> left, right;
    const Cell* cells = (const Cell*)v.cells;
    {
        size_t total = 0;
        for (int a = 0; a < n; ++a)
            for (int b = a + 1; b < n; ++b) {
                const int32_t* rp = v.rp_pool + v.rp_off[(int64_t)a * n + b];
                total += (size_t)std::max(0, rp[v.len[a] + 1] - rp[1]);
            }
        key.reserve(total); left.reserve(total); right.reserve(total);
    }
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) {
            const int64_t slot = (int64_t)a * n + b;
            const int32_t* rp = v.rp_pool + v.rp_off[slot];
            const Cell* base = cells + v.nz_off[slot];
            for (int i = 1; i <= v.len[a]; ++i)
                for (int k = rp[i]; k < rp[i + 1]; ++k) {
                    if (base[k].col < 1 || base[k].col > v.len[b]) { err = "sparse cell outside its matrix"; return MLP_E_ARG; }
                    key.push_back(base[k].val);
                    left.push_back(Residue{a, i - 1});
                    right.push_back(Residue{b, base[k].col - 1});
                }
        }
    if (key.size() > 0x7fffffffull) { err = "too many sparse cells for the alignment graph"; return MLP_E_UNSUPPORTED; }
    std::vector<int> order(key.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
    const bool times = getenv("MLP_GRAPH_TIMES") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    reference_sort(key, order);
    const double t1 = now();
    Graph g(n, v.len);
    for (size_t k = order.size(); k-- > 0;) g.offer(left[order[k]], right[order[k]]);
    const double t2 = now();
    g.alignment(residues, off, out);
    if (times) fprintf(stderr, "graph: %zu cells, sort %.2f s, insert %.2f s, columns %.2f s\n", order.size(), t1 - t0, t2 - t1, now() - t2);
    return 0;
}

}  // namespace qptail

// test hook: the permutation the graph's sort produces (tasks = 1: the OpenMP task version used in production, 0: one thread)
extern "C" int mlp_debug_reference_sort(int64_t n, const float* keys, int32_t* idx_out, int tasks) {
    if (n < 0 || !keys || !idx_out) return MLP_E_ARG;
    std::vector<float> key(keys, keys + n);
    std::vector<int> idx((size_t)n);
    for (int64_t k = 0; k < n; ++k) idx[(size_t)k] = (int)k;
    if (tasks) qptail::reference_sort(key, idx);
    else if (n > 0) qptail::reference_sort_serial(key.data(), idx.data(), 0, n - 1);
    for (int64_t k = 0; k < n; ++k) idx_out[k] = idx[(size_t)k];
    return MLP_OK;
}
