// QuickProbs flavour, host tail: guide-tree driven progressive construction and column-based refinement.
// Restates (flat arrays, no Sequence/MultiSequence objects):
//   ConstructionStage::processTree / alignAlignments      ConstructionStage.cpp:51-127
//   ParallelProbabilisticModel::buildPosterior            ParallelProbabilisticModel.cpp:301-444 (host provider)
//   ProbabilisticModel::computeAlignment                  ProbabilisticModel.cpp:345-421
//   Sequence::AddGaps / getMapping                        Sequence.cpp:67-119
//   MultiSequence::extractSubset / SortByLabel            MultiSequence.cpp:390-460
//   RefinementBase::operator() / checkAcceptance          RefinementBase.cpp:13-116
//   ColumnRefinement::initialise/split/updateColumnScores ColumnRefinement.cpp:56-174
//   det_uniform_int_distribution                          Common/deterministic_random.h:60-85
#include "qp_tail.h"
#include "../../include/mlprobs_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <random>
#include <ctime>

namespace qptail {

namespace {

struct Cell { int32_t col; float val; };

// column mapping of a gapped row: map[k] = 1-based alignment column of the k-th residue (map[0] = 0)
void row_mapping(const std::string& row, std::vector<int>& map) {
    map.clear();
    map.push_back(0);
    for (int c = 0; c < (int)row.size(); ++c)
        if (row[c] != '-') map.push_back(c + 1);
}

class HostProfilePosterior : public ProfilePosterior {
public:
    explicit HostProfilePosterior(const HostCsrView& v) : v_(v) {}
    int build(const Profile& A, const Profile& B, const WeightSpec& ws, const float** out) override {
        const int l1 = A.length(), l2 = B.length();
        std::vector<float>& dense = dense_;
        dense.assign((size_t)(l1 + 1) * (l2 + 1), 0.0f);
        const double total = ws.total(A, B);
        std::vector<std::vector<int>> mapB(B.count());
        for (int j = 0; j < B.count(); ++j) row_mapping(B.rows[j], mapB[j]);
        std::vector<int> mapA;
        const Cell* cells = (const Cell*)v_.cells;
        for (int i = 0; i < A.count(); ++i) {
            const int first = A.ids[i];
            row_mapping(A.rows[i], mapA);
            const double w1 = ws.weight_of(first);
            for (int j = 0; j < B.count(); ++j) {
                const int second = B.ids[j];
                const float w = ws.pair_weight(w1, ws.weight_of(second), total);
                const int64_t slot = (int64_t)first * v_.n + second;
                const int32_t* rp = v_.rp_pool + v_.rp_off[slot];
                const Cell* base = cells + v_.nz_off[slot];
                const int* mb = mapB[j].data();
                const int L = v_.len[first];
                for (int ii = 1; ii <= L; ++ii) {
                    float* drow = dense.data() + (size_t)mapA[ii] * (l2 + 1);
                    for (int k = rp[ii]; k < rp[ii + 1]; ++k) drow[mb[base[k].col]] += w * base[k].val;
                }
            }
        }
        *out = dense.data();
        return 0;
    }
private:
    HostCsrView v_;
    std::vector<float> dense_;
};

// Sequence::AddGaps: consume the row where the path has 'B' or `id`, gap elsewhere
std::string add_gaps(const std::string& row, const std::string& path, char id) {
    std::string out(path.size(), '-');
    const char* src = row.data();
    const char* p = path.data();
    char* dst = &out[0];
    size_t k = 0;
    const size_t n = path.size();
    for (size_t q = 0; q < n; ++q)
        if (p[q] == 'B' || p[q] == id) dst[q] = src[k++];
    return out;
}

int align_profiles(const Profile& A, const Profile& B, const WeightSpec& weights, ProfilePosterior& prov, Profile& out) {
    std::string path;
    int rc = prov.build_and_align(A, B, weights, path);
    if (rc < 0) return rc;
    if (rc > 0) {
        const float* dense = nullptr;
        rc = prov.build(A, B, weights, &dense);
        if (rc < 0) return rc;
        path = mea_path(A.length(), B.length(), dense);
    }
    // both groups arrive sorted by label (leaves, SortByLabel'ed results, index-ordered subsets): merge keeps that order
    const int total = A.count() + B.count();
    out.ids.assign(total, 0);
    out.rows.assign(total, std::string());
    std::vector<int> src(total);                      // >= 0: row of A, < 0: ~row of B
    {
        int i = 0, j = 0, k = 0;
        while (i < A.count() || j < B.count()) {
            if (j >= B.count() || (i < A.count() && A.ids[i] < B.ids[j])) { out.ids[k] = A.ids[i]; src[k++] = i++; }
            else { out.ids[k] = B.ids[j]; src[k++] = ~(j++); }
        }
    }
#pragma omp parallel for schedule(static) if (total >= 64)
    for (int k = 0; k < total; ++k)
        out.rows[k] = src[k] >= 0 ? add_gaps(A.rows[src[k]], path, 'X') : add_gaps(B.rows[~src[k]], path, 'Y');
    return 0;
}

// rows `idx` of P with the columns that are gaps in all of them removed
void extract_subset(const Profile& P, const std::vector<int>& idx, Profile& out) {
    const int len = P.length();
    const int m = (int)idx.size();
    std::vector<char> keep(len, 0);
#pragma omp parallel for schedule(static) if ((long long)m * len >= (1 << 16))
    for (int c0 = 0; c0 < len; c0 += 256) {           // column blocks: each thread ORs its own slice of `keep`
        const int c1 = std::min(len, c0 + 256);
        for (int r : idx) {
            const char* s = P.rows[r].data();
            for (int c = c0; c < c1; ++c) keep[c] |= (char)(s[c] != '-');
        }
    }
    std::vector<int> cols;
    cols.reserve(len);
    for (int c = 0; c < len; ++c) if (keep[c]) cols.push_back(c);
    out.ids.assign(m, 0);
    out.rows.assign(m, std::string());
    const int nc = (int)cols.size();
#pragma omp parallel for schedule(static) if ((long long)m * len >= (1 << 16))
    for (int k = 0; k < m; ++k) {
        const char* src = P.rows[idx[k]].data();
        std::string s((size_t)nc, '-');
        for (int q = 0; q < nc; ++q) s[q] = src[cols[q]];
        out.ids[k] = P.ids[idx[k]];
        out.rows[k] = std::move(s);
    }
}

struct ColumnRefiner {
    std::vector<std::pair<int, float>> scores;   // persists across passes exactly like the reference's member vector
    std::vector<int> gaps;
    std::mt19937 engine;

    void update_scores(const Profile& P) {
        const int n = P.count(), len = P.length();
        scores.resize(len, std::pair<int, float>(0, 0.0f));
        // the reference adds 1.0f per gap, column by column; counting first gives the same float as long as the running
        // value stays an exactly representable integer (below 2^24), otherwise fall back to the literal loop
        gaps.assign(len, 0);
#pragma omp parallel for schedule(static) if ((long long)n * len >= (1 << 16))
        for (int c0 = 0; c0 < len; c0 += 256) {
            const int c1 = std::min(len, c0 + 256);
            for (int i = 0; i < n; ++i) {
                const char* s = P.rows[i].data();
                for (int c = c0; c < c1; ++c) gaps[c] += (s[c] == '-');
            }
        }
        for (int c = 0; c < len; ++c) {
            scores[c].first = c;
            const float before = scores[c].second;
            if (before >= 0 && before + (float)gaps[c] < 16777216.0f && before == (float)(int)before) scores[c].second = before + (float)gaps[c];
            else for (int k = 0; k < gaps[c]; ++k) scores[c].second += 1.0f;
        }
        std::stable_sort(scores.begin(), scores.end(), [n](const std::pair<int, float>& a, const std::pair<int, float>& b) {
            return std::fabs((float)n / 2 - a.second) > std::fabs((float)n / 2 - b.second);
        });
        scores.erase(std::remove_if(scores.begin(), scores.end(), [](const std::pair<int, float>& e) { return e.second == 0; }), scores.end());
    }

    int draw(int lo, int hi_inclusive) {
        typedef unsigned int U;
        const U diff = (U)hi_inclusive - (U)lo + 1;
        if (diff == 0) return (int)engine();
        const U bad = std::numeric_limits<U>::max() / diff;
        for (;;) {
            const U r = (U)engine();
            if (r / diff < bad) return (int)((r % diff) + (U)lo);
        }
    }
};

}  // namespace

ProfilePosterior* make_host_provider(const HostCsrView& v) { return new HostProfilePosterior(v); }

double WeightSpec::total(const Profile& A, const Profile& B) const {
    if (mode == QP_DOUBLE) {
        double t = 0;                                         // finalSelectivity is FLT_MAX: every pair counts
        for (int a : A.ids) { const double w1 = wf[a]; for (int b : B.ids) t += w1 * (double)wf[b]; }
        return t;
    }
    if (mode == CPNP_INT) {
        float t = 0;                                          // `float totalWeights += w1 * w2` with int operands
        for (int a : A.ids) { const int w1 = wi[a]; for (int b : B.ids) t += w1 * wi[b]; }
        return (double)t;
    }
    return 1.0;
}

// glibc's rand() (TYPE_3 additive feedback, r[i] = r[i-3] + r[i-31]) seeded with 1, which is what an unseeded program
// gets; kept private so that the library neither depends on nor disturbs the host application's generator.
struct GlibcRand {
    std::vector<uint32_t> state;
    size_t pos;
    explicit GlibcRand(uint32_t seed = 1) {
        // srandom_r: seed 0 means 1; Schrage's form of 16807 * x mod (2^31 - 1) on 32-bit words, as glibc writes it
        std::vector<int64_t> s(34);
        int32_t word = seed == 0 ? 1 : (int32_t)seed;
        s[0] = word;
        for (int i = 1; i < 31; ++i) {
            const int32_t hi = word / 127773, lo = word % 127773;
            word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            s[i] = word;
        }
        for (int i = 31; i < 34; ++i) s[i] = s[i - 31];
        state.resize(34);
        for (int i = 0; i < 34; ++i) state[i] = (uint32_t)s[i];
        for (int i = 34; i < 344; ++i) state.push_back(state[i - 31] + state[i - 3]);
        pos = state.size();
    }
    int next() {
        state.push_back(state[pos - 31] + state[pos - 3]);
        const uint32_t v = state[pos++];
        if (state.size() > (1u << 16)) {                      // keep the history window small
            state.erase(state.begin(), state.end() - 64);
            pos = state.size();
        }
        return (int)(v >> 1);
    }
};

int run_cpnp_tail(int n, const int32_t* len, const uint8_t* residues, const int32_t* iweights, const int32_t* left, const int32_t* right,
                  ProfilePosterior& prov, int refine_reps, int pid, Profile& out, std::string& err) {
    if (n < 1) { err = "no sequences"; return MLP_E_ARG; }
    std::vector<long long> off(n);
    long long tot = 0;
    for (int i = 0; i < n; ++i) { off[i] = tot; tot += len[i]; }
    if (n == 1) {
        out.ids = {0};
        out.rows = {std::string((const char*)residues, (size_t)len[0])};
        return 0;
    }
    WeightSpec wtd;
    wtd.mode = WeightSpec::CPNP_INT;
    wtd.wi = iweights;
    WeightSpec flat;
    flat.mode = WeightSpec::UNWEIGHTED;
    // MSA::ProcessTree (MSA.cpp:1369-1402): post-order, weighted profile posterior, rows sorted by label after every merge
    const int total = 2 * n - 1;
    std::vector<std::unique_ptr<Profile>> prof(total);
    auto leaf = [&](int v) {
        std::unique_ptr<Profile> p(new Profile());
        p->ids = {v};
        p->rows = {std::string((const char*)residues + off[v], (size_t)len[v])};
        return p;
    };
    for (int v = n; v < total; ++v) {
        const int l = left[v], r = right[v];
        if (l < 0 || r < 0 || l >= v || r >= v) { err = "malformed guide tree"; return MLP_E_ARG; }
        if (l < n) prof[l] = leaf(l);
        if (r < n) prof[r] = leaf(r);
        if (!prof[l] || !prof[r]) { err = "guide tree node used twice"; return MLP_E_ARG; }
        prof[v].reset(new Profile());
        const int rc = align_profiles(*prof[l], *prof[r], wtd, prov, *prof[v]);
        if (rc < 0) { err = "profile posterior failed"; return rc; }
        prof[l].reset();
        prof[r].reset();
    }
    std::unique_ptr<Profile> aln = std::move(prof[total - 1]);

    // MSA::ComputeFinalAlignment (MSA.cpp:1481-1534): the pass count adapts to what the passes report
    int reps = refine_reps;
    if (pid > 3 || n > 150) reps = 0;
    if (n <= 50) reps = 2 * reps;
    GlibcRand rng;
    int ineffectiveness = 0;
    const int cutoff_iter = 100;
    Profile one, two;
    std::vector<int> g1, g2;
    for (int it = 0; it < reps; ++it) {
        // MSA::DoIterativeRefinement (MSA.cpp:1537-1623): random bipartition by position in the CURRENT row order
        int flag;
        g1.clear(); g2.clear();
        for (int i = 0; i < n; ++i) ((rng.next() % 2) ? g1 : g2).push_back(i);
        if (g1.empty() || g2.empty()) flag = 2;
        else {
            extract_subset(*aln, g1, one);
            extract_subset(*aln, g2, two);
            const int l1 = one.length(), l2 = two.length();
            // "accuracy" of the current alignment: posterior mass on its own columns, float sum in column order
            std::vector<long long> offs;
            {
                const int L = aln->length();
                offs.reserve(L);
                int i1 = 0, i2 = 0;
                for (int c = 0; c < L; ++c) {
                    bool f1 = false, f2 = false;
                    for (int k : g1) if (aln->rows[k][c] != '-') { f1 = true; break; }
                    for (int k : g2) if (aln->rows[k][c] != '-') { f2 = true; break; }
                    if (f1) ++i1;
                    if (f2) ++i2;
                    if (f1 && f2) offs.push_back((long long)i1 * (l2 + 1) + i2);
                }
            }
            float before = 0, score = 0;
            std::string path;
            std::vector<float> vals;
            int rc = prov.build_align_score(one, two, flat, offs, path, &score, vals);
            if (rc < 0) { err = "profile posterior failed"; return rc; }
            if (rc > 0) {                                     // provider without the fused path: dense matrix + host DP
                const float* dense = nullptr;
                rc = prov.build(one, two, flat, &dense);
                if (rc < 0) { err = "profile posterior failed"; return rc; }
                for (long long o : offs) before += dense[o];
                path = mea_path(l1, l2, dense, &score);
            } else {
                for (float v : vals) before += v;
            }
            std::unique_ptr<Profile> next(new Profile());
            for (int k = 0; k < one.count(); ++k) { next->ids.push_back(one.ids[k]); next->rows.push_back(add_gaps(one.rows[k], path, 'X')); }
            for (int k = 0; k < two.count(); ++k) { next->ids.push_back(two.ids[k]); next->rows.push_back(add_gaps(two.rows[k], path, 'Y')); }
            aln = std::move(next);                            // no label sort here: the row order drifts, as in the reference
            flag = (before == score) ? 1 : 0;
        }
        if (n > 20) {
            if (n < 200) {
                if (flag > 0) {
                    if (reps < 4 * n) ++reps;
                    if (flag == 1) ++ineffectiveness;
                }
                if (ineffectiveness > 2 * n && it > cutoff_iter) break;
            } else if (n > 200) reps = 10;
        }
    }
    out = std::move(*aln);
    return 0;
}

// MSA::FindSimilar (MSA.cpp:1988-2078): for every sequence x a two-means split of the others by their distance to x; S_x is
// the cluster seeded with the farthest sequence, and always holds x itself
static void find_similar(int n, const float* distances, std::vector<std::vector<int>>& sim) {
    std::vector<float> d(distances, distances + (size_t)n * n);
    for (int i = 0; i < n; ++i) d[(size_t)i * n + i] = 1.0f;
    sim.assign(n, std::vector<int>());
    std::vector<char> in1(n);
    std::vector<int> changes(n);
    for (int i = 0; i < n; ++i) {
        const float* di = d.data() + (size_t)i * n;
        float min_d = 1, max_d = 0;
        int at_min = 0, at_max = 0;
        for (int j = 0; j < n; ++j) {                         // ties: the last index wins
            if (di[j] <= min_d) { at_min = j; min_d = di[j]; }
            if (di[j] >= max_d) { at_max = j; max_d = di[j]; }
        }
        std::fill(in1.begin(), in1.end(), 0);
        // both seeds are inserted into sets: when they coincide the sequence sits in both clusters, and the membership
        // tests below look at the first cluster first
        std::vector<char> in2(n, 0);
        in1[at_max] = 1;
        in2[at_min] = 1;
        for (int j = 0; j < n; ++j)
            if (j != at_min && j != at_max) {
                const float v = d[(size_t)j * n + i];
                if (std::fabs(v - max_d) < std::fabs(v - min_d)) in1[j] = 1; else in2[j] = 1;
            }
        if (!in1[i]) { in2[i] = 0; in1[i] = 1; }
        bool changed = true;
        for (int pass = 0; pass < 100 && changed; ++pass) {
            changed = false;
            std::fill(changes.begin(), changes.end(), 0);
            float m1 = 0, m2 = 0;
            size_t c1 = 0, c2 = 0;
            for (int j = 0; j < n; ++j) if (in1[j]) { m1 += di[j]; ++c1; }
            for (int j = 0; j < n; ++j) if (in2[j]) { m2 += di[j]; ++c2; }
            m1 /= (float)c1;
            m2 /= (float)c2;
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                const float v = d[(size_t)j * n + i];
                if (in1[j]) { if (std::fabs(v - m1) > std::fabs(v - m2)) { changes[j] = 1; changed = true; } }
                else if (std::fabs(v - m2) > std::fabs(v - m1)) { changes[j] = -1; changed = true; }
            }
            if (changed)
                for (int j = 0; j < n; ++j) {
                    if (changes[j] == 1) { in1[j] = 0; in2[j] = 1; }
                    else if (changes[j] == -1) { in2[j] = 0; in1[j] = 1; }
                }
        }
        for (int j = 0; j < n; ++j) if (in1[j]) sim[i].push_back(j);
    }
}

int run_cpnp_np_tail(const HostCsrView& graph_set, const uint8_t* residues, const float* distances, ProfilePosterior& prov,
                     int refine_reps, long long seed, Profile& out, std::string& err) {
    const int n = graph_set.n;
    std::unique_ptr<Profile> aln(new Profile());
    int rc = build_graph_alignment(graph_set, residues, *aln, err);
    if (rc < 0) return rc;
    // MSA::DoRefinement (MSA.cpp:1852-1980): every sequence x in a random order: re-align x against the rest of S_x, then the
    // updated S_x against everything else.  AlignAlignments re-sorts the rows by label, so row k is always sequence k.
    int reps = refine_reps;
    if (n > 150) reps = 0;
    if (reps > 0 && !distances) { err = "distances required for the refinement"; return MLP_E_ARG; }
    std::vector<std::vector<int>> sim;
    if (reps > 0) find_similar(n, distances, sim);
    WeightSpec flat;
    flat.mode = WeightSpec::UNWEIGHTED;
    // AlignAlignments(…, nflag = false) = unweighted profile posterior, MEA path, rows merged by label: align_profiles().
    // The reference also tracks the MEA scores (`nalignscore < oalignscore` extends the pass budget), but oalignscore starts
    // at 0 and a sum of posteriors is never negative, so that branch cannot fire and the scores are not computed here.
    int cnt = 0;
    Profile one, two, self, rest, updated;
    std::vector<int> others, pick(1), remaining;
    while (cnt < reps) {
        GlibcRand rng(seed >= 0 ? (uint32_t)seed : (uint32_t)time(nullptr));     // srand(time(0)) before every sweep
        std::vector<int> pool(n), visit;
        for (int i = 0; i < n; ++i) pool[i] = i;
        while (!pool.empty()) {
            const int at = rng.next() % (int)pool.size();
            visit.push_back(pool[at]);
            pool.erase(pool.begin() + at);
        }
        for (int step = 0; step < n; ++step) {
            const int x = visit[step];
            const std::vector<int>& g1 = sim[x];
            others.clear();
            for (int j = 0, k = 0; j < n; ++j) { if (k < (int)g1.size() && g1[k] == j) ++k; else others.push_back(j); }
            ++cnt;
            if (g1.empty() || others.empty()) continue;
            extract_subset(*aln, g1, one);
            extract_subset(*aln, others, two);
            if (one.count() > 1) {
                const int at = (int)(std::find(g1.begin(), g1.end(), x) - g1.begin());
                pick[0] = at;
                remaining.clear();
                for (int k = 0; k < one.count(); ++k) if (k != at) remaining.push_back(k);
                extract_subset(one, pick, self);
                extract_subset(one, remaining, rest);
                rc = align_profiles(self, rest, flat, prov, updated);
                if (rc < 0) { err = "profile posterior failed"; return rc; }
                ++cnt;
                std::swap(one, updated);
            }
            std::unique_ptr<Profile> next(new Profile());
            rc = align_profiles(one, two, flat, prov, *next);
            if (rc < 0) { err = "profile posterior failed"; return rc; }
            aln = std::move(next);
        }
    }
    out = std::move(*aln);
    return 0;
}

std::string mea_path(int l1, int l2, const float* dense, float* score) {
    const size_t W = (size_t)l2 + 1;
    std::vector<float> two(2 * W, 0.0f);
    float* oldr = two.data();
    float* newr = two.data() + W;
    std::vector<char> tb((size_t)(l1 + 1) * W);
    for (int j = 0; j <= l2; ++j) tb[j] = 'L';
    for (int i = 1; i <= l1; ++i) {
        const float* p = dense + (size_t)i * W;
        char* t = tb.data() + (size_t)i * W;
        newr[0] = 0;
        t[0] = 'U';
        for (int j = 1; j <= l2; ++j) {
            const float x1 = p[j] + oldr[j - 1], x2 = newr[j - 1], x3 = oldr[j];
            if (x1 >= x2) {
                if (x1 >= x3) { newr[j] = x1; t[j] = 'D'; } else { newr[j] = x3; t[j] = 'U'; }
            } else if (x2 >= x3) { newr[j] = x2; t[j] = 'L'; }
            else { newr[j] = x3; t[j] = 'U'; }
        }
        std::swap(oldr, newr);
    }
    if (score) *score = oldr[l2];
    std::string path;
    int r = l1, c = l2;
    while (r != 0 || c != 0) {
        const char ch = tb[(size_t)r * W + c];
        if (ch == 'L') { --c; path.push_back('Y'); }
        else if (ch == 'U') { --r; path.push_back('X'); }
        else { --c; --r; path.push_back('B'); }
    }
    std::reverse(path.begin(), path.end());
    return path;
}

int run_tail(int n, const int32_t* len, const uint8_t* residues, const float* weights, const int32_t* left, const int32_t* right,
             ProfilePosterior& prov, const TailOptions& opt, Profile& out, std::string& err) {
    if (n < 1) { err = "no sequences"; return MLP_E_ARG; }
    WeightSpec ws;
    ws.mode = WeightSpec::QP_DOUBLE;
    ws.wf = weights;
    std::vector<long long> off(n);
    long long tot = 0;
    for (int i = 0; i < n; ++i) { off[i] = tot; tot += len[i]; }
    if (n == 1) {
        out.ids = {0};
        out.rows = {std::string((const char*)residues, (size_t)len[0])};
        return 0;
    }
    // progressive construction: children always have smaller node ids than their parent, so an ascending sweep is a post-order walk
    const int total = 2 * n - 1;
    std::vector<std::unique_ptr<Profile>> prof(total);
    auto leaf = [&](int v) {
        std::unique_ptr<Profile> p(new Profile());
        p->ids = {v};
        p->rows = {std::string((const char*)residues + off[v], (size_t)len[v])};
        return p;
    };
    for (int v = n; v < total; ++v) {
        const int l = left[v], r = right[v];
        if (l < 0 || r < 0 || l >= v || r >= v) { err = "malformed guide tree"; return MLP_E_ARG; }
        if (l < n) prof[l] = leaf(l);
        if (r < n) prof[r] = leaf(r);
        if (!prof[l] || !prof[r]) { err = "guide tree node used twice"; return MLP_E_ARG; }
        prof[v].reset(new Profile());
        const int rc = align_profiles(*prof[l], *prof[r], ws, prov, *prof[v]);
        if (rc < 0) { err = "profile posterior failed"; return rc; }
        prof[l].reset();
        prof[r].reset();
    }
    std::unique_ptr<Profile> aln = std::move(prof[total - 1]);

    // column refinement
    const int iters = opt.ref_iters > 0 ? opt.ref_iters : (opt.ref_iters == -2 ? 0 : (aln->count() > 200 ? 200 : 30));
    ColumnRefiner cr;
    if (opt.ref_seed != 0) cr.engine.seed(opt.ref_seed);
    cr.update_scores(*aln);
    const bool prepared = !cr.scores.empty();
    Profile one, two;
    std::vector<int> g1, g2;
    for (int it = 0; it < iters && prepared; ++it) {
        cr.update_scores(*aln);
        const int hi = (int)cr.scores.size();
        if (hi <= 0) continue;
        const int rnd = cr.draw(0, hi - 1);
        const int col = std::min((size_t)cr.scores[rnd].first, (size_t)aln->length() - 1);
        g1.clear(); g2.clear();
        for (int i = 0; i < aln->count(); ++i) (aln->rows[i][col] == '-' ? g1 : g2).push_back(i);
        if (g1.empty() || g2.empty()) continue;
        extract_subset(*aln, g1, one);
        extract_subset(*aln, g2, two);
        std::unique_ptr<Profile> cand(new Profile());
        const int rc = align_profiles(one, two, ws, prov, *cand);
        if (rc < 0) { err = "profile posterior failed"; return rc; }
        if (aln->length() >= cand->length()) aln = std::move(cand);
    }
    out = std::move(*aln);
    return 0;
}

}  // namespace qptail

extern "C" int mlp_qp_finish_alignment_host(int n, const int32_t* len, const uint8_t* residues, const float* weights,
                                            const int32_t* left, const int32_t* right, const int64_t* rp_off, const int64_t* nz_off,
                                            const int32_t* rp_pool, const void* cells, int ref_iters, uint32_t ref_seed,
                                            char** rows_out, int32_t* aln_len) {
    if (!len || !residues || !rows_out || !aln_len) return MLP_E_ARG;
    if (n > 1 && (!weights || !left || !right || !rp_off || !nz_off || !rp_pool || !cells)) return MLP_E_ARG;
    qptail::HostCsrView v{n, len, rp_off, nz_off, rp_pool, cells};
    std::unique_ptr<qptail::ProfilePosterior> prov(qptail::make_host_provider(v));
    qptail::TailOptions opt;
    opt.ref_iters = ref_iters;
    opt.ref_seed = ref_seed;
    qptail::Profile out;
    std::string err;
    const int rc = qptail::run_tail(n, len, residues, weights, left, right, *prov, opt, out, err);
    if (rc < 0) return rc;
    const int L = out.length();
    char* buf = (char*)std::malloc((size_t)n * (size_t)std::max(L, 1));
    if (!buf) return MLP_E_NOMEM;
    for (int i = 0; i < n; ++i) std::memcpy(buf + (size_t)i * L, out.rows[i].data(), (size_t)L);
    *rows_out = buf;
    *aln_len = L;
    return MLP_OK;
}

extern "C" void mlp_free_host(void* p) { std::free(p); }

extern "C" int mlp_cpnp_finish_alignment_host(int n, const int32_t* len, const uint8_t* residues, const int32_t* iweights,
                                              const int32_t* left, const int32_t* right, const int64_t* rp_off, const int64_t* nz_off,
                                              const int32_t* rp_pool, const void* cells, int refine_reps, int pid,
                                              char** rows_out, int32_t* aln_len, int32_t* order_out) {
    if (!len || !residues || !rows_out || !aln_len) return MLP_E_ARG;
    if (n > 1 && (!iweights || !left || !right || !rp_off || !nz_off || !rp_pool || !cells)) return MLP_E_ARG;
    qptail::HostCsrView v{n, len, rp_off, nz_off, rp_pool, cells};
    std::unique_ptr<qptail::ProfilePosterior> prov(qptail::make_host_provider(v));
    qptail::Profile out;
    std::string err;
    const int rc = qptail::run_cpnp_tail(n, len, residues, iweights, left, right, *prov, refine_reps, pid, out, err);
    if (rc < 0) return rc;
    const int L = out.length();
    char* buf = (char*)std::malloc((size_t)n * (size_t)std::max(L, 1));
    if (!buf) return MLP_E_NOMEM;
    for (int i = 0; i < n; ++i) {
        std::memcpy(buf + (size_t)i * L, out.rows[i].data(), (size_t)L);
        if (order_out) order_out[i] = out.ids[i];
    }
    *rows_out = buf;
    *aln_len = L;
    return MLP_OK;
}

extern "C" int mlp_cpnp_np_finish_alignment_host(int n, const int32_t* len, const uint8_t* residues, const float* distances,
                                                 const int64_t* rp_off, const int64_t* nz_off, const int32_t* rp_pool,
                                                 const void* cells, int refine_reps, int64_t seed, char** rows_out, int32_t* aln_len) {
    if (n < 1 || !len || !residues || !rows_out || !aln_len) return MLP_E_ARG;
    if (n > 1 && (!rp_off || !nz_off || !rp_pool || !cells)) return MLP_E_ARG;
    qptail::HostCsrView v{n, len, rp_off, nz_off, rp_pool, cells};
    std::unique_ptr<qptail::ProfilePosterior> prov(qptail::make_host_provider(v));
    qptail::Profile out;
    std::string err;
    const int rc = qptail::run_cpnp_np_tail(v, residues, distances, *prov, refine_reps, (long long)seed, out, err);
    if (rc < 0) return rc;
    const int L = out.length();
    char* buf = (char*)std::malloc((size_t)n * (size_t)std::max(L, 1));
    if (!buf) return MLP_E_NOMEM;
    for (int i = 0; i < n; ++i) std::memcpy(buf + (size_t)i * L, out.rows[i].data(), (size_t)L);
    *rows_out = buf;
    *aln_len = L;
    return MLP_OK;
}

// test hook: the first `count` values of the private glibc rand() replica after srand(seed)
extern "C" int mlp_debug_glibc_rand_seeded(uint32_t seed, int count, int32_t* out) {
    if (count < 0 || !out) return MLP_E_ARG;
    qptail::GlibcRand g(seed);
    for (int i = 0; i < count; ++i) out[i] = g.next();
    return MLP_OK;
}

// test hook: the first `count` values of the private glibc rand() replica (seed 1)
extern "C" int mlp_debug_glibc_rand(int count, int32_t* out) {
    if (count < 0 || !out) return MLP_E_ARG;
    qptail::GlibcRand g;
    for (int i = 0; i < count; ++i) out[i] = g.next();
    return MLP_OK;
}
