// TEST INFRASTRUCTURE (oracle/): harness around the UNMODIFIED reference c_p_np_aln sources.
// It is compiled against the files where they lie under /root/reference/baseMSA/C_P_NP_Aln
// (see oracle/Makefile: MSA.cpp MSAPartProbs.cpp MSAReadMatrix.cpp MSAClusterTree.cpp
// MSAGuideTree.cpp are compiled in place, -fno-access-control lets this file call the
// private members) and only CALLS the reference's own functions for the hot path:
//   ProbabilisticModel::Compute{Forward,Backward,Posterior}Matrix  ProbabilisticModel.h:153-493
//   ::ComputePostProbs                                              MSAPartProbs.cpp:665-727
//   ProbabilisticModel::ComputeAlignment                            ProbabilisticModel.h:804-864
//   SparseMatrix::SparseMatrix                                      SparseMatrix.h:55-98
//   MSA::ModelAdjustmentTest / MSA::DoRelaxation                    MSA.cpp:775-882 / 1172-1281
// The only restated glue is the three-model merge (MSA.cpp:992-1007 / 1699-1714) and the
// distance formula (MSA.cpp:1019), both one-liners.
// Output goes to a dump file (oracle/dumpfmt.h) or, in bench mode, one JSON line on stdout.
#include <string>
#include <vector>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <omp.h>
#include "dumpfmt.h"
#include "MSA.h"
#include "MSAClusterTree.h"

extern VF initDistrib, gapOpen, gapExtend, emitSingle;
extern VVF emitPairs;
extern bool enableVerbose;
extern int numThreads;
extern void init_arguments();
extern VF* ComputePostProbs(int a, int b, string seq1, string seq2);
extern double sub_matrix[26][26];
extern int subst_index[26];

// `c_p_np_aln -p 1` reseeds rand() with time(0) before every refinement sweep (MSA.cpp:1896). Defining time() here takes
// precedence over libc's for the reference objects linked into this executable, so `--fixtime T` makes that program
// reproducible without touching its sources; without the flag the real clock is returned.
static long long g_fixed_time = -1;
extern "C" time_t time(time_t* out) {
    time_t t = g_fixed_time >= 0 ? (time_t)g_fixed_time
                                 : (time_t)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    if (out) *out = t;
    return t;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Opts {
    std::string mode, fasta, out;
    int pid = -1, reps = 2, threads = 1, p1 = 0, dense = 1, g = 0;
};

static void dump_sparse(DumpWriter& w, const std::string& tag, SparseMatrix* m) {
    int L1 = m->GetSeq1Length();
    std::vector<int32_t> rowptr(L1 + 2, 0), cols;
    std::vector<float> vals;
    for (int i = 1; i <= L1; i++) {
        SafeVector<PIF>::iterator p = m->GetRowPtr(i);
        for (int k = 0; k < m->GetRowSize(i); k++) { cols.push_back(p[k].first); vals.push_back(p[k].second); }
        rowptr[i + 1] = (int32_t)cols.size();
    }
    w.i32(tag + ".rowptr", rowptr.data(), {(uint64_t)rowptr.size()});
    w.i32(tag + ".col", cols.data(), {(uint64_t)cols.size()});
    w.f32(tag + ".val", vals.data(), {(uint64_t)vals.size()});
}

int main(int argc, char** argv) {
    Opts o;
    if (argc < 3) { fprintf(stderr, "usage: ref_cpnp dump|bench <fasta> [out.bin] [--pid P] [--reps R] [--threads T] [--p1] [--nodense]\n"); return 2; }
    o.mode = argv[1]; o.fasta = argv[2];
    int ai = 3;
    if (o.mode == "dump" || o.mode == "msa") { o.out = argv[3]; ai = 4; }
    int ir = -1;
    for (; ai < argc; ai++) {
        std::string a = argv[ai];
        if (a == "--ir") { ir = atoi(argv[++ai]); continue; }
        if (a == "--fixtime") { g_fixed_time = atoll(argv[++ai]); continue; }
        if (a == "--pid") o.pid = atoi(argv[++ai]);
        else if (a == "--reps") o.reps = atoi(argv[++ai]);
        else if (a == "--threads") o.threads = atoi(argv[++ai]);
        else if (a == "--p1") o.p1 = 1;
        else if (a == "--G") o.g = 1;
        else if (a == "--nodense") o.dense = 0;
    }
    if (o.mode == "msa") {
        // the reference's complete program flow (MSA::MSA, MSA.cpp:123-187): `c_p_np_aln -p 0|1 [-ir R] -o out fasta`, with the
        // OpenMP team pinned to o.threads (1 = the deterministic summation order of BuildPosterior)
        numThreads = o.threads;
        if (o.g) {
            // `c_p_np_aln -G fasta`: the feature line on stdout.  The bundled binary always takes every core (MSA.cpp:146-151
            // ignores OMP_NUM_THREADS) and its line then changes from run to run; numThreads pins the team here.
            std::vector<const char*> gv = {"c_p_np_aln", "-G", o.fasta.c_str()};
            MSA feature((int)gv.size(), (char**)gv.data());
            return 0;
        }
        std::string irs = std::to_string(ir);
        std::vector<const char*> av = {"c_p_np_aln", "-p", o.p1 ? "1" : "0", "-o", o.out.c_str()};
        if (ir >= 0) { av.push_back("-ir"); av.push_back(irs.c_str()); }
        av.push_back(o.fasta.c_str());
        MSA whole((int)av.size(), (char**)av.data());
        return 0;
    }
    // An MSA object without running its all-in-one constructor (MSA.cpp:123-187); only the
    // members the called methods touch (numPairs, seqsPairs) are initialised.
    MSA* msa = (MSA*)calloc(1, sizeof(MSA));
    init_arguments();          // MSAReadMatrix.cpp:158-210
    msa->ReadParameters();     // MSA.cpp:444
    MultiSequence* sequences = new MultiSequence();
    sequences->LoadMFA(o.fasta, true);
    const int N = sequences->GetNumSequences();
    numThreads = o.threads;
    omp_set_num_threads(o.threads);

    double t0 = now_s();
    int variance_mean = msa->ModelAdjustmentTest(sequences);   // also overrides initDistrib[2]
    double t_viterbi = now_s() - t0;
    int pid = variance_mean % 10, vpid = variance_mean / 10;
    int pid_ref = pid;
    if (o.pid >= 0) pid = o.pid;

    ProbabilisticModel model(initDistrib, gapOpen, gapExtend, emitPairs, emitSingle);

    // pair list as MSA.cpp:912-917
    int numPairs = (N - 1) * N / 2;
    msa->numPairs = numPairs;
    msa->seqsPairs = new MSA::SeqsPair[numPairs];
    { int p = 0; for (int a = 0; a < N; a++) for (int b = a + 1; b < N; b++) { msa->seqsPairs[p].seq1 = a; msa->seqsPairs[p].seq2 = b; p++; } }

    DumpWriter* w = nullptr;
    if (o.mode == "dump") {
        w = new DumpWriter(o.out.c_str());
        w->scalar_i("n", N); w->scalar_i("pid", pid); w->scalar_i("pid_ref", pid_ref); w->scalar_i("vpid", vpid);
        w->scalar_i("p1", o.p1); w->scalar_i("reps", o.reps);
        w->scalar_f("initDistrib2", initDistrib[2]);
        std::vector<int32_t> lens(N); std::string cat;
        for (int i = 0; i < N; i++) { lens[i] = sequences->GetSequence(i)->GetLength(); cat += sequences->GetSequence(i)->GetString(); }
        w->i32("lens", lens.data(), {(uint64_t)N});
        w->u8("residues", (const uint8_t*)cat.data(), {(uint64_t)cat.size()});
        // tables (ProbabilisticModel.h:58-135), letters 'A'..'Z' only
        std::vector<float> m26(26 * 26), i26(26);
        for (int a = 0; a < 26; a++) { i26[a] = model.insProb['A' + a][0]; for (int b = 0; b < 26; b++) m26[a * 26 + b] = model.matchProb['A' + a]['A' + b]; }
        w->f32("hmm.match", m26.data(), {26, 26}); w->f32("hmm.ins", i26.data(), {26});
        w->f32("hmm.init", model.initialDistribution, {5});
        w->f32("hmm.trans", &model.transProb[0][0], {5, 5});
        w->f32("hmm.ltrans", &model.local_transProb[0][0], {3, 3});
        w->f32("hmm.rtrans", model.random_transProb, {2});
        // per-pair Viterbi statistics behind ModelAdjustmentTest (reference ComputeViterbiAlignment, counting glue as MSA.cpp:819-836)
        std::vector<int32_t> vid(numPairs), vlen(numPairs);
        for (int p = 0; p < numPairs; p++) {
            Sequence* s1 = sequences->GetSequence(msa->seqsPairs[p].seq1);
            Sequence* s2 = sequences->GetSequence(msa->seqsPairs[p].seq2);
            pair<SafeVector<char>*, float> al = model.ComputeViterbiAlignment(s1, s2);
            SafeVector<char>::iterator i1 = s1->GetDataPtr(), i2 = s2->GetDataPtr();
            int i = 1, j = 1, same = 0;
            for (SafeVector<char>::iterator it = al.first->begin(); it != al.first->end(); ++it) {
                if (*it == 'B') { if (i1[i] == i2[j]) same++; i++; j++; }
                else if (*it == 'X') i++; else if (*it == 'Y') j++;
            }
            vid[p] = same; vlen[p] = (int32_t)al.first->size();
            delete al.first;
        }
        w->i32("vit.ident", vid.data(), {(uint64_t)numPairs});
        w->i32("vit.len", vlen.data(), {(uint64_t)numPairs});
        w->scalar_i("variance_mean", variance_mean);
        {   // the -G feature line, straight from MSA::Alter_ModelAdjustmentTest (MSA.cpp:646-762)
            std::string g = msa->Alter_ModelAdjustmentTest(sequences, 1.0);
            w->u8("gline", (const uint8_t*)g.data(), {(uint64_t)g.size()});
        }
        w->f64("part.sub_raw", &sub_matrix[0][0], {26, 26});
        w->i32("part.subst_index", subst_index, {26});
    }

    SafeVector<SafeVector<SparseMatrix*> > sparseMatrices(N, SafeVector<SparseMatrix*>(N, NULL));
    VVF distances(N, VF(N, 0));
    std::vector<double> cells_by_pair(numPairs);

    t0 = now_s();
#pragma omp parallel for schedule(dynamic)
    for (int pairIdx = 0; pairIdx < numPairs; pairIdx++) {
        int a = msa->seqsPairs[pairIdx].seq1, b = msa->seqsPairs[pairIdx].seq2;
        Sequence* seq1 = sequences->GetSequence(a);
        Sequence* seq2 = sequences->GetSequence(b);
        const int L1 = seq1->GetLength(), L2 = seq2->GetLength();
        cells_by_pair[pairIdx] = (double)(L1 + 1) * (L2 + 1);
        VF *post5 = NULL, *postP = NULL, *postL = NULL, *posterior = NULL;
        if (pid <= 1 || pid == 5) {   // pid 5 = harness-only: 5-state model alone
            VF* f = model.ComputeForwardMatrix(seq1, seq2);
            VF* bk = model.ComputeBackwardMatrix(seq1, seq2);
            post5 = model.ComputePosteriorMatrix(seq1, seq2, *f, *bk);
            delete f; delete bk;
        }
        if (pid <= 1 || pid >= 3) if (pid != 5) postP = ::ComputePostProbs(a, b, seq1->GetString(), seq2->GetString());
        if (pid <= 2) {
            VF* f = model.ComputeForwardMatrix(seq1, seq2, false);
            VF* bk = model.ComputeBackwardMatrix(seq1, seq2, false);
            postL = model.ComputePosteriorMatrix(seq1, seq2, *f, *bk, false);
            delete f; delete bk;
        }
        if (pid == 2) posterior = new VF(*postL);
        else if (pid == 5) posterior = new VF(*post5);
        else if (pid >= 3) posterior = new VF(*postP);
        else {
            posterior = new VF((L1 + 1) * (L2 + 1));
            for (size_t k = 0; k < posterior->size(); k++) {
                float v1 = (*post5)[k], v2 = (*postP)[k], v3 = (*postL)[k];
                // -p 0: MSA.cpp:1001 ; -p 1: MSA.cpp:1708 (different association of the float sum)
                (*posterior)[k] = o.p1 ? sqrt((v2 * v2 + v3 * v3 + v1 * v1) / 3) : sqrt((v1 * v1 + v2 * v2 + v3 * v3) / 3);
            }
        }
        pair<SafeVector<char>*, float> alignment = model.ComputeAlignment(L1, L2, *posterior);
        float dist;
        if (!o.p1) dist = 1.0f - alignment.second / min(L1, L2);   // MSA.cpp:1019
        else {                                                     // MSA.cpp:1746-1752
            float nmatch = 0;
            for (SafeVector<char>::iterator it = alignment.first->begin(); it != alignment.first->end(); ++it) if (*it == 'B') nmatch += 1;
            dist = alignment.second / nmatch;
        }
        distances[a][b] = distances[b][a] = dist;
        sparseMatrices[a][b] = new SparseMatrix(L1, L2, *posterior);
        if (w) {
#pragma omp critical
            {
                std::string t = "pair." + std::to_string(a) + "." + std::to_string(b);
                if (o.dense) {
                    std::vector<uint64_t> d = {(uint64_t)(L1 + 1), (uint64_t)(L2 + 1)};
                    if (post5) w->f32(t + ".post5", post5->data(), d);
                    if (postP) w->f32(t + ".postP", postP->data(), d);
                    if (postL) w->f32(t + ".postL", postL->data(), d);
                    w->f32(t + ".post", posterior->data(), d);
                }
                w->scalar_f(t + ".mea", alignment.second);
                w->scalar_f(t + ".dist", dist);
                dump_sparse(*w, t + ".s0", sparseMatrices[a][b]);
            }
        }
        delete alignment.first; delete posterior; delete post5; delete postP; delete postL;
    }
    double t_post = now_s() - t0;

    t0 = now_s();
    for (int r = 0; r < o.reps; r++) {
        SafeVector<SafeVector<SparseMatrix*> > nw = msa->DoRelaxation(sequences, sparseMatrices);
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { delete sparseMatrices[i][j]; sparseMatrices[i][j] = nw[i][j]; }
        if (w) for (int a = 0; a < N; a++) for (int b = a + 1; b < N; b++)
            dump_sparse(*w, "pair." + std::to_string(a) + "." + std::to_string(b) + ".s" + std::to_string(r + 1), sparseMatrices[a][b]);
    }
    double t_relax = now_s() - t0;

    if (w) {
        std::vector<float> dm(N * N);
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) dm[i * N + j] = distances[i][j];
        w->f32("distances", dm.data(), {(uint64_t)N, (uint64_t)N});
        delete w;
    }
    double cells = 0; for (double c : cells_by_pair) cells += c;
    int nmodels = (pid <= 1) ? 3 : 1;
    printf("{\"tool\": \"ref_cpnp\", \"n\": %d, \"pairs\": %d, \"pid\": %d, \"pid_ref\": %d, \"threads\": %d, \"cells\": %.0f, \"models\": %d, "
           "\"t_viterbi_s\": %.6f, \"t_posterior_s\": %.6f, \"t_relax_s\": %.6f, \"reps\": %d}\n",
           N, numPairs, pid, pid_ref, o.threads, cells, nmodels, t_viterbi, t_post, t_relax, o.reps);
    return 0;
}
