// TEST INFRASTRUCTURE (oracle/): harness around the UNMODIFIED reference QuickProbs CPU sources,
// compiled where they lie under /root/reference/realign/QuickProbs/src (oracle/Makefile).
// It drives the reference's own public stage objects exactly as ExtendedMSA::doAlign does
// (ExtendedMSA.cpp:66-150): PosteriorStage -> ClusterTree -> weights / subtree distances ->
// ConsistencyStage, and dumps what each stage produced. Nothing of the hot path is restated
// here; -fno-access-control only lets the dumps read protected tables.
#include <string>
#include <vector>
#include <iostream>
#include <fstream>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <omp.h>
#include "dumpfmt.h"
#include "Alignment/Alignment.h"
#include "Alignment/Multiple/PosteriorStage.h"
#include "Alignment/Multiple/ConsistencyStage.h"
#include "Alignment/Multiple/ConstructionStage.h"
#include "Alignment/Multiple/ColumnRefinement.h"
#include "Alignment/Multiple/ClusterTree.h"
#include "Alignment/Multiple/PartitionFunction.h"
#include "Alignment/Multiple/ExpPartitionFunctionParams.h"
#include "Alignment/Multiple/ParallelProbabilisticModel.h"
#include "Alignment/DataStructures/ContiguousMultiSequence.h"

using namespace quickprobs;

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void dump_sparse(DumpWriter& w, const std::string& tag, const SparseMatrixType* m) {
    int L1 = m->getSeq1Length();
    std::vector<int32_t> rowptr(L1 + 2, 0);
    std::vector<uint16_t> cols, codes;
    std::vector<float> vals;
    for (int i = 1; i <= L1; i++) {
        auto p = m->getRowPtr(i);
        for (int k = 0; k < m->getRowSize(i); k++) {
            cols.push_back((uint16_t)p[k].getColumn());
            codes.push_back(p[k].second);
            vals.push_back(p[k].getValue());
        }
        rowptr[i + 1] = (int32_t)cols.size();
    }
    w.i32(tag + ".rowptr", rowptr.data(), {(uint64_t)rowptr.size()});
    w.u16(tag + ".col", cols.data(), {(uint64_t)cols.size()});
    w.u16(tag + ".code", codes.data(), {(uint64_t)codes.size()});
    w.f32(tag + ".val", vals.data(), {(uint64_t)vals.size()});
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_qp dump|bench|msa <fasta> [out.bin|out.fasta] [--threads T] [--nodense] [--iters K] [--ref-count R]\n"); return 2; }
    std::string mode = argv[1], fasta = argv[2], out;
    int ai = 3, threads = 1, dense = 1, iters = -1, refcount = -1;
    if (mode == "dump" || mode == "msa") { out = argv[3]; ai = 4; }
    for (; ai < argc; ai++) {
        std::string a = argv[ai];
        if (a == "--threads") threads = atoi(argv[++ai]);
        else if (a == "--nodense") dense = 0;
        else if (a == "--iters") iters = atoi(argv[++ai]);
        else if (a == "--ref-count") refcount = atoi(argv[++ai]);
    }
    // same configuration path as Console/main.cpp:30-46
    auto config = std::shared_ptr<Configuration>(new Configuration());
    std::string tstr = std::to_string(threads);
    const char* av[] = {"quickprobs", fasta.c_str(), "-t", tstr.c_str()};
    if (!config->parse(4, (char**)av)) { fprintf(stderr, "config parse failed\n"); return 3; }
    config->optimisation.useDoublePartition = true;
    if (config->hardware.numThreads <= 0) config->hardware.numThreads = omp_get_num_procs();
    if (iters >= 0) config->algorithm.consistency.itertions = iters;
    if (refcount >= 0) config->algorithm.refinement.iterations = refcount;

    MultiSequence* sequences = new MultiSequence();
    sequences->LoadMFA(fasta, true);
    const int N = sequences->count();
    ContiguousMultiSequence cms(*sequences);
    ISequenceSet* set = &cms;

    Array<float> distances(N);
    Array<SparseMatrixType*> sparse(N);
    PosteriorStage ps(config);
    double t0 = now_s();
    ps(*set, distances, sparse);
    double t_post = now_s() - t0;

    DumpWriter* w = nullptr;
    double cells = 0;
    for (int a = 0; a < N; a++) for (int b = a + 1; b < N; b++)
        cells += (double)(set->GetSequence(a)->GetLength() + 1) * (set->GetSequence(b)->GetLength() + 1);
    if (mode == "dump") {
        w = new DumpWriter(out.c_str());
        w->scalar_i("n", N);
        std::vector<int32_t> lens(N); std::string cat;
        for (int i = 0; i < N; i++) {
            Sequence* s = set->GetSequence(i);
            lens[i] = s->GetLength();
            cat.append(s->getData() + 1, s->GetLength());
        }
        w->i32("lens", lens.data(), {(uint64_t)N});
        w->u8("residues", (const uint8_t*)cat.data(), {(uint64_t)cat.size()});
        auto model = ps.getModel();
        std::vector<float> m26(26 * 26), i26(26);
        for (int a = 0; a < 26; a++) { i26[a] = model->insProb['A' + a][0]; for (int b = 0; b < 26; b++) m26[a * 26 + b] = model->matchProb['A' + a]['A' + b]; }
        w->f32("hmm.match", m26.data(), {26, 26}); w->f32("hmm.ins", i26.data(), {26});
        w->f32("hmm.init", model->initialDistribution, {5});
        w->f32("hmm.trans", &model->transProb[0][0], {5, 5});
        auto& raw = dynamic_cast<ExpPartitionFunctionParams<double>&>(*ps.function->params).raw;
        w->f64("part.sub", raw.subMatrix, {26, 26});
        double g[4] = {raw.termGapOpen, raw.termGapExtend, raw.gapOpen, raw.gapExt};
        w->f64("part.gaps", g, {4});
        BufferSet buf((cms.maxLength + 1) * (cms.maxLength + 1));
        for (int a = 0; a < N; a++) for (int b = a + 1; b < N; b++) {
            std::string t = "pair." + std::to_string(a) + "." + std::to_string(b);
            if (dense) {
                Sequence* s1 = set->GetSequence(a); Sequence* s2 = set->GetSequence(b);
                float d = 0;
                ps.computePairwise(*s1, *s2, buf, d);
                w->f32(t + ".post", buf.f0(), {(uint64_t)(s1->GetLength() + 1), (uint64_t)(s2->GetLength() + 1)});
                w->f32(t + ".postP", buf.f2(), {(uint64_t)(s1->GetLength() + 1), (uint64_t)(s2->GetLength() + 1)});
                w->f32(t + ".post5", buf.f1(), {(uint64_t)(s1->GetLength() + 1), (uint64_t)(s2->GetLength() + 1)});
            }
            w->scalar_f(t + ".dist", distances[a][b]);
            dump_sparse(*w, t + ".s0", sparse[a][b]);
            dump_sparse(*w, t + ".t0", sparse[b][a]);
        }
        w->f32("distances", distances.getData().data(), {(uint64_t)N, (uint64_t)N});
    }

    // ExtendedMSA.cpp:86-108,169-170
    double t1 = now_s();
    ClusterTree tree(distances);
    tree();
    auto weights = tree.getWeights();
    Array<float> cd = tree.calculateSubtreeDistances();
    for (float& x : weights) x = std::max(x, config->algorithm.consistency.saturation);
    double t_tree = now_s() - t1;
    if (w) {
        w->f32("weights", weights.data(), {(uint64_t)N});
        w->f32("seldist", cd.getData().data(), {(uint64_t)N, (uint64_t)N});
        w->f32("distances_after_tree", distances.getData().data(), {(uint64_t)N, (uint64_t)N});
    }
    ConsistencyStage cs(config);
    t1 = now_s();
    cs(weights.data(), *set, cd, sparse);
    double t_cons = now_s() - t1;
    if (w) {
        w->scalar_i("cons.iterations", cs.iterations);
        w->scalar_f("cons.selfweight", cs.selfweight);
        for (int a = 0; a < N; a++) for (int b = a + 1; b < N; b++) {
            std::string t = "pair." + std::to_string(a) + "." + std::to_string(b);
            dump_sparse(*w, t + ".sF", sparse[a][b]);
            dump_sparse(*w, t + ".tF", sparse[b][a]);
        }
        delete w;
    }
    if (mode == "msa") {
        // ExtendedMSA.cpp:235-252: final weights, thread count of the model, construction then refinement
        auto fw = tree.getWeights();
        for (float& x : fw) x = std::max(x, config->algorithm.finalSaturation);
        auto model = ps.getModel();
        model->setNumThreads(std::max(std::min(config->hardware.numThreads / 2, 8), 1));
        omp_set_num_threads(model->getNumThreads());
        auto constructor = std::shared_ptr<ConstructionStage>(new ConstructionStage(config));
        ColumnRefinement refiner(config, constructor);
        t1 = now_s();
        auto alignment = (*constructor)(fw.data(), cd, &tree, *sequences, sparse, *model);
        double t_build = now_s() - t1;
        {
            std::ofstream f0((out + ".construct").c_str(), std::ios::binary);
            alignment->WriteMFA(f0);
        }
        t1 = now_s();
        alignment = refiner(tree, fw.data(), cd, sparse, *model, std::move(alignment));
        double t_ref = now_s() - t1;
        std::ofstream f(out.c_str(), std::ios::binary);
        alignment->WriteMFA(f);
        printf("{\"tool\": \"ref_qp\", \"mode\": \"msa\", \"t_construct_s\": %.6f, \"t_refine_s\": %.6f}\n", t_build, t_ref);
    }
    printf("{\"tool\": \"ref_qp\", \"n\": %d, \"pairs\": %d, \"threads\": %d, \"cells\": %.0f, \"models\": 2, "
           "\"t_posterior_s\": %.6f, \"t_tree_s\": %.6f, \"t_relax_s\": %.6f, \"reps\": %d}\n",
           N, N * (N - 1) / 2, config->hardware.numThreads, cells, t_post, t_tree, t_cons, cs.iterations);
    return 0;
}
