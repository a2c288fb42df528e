// c_p_np_aln_b200: command-line drop-in for `c_p_np_aln` (baseMSA/C_P_NP_Aln, MSA::MSA MSA.cpp:123-187) for the ways
// MLProbs calls it (utils/prepare_features_4_classifier_1.py:12, utils/classifier_c_p_np_aln.py:14,37):
//   c_p_np_aln -G file      -> the feature line of MSA::Alter_ModelAdjustmentTest on stdout
//   c_p_np_aln -p 0 file    -> model selection, all-pairs posteriors, tree, consistency, progressive alignment, refinement
//   c_p_np_aln -p 1 file    -> model selection, all-pairs posteriors, consistency, alignment graph, similar-set refinement
// every O(N^2 L^2) stage on the GPU through the C ABI of include/mlprobs_b200.h.  The reference reseeds rand() from the wall
// clock before every refinement sweep of -p 1 (MSA.cpp:1896); so does this program, unless `--seed S` (an extension, also
// read from the environment variable MLP_CPNP_SEED) pins the value the clock would have returned.
// Options kept: -p, -G, -o/--outfile, -c/--consistency, -ir/--iterative-refinement, -v.  No CPU fallback.
#include "../../include/mlprobs_b200.h"
#include "serve.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <chrono>
#include <dirent.h>
#include <sys/stat.h>

namespace {

struct Input { std::vector<std::string> headers, seqs; };

// Sequence::Sequence(FileBuffer&, stripGaps = true), Sequence.h:52-122
bool load_mfa(const std::string& path, Input& in) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.is_open()) { std::cerr << "ERROR: Could not open file '" << path << "' for reading." << std::endl; throw mlpserve::Exit{1}; }
    std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t p = 0;
    while (p < all.size()) {
        std::string header;
        while (p < all.size()) {                               // first non-blank line
            size_t e = all.find('\n', p);
            if (e == std::string::npos) e = all.size();
            header = all.substr(p, e - p);
            p = std::min(all.size(), e + 1);
            if (!header.empty()) break;
        }
        if (header.empty() || header[0] != '>') break;
        header = header.substr(1);
        while (!header.empty() && isspace((unsigned char)header[0])) header.erase(0, 1);
        while (!header.empty() && isspace((unsigned char)header[header.size() - 1])) header.erase(header.size() - 1);
        std::string data;
        while (p < all.size() && all[p] != '>') {
            char ch = all[p++];
            if (isspace((unsigned char)ch)) continue;
            if (ch == '.' || ch == '-') continue;               // stripGaps
            if (!((ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z'))) {
                std::cerr << "ERROR: Unknown character encountered: " << ch << std::endl;
                throw mlpserve::Exit{1};
            }
            if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 'a' + 'A');
            data.push_back(ch);
        }
        if (data.empty()) break;                               // an empty record ends the file for the reference too
        in.headers.push_back(header);
        in.seqs.push_back(data);
    }
    return !in.seqs.empty();
}

void write_mfa(std::ostream& out, const std::string& header, const char* row, int len) {
    out << ">" << header << "\n";
    for (int p = 0; p < len; p += 60) { out.write(row + p, std::min(60, len - p)); out << "\n"; }
}

int fail(mlp_ctx* ctx, const char* what, int rc) {
    std::fprintf(stderr, "c_p_np_aln_b200: %s failed (%d)%s%s\n", what, rc, ctx ? ": " : "", ctx ? mlp_last_error(ctx) : "");
    return 1;                                          // the context belongs to main (or to the server): never destroyed here
}

}  // namespace

// one family (the reference loads every positional file into the same sequence set, MSA.cpp:133-137); returns the exit status
int run_file(mlp_ctx* ctx, const std::vector<std::string>& infiles, const std::string& outfile, int getpid, int reps, int refine, int verbose,
             int program, long long seed) {
    const std::string& infile = infiles[0];
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    auto fail = [&](mlp_ctx* c, const char* what, int rc) {
        std::fprintf(stderr, "c_p_np_aln_b200: %s: %s failed (%d): %s\n", infile.c_str(), what, rc, c ? mlp_last_error(c) : "");
        return 1;
    };
    Input in;
    for (const std::string& f : infiles) load_mfa(f, in);
    if (in.seqs.empty()) { std::cerr << "ERROR: No sequences read." << std::endl; return 1; }
    const int n = (int)in.seqs.size();
    std::ofstream fout;
    if (!outfile.empty()) {
        fout.open(outfile.c_str(), std::ios::binary | std::ios::out | std::ios::trunc);
        if (!fout.is_open()) { std::cerr << "ERROR: Failed to open the file " << outfile << std::endl; return 1; }
    }
    std::ostream& out = outfile.empty() ? std::cout : fout;

    int rc = 0;
    if (n < 2) {
        // one sequence: the reference crashes (-p 0/1: SIGSEGV, -G: SIGFPE) with nothing on stdout, and MLProbs reads the
        // non-zero status as "fall back to quickprobs" (utils/classifier_c_p_np_aln.py:40-41): same contract here
        std::cerr << "ERROR: at least two sequences are required." << std::endl;
        return 1;
    }
    std::vector<int32_t> len(n);
    std::string cat;
    for (int i = 0; i < n; ++i) { len[i] = (int32_t)in.seqs[i].size(); cat += in.seqs[i]; }
    const long long npairs = (long long)n * (n - 1) / 2;
    mlp_hmm_tables hmm;
    mlp_part_tables part;
    if ((rc = mlp_default_tables(MLP_CPNP_P0, 0.700645f, &hmm, &part))) return fail(ctx, "mlp_default_tables", rc);
    if ((rc = mlp_set_tables(ctx, &hmm, &part))) return fail(ctx, "mlp_set_tables", rc);
    if ((rc = mlp_set_sequences(ctx, n, len.data(), (const uint8_t*)cat.data()))) return fail(ctx, "mlp_set_sequences", rc);
    std::vector<int32_t> ident(npairs), alen(npairs);
    if (getpid) {
        // MSA::Alter_ModelAdjustmentTest(sequences, 1.0), MSA.cpp:154-165,646-762
        long long cap = 0;
        for (int a = 0; a < n; ++a) for (int b = a + 1; b < n; ++b) cap += len[a] + len[b];
        std::vector<char> aln((size_t)cap + 16);
        std::vector<int64_t> off(npairs + 1);
        if ((rc = mlp_viterbi_all_pairs_ex(ctx, ident.data(), alen.data(), aln.data(), off.data()))) return fail(ctx, "mlp_viterbi_all_pairs_ex", rc);
        char line[512];
        if ((rc = mlp_cpnp_g_features(n, len.data(), (const uint8_t*)cat.data(), aln.data(), off.data(), 1.0f, line, (int)sizeof line)))
            return fail(ctx, "mlp_cpnp_g_features", rc);
        out << line << "\n";
        return 0;
    }
    // MSA::ModelAdjustmentTest, MSA.cpp:775-882
    const double t1 = now();
    if ((rc = mlp_viterbi_all_pairs(ctx, ident.data(), alen.data()))) return fail(ctx, "mlp_viterbi_all_pairs", rc);
    float identity = 0, sigma = 0, init2 = 0;
    const int variance_mean = mlp_cpnp_model_adjustment(npairs, ident.data(), alen.data(), &identity, &sigma, &init2);
    if (variance_mean < 0) return fail(ctx, "mlp_cpnp_model_adjustment", variance_mean);
    const int pid = variance_mean % 10, vpid = variance_mean / 10;
    if ((rc = mlp_default_tables(MLP_CPNP_P0, init2, &hmm, &part))) return fail(ctx, "mlp_default_tables", rc);
    if ((rc = mlp_set_tables(ctx, &hmm, &part))) return fail(ctx, "mlp_set_tables", rc);
    const double t2 = now();
    const uint32_t mask = pid <= 1 ? (MLP_M_HMM5 | MLP_M_PART | MLP_M_LOCAL) : (pid == 2 ? MLP_M_LOCAL : MLP_M_PART);   // MSA.cpp:946-1010
    rc = mlp_posterior_all_pairs(ctx, program == 1 ? MLP_CPNP_P1 : MLP_CPNP_P0, mask, 0.01f);   // -p 1: ArrangePosteriorProbs, MSA.cpp:1635-1766
    if (rc == MLP_E_OVERFLOW) { std::printf("ERROR: huge val error for zM\n"); return 1; }            // MSAPartProbs.cpp:547-589
    if (rc) return fail(ctx, "mlp_posterior_all_pairs", rc);
    const double t3 = now();
    if (program == 1) {
        // MSA::npdoAlign, MSA.cpp:1117-1136: the same relaxation, then the alignment graph and DoRefinement; rows in input order
        for (int r = 0; r < reps; ++r)
            if ((rc = mlp_relax(ctx, MLP_CPNP_P0, nullptr, nullptr, 0.0f, 0.0f, 0.01f))) return fail(ctx, "mlp_relax", rc);
        const double t4 = now();
        char* rows = nullptr;
        int32_t cols = 0;
        if ((rc = mlp_cpnp_np_finish_alignment(ctx, refine, seed, &rows, &cols))) return fail(ctx, "mlp_cpnp_np_finish_alignment", rc);
        for (int k = 0; k < n; ++k) write_mfa(out, in.headers[k], rows + (size_t)k * cols, cols);
        if (verbose) std::fprintf(stderr, "c_p_np_aln_b200: %d sequences, model class %d, %d columns; ms: load+upload %.1f, viterbi %.1f, posterior %.1f, consistency %.1f, graph+refinement %.1f\n",
                                  n, variance_mean, cols, t1 - t0, t2 - t1, t3 - t2, t4 - t3, now() - t4);
        mlp_free_host(rows);
        return 0;
    }
    std::vector<float> dist((size_t)n * n);
    std::vector<int32_t> weights(n), left(2 * n - 1), right(2 * n - 1), order(n);
    if ((rc = mlp_get_distances(ctx, dist.data()))) return fail(ctx, "mlp_get_distances", rc);
    if ((rc = mlp_cpnp_guide_tree(n, dist.data(), vpid, weights.data(), left.data(), right.data()))) return fail(ctx, "mlp_cpnp_guide_tree", rc);
    for (int r = 0; r < reps; ++r)
        if ((rc = mlp_relax(ctx, MLP_CPNP_P0, nullptr, nullptr, 0.0f, 0.0f, 0.01f))) return fail(ctx, "mlp_relax", rc);
    const double t4 = now();
    char* rows = nullptr;
    int32_t cols = 0;
    if ((rc = mlp_cpnp_finish_alignment(ctx, weights.data(), left.data(), right.data(), refine, pid, &rows, &cols, order.data())))
        return fail(ctx, "mlp_cpnp_finish_alignment", rc);
    for (int k = 0; k < n; ++k) write_mfa(out, in.headers[order[k]], rows + (size_t)k * cols, cols);
    if (verbose) std::fprintf(stderr, "c_p_np_aln_b200: %d sequences, model class %d, %d columns; ms: load+upload %.1f, viterbi %.1f, posterior %.1f, tree+consistency %.1f, alignment+refinement %.1f\n",
                              n, variance_mean, cols, t1 - t0, t2 - t1, t3 - t2, t4 - t3, now() - t4);
    mlp_free_host(rows);
    return 0;
}

// the program proper; `shared` (persistent-process mode, serve.h) points to a context that outlives the call
static int tool_main(int argc, char** argv, mlp_ctx** shared) {
    std::string outfile;
    std::vector<std::string> infiles;
    int program = 0, getpid = 0, reps = 2, refine = 100, device = 0, verbose = 0;
    long long seed = -1;
    if (const char* e = std::getenv("MLP_CPNP_SEED")) seed = std::atoll(e);
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::cerr << "ERROR: Must specify a value after option " << name << "." << std::endl; throw mlpserve::Exit{1}; }
            return argv[++i];
        };
        if (a == "-o" || a == "--outfile") outfile = need("-o");
        else if (a == "-p" || a == "--program") {
            program = std::atoi(need("-p"));
            if (program < 0 || program > 1) { std::cerr << "ERROR: For option -p, integer must be 0 or 1." << std::endl; return 1; }
        }
        else if (a == "-G" || a == "--getPID") getpid = 1;
        else if (a == "-c" || a == "--consistency") reps = std::atoi(need("-c"));
        else if (a == "-ir" || a == "--iterative-refinement") refine = std::atoi(need("-ir"));
        else if (a == "-d" || a == "--device") device = std::atoi(need("-d"));
        else if (a == "--seed") seed = std::atoll(need("--seed"));
        else if (a == "-v" || a == "--verbose") verbose = 1;
        else if (a == "-clustalw" || a == "-timeon" || a == "-timeoff") {}            // parsed and without effect on the output in the reference too (MSA.cpp:182,404-409)
        else if (a == "-version") { std::cerr << "c_p_np_aln_b200 (drop-in for PNPProbs c_p_np_aln)" << std::endl; return 1; }   // MSA.cpp:417-420: stderr, exit 1
        else if (a == "-a" || a == "--alignment-order" || a == "-annot" || a == "-co" || a == "--cutoff") {
            std::cerr << "ERROR: option " << a << " is not supported by c_p_np_aln_b200" << std::endl;
            return 1;
        }
        else if (!a.empty() && a[0] == '-') { std::cerr << "ERROR: Unrecognized option: " << a << std::endl; return 1; }   // MSA.cpp:422-425
        else infiles.push_back(a);
    }
    if (infiles.empty()) { std::fprintf(stderr, "usage: c_p_np_aln_b200 (-G | -p 0 | -p 1 [--seed S]) [-o outfile] [-c reps] [-ir passes] <fasta>\n"); return 1; }   // the reference prints its usage on stderr and exits 1
    struct stat si, so;
    const std::string& infile = infiles[0];
    const bool dir_mode = infiles.size() == 1 && !outfile.empty() && stat(infile.c_str(), &si) == 0 && S_ISDIR(si.st_mode) &&
                          stat(outfile.c_str(), &so) == 0 && S_ISDIR(so.st_mode);   // extension: a directory of families, one CUDA context
    if (!dir_mode) {
        Input probe;
        for (const std::string& f : infiles) load_mfa(f, probe);
        if (probe.seqs.empty()) { std::cerr << "ERROR: No sequences read." << std::endl; return 1; }
        if (probe.seqs.size() < 2) { std::cerr << "ERROR: at least two sequences are required." << std::endl; return 1; }   // see run_file
    }
    mlp_ctx* ctx = shared ? *shared : nullptr;
    int rc = ctx ? 0 : mlp_create(device, &ctx);                 // no CUDA device -> stop: nothing falls back to the CPU
    if (rc) return fail(nullptr, "mlp_create (a CUDA device is required)", rc);
    if (shared) *shared = ctx;
    int status = 0;
    if (dir_mode) {
        std::vector<std::string> names;
        if (DIR* d = opendir(infile.c_str())) {
            while (struct dirent* e = readdir(d)) {
                struct stat st;
                const std::string path = infile + "/" + e->d_name;
                if (stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode)) names.push_back(e->d_name);
            }
            closedir(d);
        }
        std::sort(names.begin(), names.end());
        for (const std::string& nm : names) {
            const int r1 = run_file(ctx, std::vector<std::string>(1, infile + "/" + nm), outfile + "/" + nm, getpid, reps, refine, verbose, program, seed);
            if (r1) status = r1;
        }
    } else status = run_file(ctx, infiles, outfile, getpid, reps, refine, verbose, program, seed);
    if (!shared) mlp_destroy(ctx);
    return status;
}

int main(int argc, char** argv) { return mlpserve::run("c_p_np_aln_b200", argc, argv, tool_main); }
