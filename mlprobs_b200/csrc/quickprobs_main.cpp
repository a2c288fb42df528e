// quickprobs_b200: command-line drop-in for `quickprobs` (realign/QuickProbs, Console/main.cpp:18-66) in its default
// protein configuration, every stage on the GPU through the C ABI of include/mlprobs_b200.h:
//   FASTA in (SequenceIO::loadFasta SequenceIO.cpp:98-155, checkAndCorrect :70-93) -> posterior stage -> UPGMA tree ->
//   consistency (1 repetition above 50 sequences, else 2; last one unfiltered, ConsistencyStage.cpp:73-123) ->
//   progressive construction + column refinement -> FASTA out, 60 columns per line (SequenceIO::saveFasta :175-193).
// Options kept: positional infile, -o/--outfile, -c/--con-iters, -r/--ref-count, --ref-seed, -t/--num-threads (accepted,
// ignored: there is no thread team), -v.  Anything else the reference offers (nucleotide mode, other trees, OpenCL
// device selection, ClustalW output) is refused loudly rather than approximated.  There is no CPU fallback.
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

bool load_fasta(const std::string& path, Input& in, std::string& err) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.is_open()) { err = "unable to open input file"; return false; }
    std::string line;
    bool have = false;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        if (line[0] == '>') {
            std::string h = line.substr(1);
            while (!h.empty() && isspace((unsigned char)h[0])) h.erase(0, 1);
            while (!h.empty() && isspace((unsigned char)h[h.size() - 1])) h.erase(h.size() - 1);
            in.headers.push_back(h);
            in.seqs.push_back(std::string());
            have = true;
        } else if (have) {
            if (line[line.size() - 1] == '\r') line.erase(line.size() - 1);
            in.seqs.back() += line;
        }
    }
    if (in.seqs.empty()) { err = "no sequences read"; return false; }
    bool ok = true;
    for (auto& s : in.seqs)
        for (auto& c : s) {
            if (isalpha((unsigned char)c)) c = (char)toupper((unsigned char)c);
            else { std::cout << "illegal sequence character:" << c << std::endl; ok = false; }
        }
    if (!ok) { err = "Illegal characters in sequence set!"; return false; }
    for (auto& s : in.seqs) if (s.empty()) { err = "empty sequence in input"; return false; }
    return true;
}

void write_fasta(std::ostream& out, const std::vector<std::string>& headers, const char* rows, int n, int len) {
    for (int i = 0; i < n; ++i) {
        out << ">" << headers[i] << "\n";
        const char* r = rows + (size_t)i * len;
        for (int p = 0; p < len; p += 60) { out.write(r + p, std::min(60, len - p)); out << "\n"; }
    }
}

int fail(mlp_ctx* ctx, const char* what, int rc) {
    std::fprintf(stderr, "quickprobs_b200: %s failed (%d)%s%s\n", what, rc, ctx ? ": " : "", ctx ? mlp_last_error(ctx) : "");
    return 1;                                          // the context belongs to main (or to the server): never destroyed here
}

}  // namespace

// one input file -> one alignment; returns the process exit status for this file (0 = ok)
int align_file(mlp_ctx* ctx, const std::string& infile, const std::string& outfile, int con_iters, int ref_count, unsigned ref_seed, int verbose) {
    auto fail = [&](mlp_ctx* c, const char* what, int rc) {
        std::fprintf(stderr, "quickprobs_b200: %s: %s failed (%d): %s\n", infile.c_str(), what, rc, c ? mlp_last_error(c) : "");
        return 1;
    };
    int rc = 0;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    Input in;
    std::string err;
    if (!load_fasta(infile, in, err)) { std::fprintf(stderr, "quickprobs_b200: %s\n", err.c_str()); return 255; }
    const int n = (int)in.seqs.size();
    std::ofstream fout;
    if (!outfile.empty()) {
        fout.open(outfile.c_str(), std::ios::binary | std::ios::out | std::ios::trunc);
        if (!fout.is_open()) { std::fprintf(stderr, "quickprobs_b200: unable to open output file\n"); return 255; }
    }
    std::ostream& out = outfile.empty() ? std::cout : fout;

    if (n == 1) {
        write_fasta(out, in.headers, in.seqs[0].data(), 1, (int)in.seqs[0].size());
        return 0;
    }
    std::vector<int32_t> len(n);
    std::string cat;
    for (int i = 0; i < n; ++i) { len[i] = (int32_t)in.seqs[i].size(); cat += in.seqs[i]; }
    mlp_hmm_tables hmm;
    mlp_part_tables part;
    if ((rc = mlp_default_tables(MLP_QP, 0.0f, &hmm, &part))) return fail(ctx, "mlp_default_tables", rc);
    if ((rc = mlp_set_tables(ctx, &hmm, &part))) return fail(ctx, "mlp_set_tables", rc);
    if ((rc = mlp_set_sequences(ctx, n, len.data(), (const uint8_t*)cat.data()))) return fail(ctx, "mlp_set_sequences", rc);
    const double t1 = now();
    if ((rc = mlp_posterior_all_pairs(ctx, MLP_QP, MLP_M_HMM5 | MLP_M_PART, 0.01f))) return fail(ctx, "mlp_posterior_all_pairs", rc);
    const double t2 = now();
    std::vector<float> weights(n), seldist;
    std::vector<int32_t> left(2 * n - 1), right(2 * n - 1);
    // guide tree on the device, weights saturated at 1e-6 (consistency.saturation == finalSaturation); families too large for the
    // single-CTA clustering take the host tree on a read-back of the distances
    bool resident = true;
    rc = mlp_qp_guide_tree_device(ctx, 1e-6f, weights.data(), nullptr, left.data(), right.data(), nullptr);
    if (rc == MLP_E_UNSUPPORTED) {
        resident = false;
        std::vector<float> dist((size_t)n * n);
        seldist.resize((size_t)n * n);
        if ((rc = mlp_get_distances(ctx, dist.data()))) return fail(ctx, "mlp_get_distances", rc);
        if ((rc = mlp_qp_guide_tree_ex(n, dist.data(), weights.data(), seldist.data(), nullptr, left.data(), right.data())))
            return fail(ctx, "mlp_qp_guide_tree_ex", rc);
        for (auto& w : weights) w = std::max(w, 1e-6f);
    } else if (rc) return fail(ctx, "mlp_qp_guide_tree_device", rc);
    const int iters = con_iters >= 0 ? con_iters : (n > 50 ? 1 : 2);
    for (int it = 0; it < iters; ++it) {
        const float cutoff = (it == iters - 1) ? 1e-5f : 0.01f;
        if ((rc = mlp_relax(ctx, MLP_QP, resident ? nullptr : weights.data(), resident ? nullptr : seldist.data(), 200.0f, 3.0f, cutoff))) return fail(ctx, "mlp_relax", rc);
    }
    const double t3 = now();
    char* rows = nullptr;
    int32_t alen = 0;
    if ((rc = mlp_qp_finish_alignment(ctx, weights.data(), left.data(), right.data(), ref_count, ref_seed, &rows, &alen)))
        return fail(ctx, "mlp_qp_finish_alignment", rc);
    write_fasta(out, in.headers, rows, n, alen);
    if (verbose) std::fprintf(stderr, "quickprobs_b200: %d sequences, %d columns; ms: load+upload %.1f, posterior %.1f, tree+consistency %.1f, construction+refinement %.1f\n",
                              n, alen, t1 - t0, t2 - t1, t3 - t2, now() - t3);
    mlp_free_host(rows);
    return 0;
}

// the program proper; `shared` (persistent-process mode, serve.h) points to a context that outlives the call
static int tool_main(int argc, char** argv, mlp_ctx** shared) {
    std::string infile, outfile;
    int con_iters = -1, ref_count = -1, device = 0, verbose = 0;
    unsigned ref_seed = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "quickprobs_b200: option %s needs a value\n", name); throw mlpserve::Exit{2}; }
            return argv[++i];
        };
        // ProgramOptions::parse (Common/ProgramOptions.cpp:15-48): any number of leading dashes, long or short name
        std::string name = a;
        while (!name.empty() && name[0] == '-') name.erase(0, 1);
        const bool dashed = !a.empty() && a[0] == '-';
        if (!dashed) {
            if (infile.empty()) infile = a;
            else { std::fprintf(stderr, "quickprobs_b200: more than one input file\n"); return 2; }
        }
        else if (name == "o" || name == "outfile") outfile = need("-o");
        else if (name == "c" || name == "con-iters") con_iters = std::atoi(need("-c"));
        else if (name == "r" || name == "ref-count") ref_count = std::atoi(need("-r"));
        else if (name == "ref-seed") ref_seed = (unsigned)std::strtoul(need("--ref-seed"), nullptr, 10);
        else if (name == "t" || name == "num-threads" || name == "p" || name == "platform" || name == "mem-limit" || name == "ref-threads")
            (void)need(a.c_str());                   // host threads, OpenCL platform, memory limit: no meaning here, value skipped
        else if (name == "d" || name == "device") device = std::atoi(need("-d"));    // the reference's OpenCL device id; here the CUDA device
        else if (name == "v" || name == "verbose") verbose = 1;
        else { std::fprintf(stderr, "quickprobs_b200: unsupported option %s\n", a.c_str()); return 2; }   // incl. -n/--nucleotide, -l/--clustalw
    }
    if (infile.empty()) {
        // the reference prints its usage on stderr and returns 0 when the positional argument is missing (Console/main.cpp:27-36)
        std::fprintf(stderr, "usage: quickprobs_b200 <infile | indir> [-o outfile | outdir] [-c con-iters] [-r ref-count] [--ref-seed S] [-d cuda-device]\n");
        return 0;
    }
    // directory mode (Configuration.cpp:248-267): every regular file of the input directory is aligned into a file of the same
    // name in the output directory -- one process, one CUDA context for all of them
    struct stat si, so;
    const bool dir_mode = !outfile.empty() && stat(infile.c_str(), &si) == 0 && S_ISDIR(si.st_mode) &&
                          stat(outfile.c_str(), &so) == 0 && S_ISDIR(so.st_mode);
    if (!dir_mode) {                                   // input errors are reported before a device is required
        Input probe;
        std::string err;
        if (!load_fasta(infile, probe, err)) { std::fprintf(stderr, "quickprobs_b200: %s\n", err.c_str()); return 255; }
    }
    mlp_ctx* ctx = shared ? *shared : nullptr;
    int rc = ctx ? 0 : mlp_create(device, &ctx);                 // no CUDA device -> MLP_E_NO_DEVICE: stop here, nothing falls back to the CPU
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
            const int r1 = align_file(ctx, infile + "/" + nm, outfile + "/" + nm, con_iters, ref_count, ref_seed, verbose);
            if (r1) status = r1;
        }
    } else status = align_file(ctx, infile, outfile, con_iters, ref_count, ref_seed, verbose);
    if (!shared) mlp_destroy(ctx);
    return status;
}

int main(int argc, char** argv) { return mlpserve::run("quickprobs_b200", argc, argv, tool_main); }
