// Persistent-process mode of the two drop-in executables (MLProbs' driver starts `c_p_np_aln` two or three
// times and `quickprobs` once per region file, utils/do_realign.py:52-63, utils/classifier_c_p_np_aln.py:14,37).
//
// Creating a CUDA context costs 2-3 s on a B200 box; the alignment of a typical benchmark family costs 10-50 ms.  With
// MLP_B200_SERVER=1 in the environment the executable becomes a thin client: it hands its command line, working directory
// and MLP_* environment to a server process of the same program (started on first use, one per user / tool, listening on a
// unix socket under /tmp, exiting after MLP_B200_SERVER_IDLE seconds without a request, default 120) and relays the
// server's stdout, stderr and exit status.  The server runs the very same tool_main() on one long-lived mlp_ctx, so the
// outputs are those of the stand-alone program; if no server can be reached the program simply runs in-process (on the
// GPU as always -- there is no CPU path to fall back to).  Without the variable nothing of this is active.
#pragma once
#include "../../include/mlprobs_b200.h"
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <fcntl.h>
#include <signal.h>
#include <sys/select.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <unistd.h>

extern char** environ;

namespace mlpserve {

struct Exit { int code; };                                         // thrown instead of std::exit() inside tool_main
typedef int (*ToolMain)(int argc, char** argv, mlp_ctx** shared);  // shared != nullptr: keep / reuse the context it points to

namespace detail {

inline bool write_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) { const ssize_t w = ::write(fd, c, n); if (w <= 0) { if (errno == EINTR) continue; return false; } c += w; n -= (size_t)w; }
    return true;
}
inline bool read_all(int fd, void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n) { const ssize_t r = ::read(fd, c, n); if (r <= 0) { if (r < 0 && errno == EINTR) continue; return false; } c += r; n -= (size_t)r; }
    return true;
}
inline bool put_str(int fd, const std::string& s) { const uint32_t n = (uint32_t)s.size(); return write_all(fd, &n, 4) && (n == 0 || write_all(fd, s.data(), n)); }
inline bool get_str(int fd, std::string& s) {
    uint32_t n = 0;
    if (!read_all(fd, &n, 4) || n > (1u << 30)) return false;
    s.resize(n);
    return n == 0 || read_all(fd, &s[0], n);
}
inline std::string socket_path(const char* tool) {
    char buf[128];
    std::snprintf(buf, sizeof buf, "/tmp/mlprobs_b200_%u_%s.sock", (unsigned)getuid(), tool);
    return buf;
}
inline int connect_to(const std::string& path) {
    const int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    sockaddr_un a; std::memset(&a, 0, sizeof a); a.sun_family = AF_UNIX;
    std::strncpy(a.sun_path, path.c_str(), sizeof a.sun_path - 1);
    if (::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof a) != 0) { ::close(fd); return -1; }
    return fd;
}
inline std::string slurp_and_reset(int fd) {
    std::string s;
    const off_t n = ::lseek(fd, 0, SEEK_END);
    if (n > 0) { s.resize((size_t)n); ::lseek(fd, 0, SEEK_SET); if (!read_all(fd, &s[0], (size_t)n)) s.clear(); }
    ::lseek(fd, 0, SEEK_SET);
    if (::ftruncate(fd, 0) != 0) {}
    return s;
}

const uint32_t kMagic = 0x4d4c5032u;   // "MLP2"

// ---- server: one request at a time on one long-lived context
inline int serve(const std::string& path, ToolMain fn) {
    ::signal(SIGPIPE, SIG_IGN);
    const int ls = ::socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0) return 1;
    sockaddr_un a; std::memset(&a, 0, sizeof a); a.sun_family = AF_UNIX;
    std::strncpy(a.sun_path, path.c_str(), sizeof a.sun_path - 1);
    const mode_t old = ::umask(0177);
    const int br = ::bind(ls, reinterpret_cast<sockaddr*>(&a), sizeof a);
    ::umask(old);
    if (br != 0 || ::listen(ls, 16) != 0) return 1;             // another server owns the socket: nothing to do
    int idle = 120;
    if (const char* e = std::getenv("MLP_B200_SERVER_IDLE")) idle = std::max(1, std::atoi(e));
    char tmpl_out[] = "/tmp/mlprobs_b200_out_XXXXXX", tmpl_err[] = "/tmp/mlprobs_b200_err_XXXXXX";
    const int cap_out = ::mkstemp(tmpl_out), cap_err = ::mkstemp(tmpl_err);
    if (cap_out < 0 || cap_err < 0) return 1;
    ::unlink(tmpl_out); ::unlink(tmpl_err);
    mlp_ctx* shared = nullptr;
    for (;;) {
        fd_set rs; FD_ZERO(&rs); FD_SET(ls, &rs);
        timeval tv; tv.tv_sec = idle; tv.tv_usec = 0;
        const int sr = ::select(ls + 1, &rs, nullptr, nullptr, &tv);
        if (sr == 0) break;                                        // idle: give the GPU back
        if (sr < 0) { if (errno == EINTR) continue; break; }
        const int c = ::accept(ls, nullptr, nullptr);
        if (c < 0) continue;
        uint32_t magic = 0, argc = 0, nenv = 0;
        std::string cwd;
        std::vector<std::string> args, envs;
        bool ok = read_all(c, &magic, 4) && magic == kMagic && read_all(c, &argc, 4) && argc < 4096;
        for (uint32_t k = 0; ok && k < argc; ++k) { std::string s; ok = get_str(c, s); args.push_back(s); }
        ok = ok && get_str(c, cwd) && read_all(c, &nenv, 4) && nenv < 4096;
        for (uint32_t k = 0; ok && k < nenv; ++k) { std::string s; ok = get_str(c, s); envs.push_back(s); }
        if (!ok) { ::close(c); continue; }
        // the request's view of the world: working directory, MLP_* variables, stdout / stderr captured in files
        if (::chdir(cwd.c_str()) != 0) {}
        std::vector<std::string> set_names;
        for (const std::string& e : envs) {
            const size_t eq = e.find('=');
            if (eq == std::string::npos) continue;
            ::setenv(e.substr(0, eq).c_str(), e.substr(eq + 1).c_str(), 1);
            set_names.push_back(e.substr(0, eq));
        }
        std::fflush(stdout); std::fflush(stderr); std::cout.flush(); std::cerr.flush();
        const int keep_out = ::dup(1), keep_err = ::dup(2);
        ::dup2(cap_out, 1); ::dup2(cap_err, 2);
        std::vector<char*> argv;
        for (std::string& s : args) argv.push_back(&s[0]);
        argv.push_back(nullptr);
        int32_t status = 1;
        try { status = fn((int)args.size(), argv.data(), &shared); }
        catch (const Exit& e) { status = e.code; }
        catch (const std::exception& e) { std::fprintf(stderr, "%s: %s\n", args.empty() ? "mlprobs_b200" : args[0].c_str(), e.what()); status = 1; }
        std::cout.flush(); std::cerr.flush(); std::fflush(stdout); std::fflush(stderr);
        std::cout.clear(); std::cerr.clear();
        ::dup2(keep_out, 1); ::dup2(keep_err, 2);
        ::close(keep_out); ::close(keep_err);
        for (const std::string& nm : set_names) ::unsetenv(nm.c_str());
        const std::string so = slurp_and_reset(cap_out), se = slurp_and_reset(cap_err);
        if (write_all(c, &status, 4)) { put_str(c, so) && put_str(c, se); }
        ::close(c);
    }
    ::close(ls);
    ::unlink(path.c_str());
    if (shared) mlp_destroy(shared);
    return 0;
}

// ---- client: returns the exit status of the request, or -1 when no server could be used (caller runs in-process)
inline int client(const char* tool, int argc, char** argv) {
    const std::string path = socket_path(tool);
    int fd = connect_to(path);
    if (fd < 0) {
        // no server yet (or a stale socket file): start one from this very executable, detached from the caller
        ::unlink(path.c_str());
        char self[4096];
        const ssize_t n = ::readlink("/proc/self/exe", self, sizeof self - 1);
        if (n <= 0) return -1;
        self[n] = 0;
        const pid_t pid = ::fork();
        if (pid < 0) return -1;
        if (pid == 0) {
            if (::fork() != 0) ::_exit(0);                      // double fork: the server is not a child of the caller
            ::setsid();
            const int dn = ::open("/dev/null", O_RDWR);
            if (dn >= 0) { ::dup2(dn, 0); ::dup2(dn, 1); ::dup2(dn, 2); if (dn > 2) ::close(dn); }
            char* sargv[] = {self, const_cast<char*>("--serve"), const_cast<char*>(path.c_str()), nullptr};
            ::execv(self, sargv);
            ::_exit(127);
        }
        int st = 0;
        ::waitpid(pid, &st, 0);
        for (int tries = 0; tries < 400 && fd < 0; ++tries) {  // the server binds its socket before it touches CUDA
            ::usleep(25000);
            fd = connect_to(path);
        }
        if (fd < 0) return -1;
    }
    char cwd[4096];
    if (!::getcwd(cwd, sizeof cwd)) { ::close(fd); return -1; }
    std::vector<std::string> envs;
    for (char** e = environ; e && *e; ++e)
        if (std::strncmp(*e, "MLP_", 4) == 0 && std::strncmp(*e, "MLP_B200_SERVER", 15) != 0) envs.push_back(*e);
    const uint32_t ac = (uint32_t)argc, ne = (uint32_t)envs.size();
    bool ok = write_all(fd, &kMagic, 4) && write_all(fd, &ac, 4);
    for (int k = 0; ok && k < argc; ++k) ok = put_str(fd, argv[k]);
    ok = ok && put_str(fd, cwd) && write_all(fd, &ne, 4);
    for (size_t k = 0; ok && k < envs.size(); ++k) ok = put_str(fd, envs[k]);
    int32_t status = 1;
    std::string so, se;
    ok = ok && read_all(fd, &status, 4) && get_str(fd, so) && get_str(fd, se);
    ::close(fd);
    if (!ok) return -1;                                          // the server went away mid-request: run it here instead
    if (!so.empty()) write_all(1, so.data(), so.size());
    if (!se.empty()) write_all(2, se.data(), se.size());
    return status & 0xff;
}

}  // namespace detail

// main() of both executables: server, client or plain in-process run
inline int run(const char* tool, int argc, char** argv, ToolMain fn) {
    if (argc >= 3 && std::strcmp(argv[1], "--serve") == 0) return detail::serve(argv[2], fn);
    const char* mode = std::getenv("MLP_B200_SERVER");
    if (mode && *mode && std::strcmp(mode, "0") != 0) {
        const int st = detail::client(tool, argc, argv);
        if (st >= 0) return st;
    }
    try { return fn(argc, argv, nullptr); }
    catch (const Exit& e) { return e.code; }
}

}  // namespace mlpserve
