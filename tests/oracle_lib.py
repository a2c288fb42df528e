"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

QP, CPNP_P0, CPNP_P1 = 0, 1, 2
M_HMM5, M_PART, M_LOCAL = 1, 2, 4


class HmmTables(C.Structure):
    _fields_ = [("init", C.c_float * 5), ("trans", (C.c_float * 5) * 5), ("match", (C.c_float * 26) * 26),
                ("ins", C.c_float * 26), ("ltrans", (C.c_float * 3) * 3), ("rtrans", C.c_float * 2)]


class PartTables(C.Structure):
    _fields_ = [("sub", (C.c_double * 26) * 26), ("go", C.c_double), ("ge", C.c_double),
                ("tgo", C.c_double), ("tge", C.c_double)]


class CsrSetC(C.Structure):
    _fields_ = [("n", C.c_int), ("len", C.c_void_p), ("rp_off", C.c_void_p), ("nz_off", C.c_void_p),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
                ("rp_cap", C.c_int64), ("nz_cap", C.c_int64), ("rp_used", C.c_int64), ("nz_used", C.c_int64)]


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "mlp_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_log_add.restype = C.c_float
        _lib.orc_log_add.argtypes = [C.c_float, C.c_float]
        _lib.orc_exp.restype = C.c_float
        _lib.orc_exp.argtypes = [C.c_float]
        _lib.orc_combine_qp.restype = C.c_float
        _lib.orc_mea_score.restype = C.c_float
        _lib.orc_init_distrib2_for_identity.restype = C.c_float
        _lib.orc_init_distrib2_for_identity.argtypes = [C.c_float]
        _lib.orc_build_hmm.argtypes = [C.c_float, C.c_void_p]
        _lib.orc_sparsify.restype = C.c_int64
    return _lib


def hmm_tables(init_distrib2=0.700645):
    t = HmmTables()
    lib().orc_build_hmm(C.c_float(init_distrib2), C.byref(t))
    return t


def part_tables(flavour):
    t = PartTables()
    (lib().orc_build_part_qp if flavour == QP else lib().orc_build_part_cpnp)(C.byref(t))
    return t


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pair_posterior(flavour, mask, ht, pt, s1: bytes, s2: bytes):
    L1, L2 = len(s1), len(s2)
    post = np.zeros((L1 + 1, L2 + 1), np.float32)
    dist = C.c_float(0)
    rc = lib().orc_pair_posterior(flavour, mask, C.byref(ht), C.byref(pt), s1, L1, s2, L2, _p(post), C.byref(dist))
    return post, dist.value, rc


def model_posterior(which, ht, pt, s1: bytes, s2: bytes, qp_quirk=0):
    L1, L2 = len(s1), len(s2)
    post = np.zeros((L1 + 1, L2 + 1), np.float32)
    tot = C.c_float(0)
    if which == "hmm5":
        lib().orc_hmm5_posterior(C.byref(ht), s1, L1, s2, L2, qp_quirk, _p(post), C.byref(tot))
    elif which == "local":
        lib().orc_local_posterior(C.byref(ht), s1, L1, s2, L2, _p(post), C.byref(tot))
    elif which == "part_qp":
        lib().orc_part_posterior_qp(C.byref(pt), s1, L1, s2, L2, _p(post))
    elif which == "part_cpnp":
        lib().orc_part_posterior_cpnp(C.byref(pt), s1, L1, s2, L2, _p(post))
    return post, tot.value


def viterbi(ht, s1: bytes, s2: bytes):
    ident = C.c_int(0); ln = C.c_int(0)
    aln = C.create_string_buffer(len(s1) + len(s2) + 2)
    lib().orc_viterbi.restype = C.c_float
    p = lib().orc_viterbi(C.byref(ht), s1, len(s1), s2, len(s2), C.byref(ident), C.byref(ln), aln)
    return p, ident.value, ln.value, aln.raw[:ln.value]


def model_adjustment(n_identical, aln_len):
    ni = np.ascontiguousarray(n_identical, np.int32); al = np.ascontiguousarray(aln_len, np.int32)
    ident = C.c_float(0); sig = C.c_float(0); i2 = C.c_float(0)
    vm = lib().orc_model_adjustment(len(ni), _p(ni), _p(al), C.byref(ident), C.byref(sig), C.byref(i2))
    return vm, ident.value, sig.value, i2.value


def g_features(ht, seqs, theta=1.0):
    lens, off, cat = _seqs(seqs)
    line = C.create_string_buffer(256)
    rc = lib().orc_g_features(C.byref(ht), len(seqs), _p(lens), cat, _p(off), C.c_float(theta), line, 256)
    return rc, line.value


class CsrSet:
    """Both orientations of every pair, pooled (mirrors orc_csr_set)."""

    def __init__(self, lens, nz_cap=None):
        self.lens = np.ascontiguousarray(lens, np.int32)
        n = len(lens)
        self.n = n
        tot = int(self.lens.sum())
        self.rp_cap = (n - 1) * (tot + 2 * n) + 16
        if nz_cap is None:
            nz_cap = 16 * (n - 1) * tot + 1024
        self.nz_cap = int(nz_cap)
        self.rp_off = np.full(n * n, -1, np.int64)
        self.nz_off = np.full(n * n, -1, np.int64)
        self.rowptr = np.zeros(self.rp_cap, np.int32)
        self.col = np.zeros(self.nz_cap, np.int32)
        self.val = np.zeros(self.nz_cap, np.float32)
        self.c = CsrSetC(n, _p(self.lens), _p(self.rp_off), _p(self.nz_off), _p(self.rowptr), _p(self.col),
                         _p(self.val), self.rp_cap, self.nz_cap, 0, 0)

    def get(self, a, b):
        La = int(self.lens[a])
        ro = int(self.rp_off[a * self.n + b]); no = int(self.nz_off[a * self.n + b])
        rp = self.rowptr[ro:ro + La + 2]
        nz = int(rp[La + 1])
        return rp, self.col[no:no + nz], self.val[no:no + nz]


def _seqs(seqs):
    lens = np.array([len(s) for s in seqs], np.int32)
    off = np.zeros(len(seqs), np.int64)
    off[1:] = np.cumsum(lens)[:-1]
    return lens, off, b"".join(seqs)


def posterior_stage(flavour, mask, ht, pt, seqs, cutoff=0.01, threads=1, nz_cap=None):
    lens, off, cat = _seqs(seqs)
    n = len(seqs)
    dist = np.zeros((n, n), np.float32)
    out = CsrSet(lens, nz_cap)
    rc = lib().orc_posterior_stage(flavour, mask, C.byref(ht), C.byref(pt), n, _p(lens), cat, _p(off),
                                   C.c_float(cutoff), _p(dist), C.byref(out.c), threads)
    return dist, out, rc


def relax_cpnp(inp: CsrSet, cutoff=0.01, threads=1):
    out = CsrSet(inp.lens, inp.nz_cap)
    rc = lib().orc_relax_cpnp(C.byref(inp.c), C.c_float(cutoff), C.byref(out.c), threads)
    assert rc == 0
    return out


def relax_qp(inp: CsrSet, weights, seldist, cutoff, selectivity=200.0, selfweight=3.0, threads=1):
    out = CsrSet(inp.lens, inp.nz_cap)
    w = np.ascontiguousarray(weights, np.float32)
    sd = np.ascontiguousarray(seldist, np.float32)
    rc = lib().orc_relax_qp(C.byref(inp.c), _p(w), _p(sd), C.c_float(selectivity), C.c_float(selfweight),
                            C.c_float(cutoff), C.byref(out.c), threads)
    assert rc == 0
    return out
