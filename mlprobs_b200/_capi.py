"""ctypes binding of libmlprobs_b200.so (include/mlprobs_b200.h).

The library is the product; this module only marshals numpy arrays across the C ABI.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is visible, the calls raise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLP_B200_LIB") or os.path.join(_HERE, "libmlprobs_b200.so")   # MLP_B200_LIB: developer override (A/B builds)

QP, CPNP_P0, CPNP_P1 = 0, 1, 2
M_HMM5, M_PART, M_LOCAL = 1, 2, 4
K_NAMES = ["part_fwd", "part_rev", "hmm_fwd", "hmm_bwd", "local_fwd", "local_bwd", "final", "relax"]

ERRORS = {0: "ok", -1: "no CUDA device", -2: "CUDA error", -3: "bad argument", -4: "bad state",
          -5: "capacity exhausted", -6: "partition function overflow", -7: "unsupported", -8: "NCCL error"}

EXPORTS = ["mlp_default_tables", "mlp_create", "mlp_destroy", "mlp_last_error", "mlp_configure", "mlp_set_tables",
           "mlp_set_sequences", "mlp_set_shard", "mlp_posterior_all_pairs", "mlp_get_distances", "mlp_relax",
           "mlp_get_csr", "mlp_total_cells", "mlp_get_csr_bulk", "mlp_debug_pair_dense", "mlp_nccl_unique_id",
           "mlp_comm_init", "mlp_exchange", "mlp_last_stats", "mlp_qp_guide_tree", "mlp_shard_pairs", "mlp_csr_layout", "mlp_get_csr_raw",
           "mlp_alloc_pinned", "mlp_free_pinned", "mlp_viterbi_all_pairs", "mlp_cpnp_model_adjustment", "mlp_viterbi_all_pairs_ex",
           "mlp_cpnp_g_features", "mlp_qp_guide_tree_ex", "mlp_qp_finish_alignment_host", "mlp_qp_finish_alignment",
           "mlp_free_host", "mlp_get_csr_packed", "mlp_cpnp_guide_tree", "mlp_cpnp_finish_alignment_host",
           "mlp_cpnp_finish_alignment", "mlp_debug_glibc_rand", "mlp_column_scores", "mlp_exchange_begin", "mlp_exchange_end",
           "mlp_exchange_distances", "mlp_exchange_needed", "mlp_qp_guide_tree_device", "mlp_debug_set_distances", "mlp_debug_loc_counters", "mlp_stream_begin", "mlp_stream_end", "mlp_restrict_pairs", "mlp_shard_pairs_within", "mlp_set_digest", "mlp_get_csr_packed_begin", "mlp_get_csr_packed_end"]


class HmmTables(C.Structure):
    _fields_ = [("init", C.c_float * 5), ("trans", (C.c_float * 5) * 5), ("match", (C.c_float * 26) * 26),
                ("ins", C.c_float * 26), ("ltrans", (C.c_float * 3) * 3), ("rtrans", C.c_float * 2)]


class PartTables(C.Structure):
    _fields_ = [("sub", (C.c_double * 26) * 26), ("go", C.c_double), ("ge", C.c_double),
                ("tgo", C.c_double), ("tge", C.c_double)]


class StageStats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_kernel", C.c_double * 8), ("launches", C.c_int64),
                ("cells", C.c_int64), ("pairs", C.c_int64), ("nnz", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class MlpError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("mlprobs_b200: %s (%d)%s" % (ERRORS.get(code, "error"), code, (": " + detail) if detail else ""))


_lib = None


def load():
    """Load the C-ABI library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmlprobs_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.mlp_last_error.restype = C.c_char_p
        lib.mlp_last_error.argtypes = [C.c_void_p]
        lib.mlp_destroy.restype = None
        lib.mlp_destroy.argtypes = [C.c_void_p]
        lib.mlp_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.mlp_default_tables.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        lib.mlp_configure.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        lib.mlp_set_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_set_sequences.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p]
        lib.mlp_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.mlp_posterior_all_pairs.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_float]
        lib.mlp_get_distances.argtypes = [C.c_void_p, C.c_void_p]
        lib.mlp_relax.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float]
        lib.mlp_get_csr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        lib.mlp_total_cells.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        lib.mlp_get_csr_bulk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_debug_pair_dense.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        lib.mlp_nccl_unique_id.argtypes = [C.c_void_p]
        lib.mlp_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.mlp_exchange.argtypes = [C.c_void_p]
        lib.mlp_exchange_distances.argtypes = [C.c_void_p]
        lib.mlp_exchange_needed.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        lib.mlp_qp_guide_tree_device.argtypes = [C.c_void_p, C.c_float] + [C.c_void_p] * 5
        lib.mlp_stream_begin.argtypes = [C.c_void_p, C.c_int]
        lib.mlp_stream_end.argtypes = [C.c_void_p, C.c_void_p]
        lib.mlp_restrict_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        lib.mlp_set_digest.argtypes = [C.c_void_p, C.c_void_p]
        lib.mlp_last_stats.argtypes = [C.c_void_p, C.c_void_p]
        lib.mlp_csr_layout.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.mlp_get_csr_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_get_csr_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_get_csr_packed_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_get_csr_packed_end.argtypes = [C.c_void_p]
        lib.mlp_qp_guide_tree_ex.argtypes = [C.c_int] + [C.c_void_p] * 6
        lib.mlp_alloc_pinned.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
        lib.mlp_free_pinned.argtypes = [C.c_void_p]
        lib.mlp_free_pinned.restype = None
        lib.mlp_viterbi_all_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_viterbi_all_pairs_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mlp_cpnp_g_features.argtypes = [C.c_int, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_float, C.c_char_p, C.c_int]
        lib.mlp_cpnp_model_adjustment.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.mlp_qp_guide_tree.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def default_tables(flavour, init_distrib2=0.700645):
    h, p = HmmTables(), PartTables()
    rc = load().mlp_default_tables(flavour, C.c_float(init_distrib2), C.byref(h), C.byref(p))
    if rc:
        raise MlpError(rc)
    return h, p


def cpnp_model_adjustment(n_identical, align_len):
    """-> (variance_mean, identity, sigma, init_distrib2) exactly as MSA::ModelAdjustmentTest derives them."""
    ni = np.ascontiguousarray(n_identical, np.int32); al = np.ascontiguousarray(align_len, np.int32)
    ident = C.c_float(0); sig = C.c_float(0); i2 = C.c_float(0)
    vm = load().mlp_cpnp_model_adjustment(len(ni), ni.ctypes.data_as(C.c_void_p), al.ctypes.data_as(C.c_void_p),
                                          C.byref(ident), C.byref(sig), C.byref(i2))
    if vm < 0:
        raise MlpError(vm)
    return vm, ident.value, sig.value, i2.value


def cpnp_g_features(seqs, aln, aln_off, theta=1.0):
    """The `c_p_np_aln -G` line from the Viterbi alignments (host only)."""
    lens = np.array([len(s) for s in seqs], np.int32)
    line = C.create_string_buffer(256)
    rc = load().mlp_cpnp_g_features(len(seqs), lens.ctypes.data_as(C.c_void_p), b"".join(seqs), aln.ctypes.data_as(C.c_void_p),
                                    aln_off.ctypes.data_as(C.c_void_p), C.c_float(theta), line, 256)
    if rc:
        raise MlpError(rc)
    return line.value


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    rc = load().mlp_nccl_unique_id(buf)
    if rc:
        raise MlpError(rc)
    return bytes(buf)


def qp_guide_tree(distances):
    """UPGMA tree -> (weights, subtree distances, parent array, distances after the in-place update)."""
    d = np.array(distances, np.float32, copy=True, order="C")
    n = d.shape[0]
    w = np.zeros(n, np.float32)
    sd = np.zeros((n, n), np.float32)
    par = np.zeros(2 * n - 1, np.int32)
    rc = load().mlp_qp_guide_tree(n, d.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                  sd.ctypes.data_as(C.c_void_p), par.ctypes.data_as(C.c_void_p))
    if rc:
        raise MlpError(rc)
    return w, sd, par, d


def qp_guide_tree_ex(distances):
    """UPGMA tree -> dict(weights, seldist, parent, left, right, dist_after); children as mlp_qp_finish_alignment* wants them."""
    d = np.array(distances, np.float32, copy=True, order="C")
    n = d.shape[0]
    w = np.zeros(n, np.float32)
    sd = np.zeros((n, n), np.float32)
    par, left, right = (np.zeros(2 * n - 1, np.int32) for _ in range(3))
    rc = load().mlp_qp_guide_tree_ex(n, _ptr(d), _ptr(w), _ptr(sd), _ptr(par), _ptr(left), _ptr(right))
    if rc:
        raise MlpError(rc)
    return {"weights": w, "seldist": sd, "parent": par, "left": left, "right": right, "dist_after": d}


def _take_rows(n, rows_p, alen):
    lib = load()
    lib.mlp_free_host.argtypes = [C.c_void_p]
    lib.mlp_free_host.restype = None
    raw = C.string_at(rows_p, n * alen.value)
    lib.mlp_free_host(rows_p)
    L = alen.value
    return [raw[i * L:(i + 1) * L] for i in range(n)]


def qp_finish_alignment_host(seqs, weights, left, right, rp_off, nz_off, rp_pool, cells, ref_iters=-1, ref_seed=0):
    """Progressive construction + column refinement from a HOST copy of the pooled sparse set (no GPU work).
    cells: structured/2-column array of {int32 column, float32 value}. Returns the aligned rows (bytes) in input order."""
    n = len(seqs)
    lens = np.array([len(x) for x in seqs], np.int32)
    cat = np.frombuffer(b"".join(seqs), np.uint8)
    cv = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
    keep = [cv(weights, np.float32), cv(left, np.int32), cv(right, np.int32), cv(rp_off, np.int64), cv(nz_off, np.int64),
            cv(rp_pool, np.int32), None if cells is None else np.ascontiguousarray(cells)]
    rows_p = C.c_void_p(0)
    alen = C.c_int32(0)
    lib = load()
    lib.mlp_qp_finish_alignment_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_void_p] * 7 + [C.c_int, C.c_uint32,
                                                 C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    rc = lib.mlp_qp_finish_alignment_host(n, _ptr(lens), _ptr(cat), *[_ptr(k) for k in keep], int(ref_iters), int(ref_seed),
                                          C.byref(rows_p), C.byref(alen))
    if rc:
        raise MlpError(rc)
    return _take_rows(n, rows_p, alen)


def cpnp_guide_tree(distances, variance_id):
    """c_p_np_aln's UPGMA (MSAClusterTree::create(vpid)) -> dict(weights int32, left, right, dist_after)."""
    d = np.array(distances, np.float32, copy=True, order="C")
    n = d.shape[0]
    w = np.zeros(n, np.int32)
    left, right = (np.zeros(2 * n - 1, np.int32) for _ in range(2))
    lib = load()
    lib.mlp_cpnp_guide_tree.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.mlp_cpnp_guide_tree(n, _ptr(d), int(variance_id), _ptr(w), _ptr(left), _ptr(right))
    if rc:
        raise MlpError(rc)
    return {"weights": w, "left": left, "right": right, "dist_after": d}


def cpnp_finish_alignment_host(seqs, iweights, left, right, rp_off, nz_off, rp_pool, cells, refine_reps=100, pid=0):
    """c_p_np_aln -p 0 tail from a HOST copy of the pooled set -> (rows in the reference's output order, order of input indices)."""
    n = len(seqs)
    lens = np.array([len(x) for x in seqs], np.int32)
    cat = np.frombuffer(b"".join(seqs), np.uint8)
    keep = [np.ascontiguousarray(iweights, np.int32), np.ascontiguousarray(left, np.int32), np.ascontiguousarray(right, np.int32),
            np.ascontiguousarray(rp_off, np.int64), np.ascontiguousarray(nz_off, np.int64), np.ascontiguousarray(rp_pool, np.int32),
            np.ascontiguousarray(cells)]
    rows_p = C.c_void_p(0)
    alen = C.c_int32(0)
    order = np.zeros(n, np.int32)
    lib = load()
    lib.mlp_cpnp_finish_alignment_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_void_p] * 7 + [C.c_int, C.c_int,
                                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_void_p]
    rc = lib.mlp_cpnp_finish_alignment_host(n, _ptr(lens), _ptr(cat), *[_ptr(k) for k in keep], int(refine_reps), int(pid),
                                            C.byref(rows_p), C.byref(alen), _ptr(order))
    if rc:
        raise MlpError(rc)
    return _take_rows(n, rows_p, alen), order


def cpnp_np_finish_alignment_host(seqs, distances, rp_off, nz_off, rp_pool, cells, refine_reps=100, seed=-1):
    """c_p_np_aln -p 1 tail (alignment graph + similar-set refinement) from a HOST copy of the pooled set -> rows in input order.
    seed < 0 reseeds from the wall clock before every sweep like the reference; seed >= 0 pins it."""
    n = len(seqs)
    lens = np.array([len(x) for x in seqs], np.int32)
    cat = np.frombuffer(b"".join(seqs), np.uint8)
    keep = [np.ascontiguousarray(distances, np.float32), np.ascontiguousarray(rp_off, np.int64), np.ascontiguousarray(nz_off, np.int64),
            np.ascontiguousarray(rp_pool, np.int32), np.ascontiguousarray(cells)]
    rows_p = C.c_void_p(0)
    alen = C.c_int32(0)
    lib = load()
    lib.mlp_cpnp_np_finish_alignment_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_void_p] * 5 + [C.c_int, C.c_int64,
                                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    rc = lib.mlp_cpnp_np_finish_alignment_host(n, _ptr(lens), _ptr(cat), *[_ptr(k) for k in keep], int(refine_reps), int(seed),
                                               C.byref(rows_p), C.byref(alen))
    if rc:
        raise MlpError(rc)
    return _take_rows(n, rows_p, alen)


def column_scores(rows):
    """calculateColScore of MLProbs' Python driver for an alignment given as equal-length byte rows ->
    (col_score float64 array, mean, sd, peak_length_ratio)."""
    n = len(rows)
    L = len(rows[0]) if n else 0
    buf = np.frombuffer(b"".join(rows), np.uint8)
    out = np.zeros(L, np.float64)
    mean, sd, ratio = C.c_double(0), C.c_double(0), C.c_double(0)
    lib = load()
    lib.mlp_column_scores.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rc = lib.mlp_column_scores(n, L, _ptr(buf), _ptr(out), C.byref(mean), C.byref(sd), C.byref(ratio))
    if rc:
        raise MlpError(rc)
    return out, mean.value, sd.value, ratio.value


def debug_glibc_rand(count):
    out = np.zeros(count, np.int32)
    lib = load()
    lib.mlp_debug_glibc_rand.argtypes = [C.c_int, C.c_void_p]
    rc = lib.mlp_debug_glibc_rand(count, _ptr(out))
    if rc:
        raise MlpError(rc)
    return out


def debug_glibc_rand_seeded(seed, count):
    out = np.zeros(count, np.int32)
    lib = load()
    lib.mlp_debug_glibc_rand_seeded.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    rc = lib.mlp_debug_glibc_rand_seeded(int(seed) & 0xffffffff, count, _ptr(out))
    if rc:
        raise MlpError(rc)
    return out


def shard_pairs(lens, rank, world):
    """Pairs (a,b) owned by `rank` of `world`, in device processing order (host-only, no GPU needed)."""
    lens = np.ascontiguousarray(lens, np.int32)
    lib = load()
    lib.mlp_shard_pairs.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
    cnt = C.c_int64(0)
    rc = lib.mlp_shard_pairs(len(lens), lens.ctypes.data_as(C.c_void_p), rank, world, None, C.byref(cnt))
    if rc:
        raise MlpError(rc)
    out = np.zeros((cnt.value, 2), np.int32)
    rc = lib.mlp_shard_pairs(len(lens), lens.ctypes.data_as(C.c_void_p), rank, world, out.ctypes.data_as(C.c_void_p), C.byref(cnt))
    if rc:
        raise MlpError(rc)
    return out


def shard_pairs_within(lens, rank, world, seldist, selectivity=200.0):
    """The pairs of shard_pairs(rank, world) whose subtree-size distance is within the selectivity -- what mlp_restrict_pairs keeps
    (host-only, no GPU needed)."""
    lens = np.ascontiguousarray(lens, np.int32)
    sd = np.ascontiguousarray(seldist, np.float32)
    lib = load()
    lib.mlp_shard_pairs_within.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(C.c_int64)]
    cnt = C.c_int64(0)
    rc = lib.mlp_shard_pairs_within(len(lens), lens.ctypes.data_as(C.c_void_p), rank, world, sd.ctypes.data_as(C.c_void_p), C.c_float(selectivity), None, C.byref(cnt))
    if rc:
        raise MlpError(rc)
    out = np.zeros((cnt.value, 2), np.int32)
    rc = lib.mlp_shard_pairs_within(len(lens), lens.ctypes.data_as(C.c_void_p), rank, world, sd.ctypes.data_as(C.c_void_p), C.c_float(selectivity),
                                    out.ctypes.data_as(C.c_void_p), C.byref(cnt))
    if rc:
        raise MlpError(rc)
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PinnedCsrBuffers:
    """Page-locked host buffers (mlp_alloc_pinned) for Engine.csr_raw; grown on demand, freed on close()."""

    def __init__(self, n, rp_total, cells):
        self._ptrs = []
        self.cap_rp = self.cap_cells = self.n = 0
        self.ensure(n, rp_total, cells)

    def _pinned(self, nbytes, dtype, count):
        p = C.c_void_p()
        rc = load().mlp_alloc_pinned(max(int(nbytes), 16), C.byref(p))
        if rc:
            raise MlpError(rc, "pinned host allocation failed")
        self._ptrs.append(p)
        buf = (C.c_char * max(int(nbytes), 16)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def ensure(self, n, rp_total, cells):
        if n != self.n:
            self.nz_off = self._pinned(n * n * 8, np.int64, n * n)
            self.nz_cnt = self._pinned(n * n * 4, np.int32, n * n)
            self.n = n
        if rp_total > self.cap_rp:
            self.cap_rp = int(rp_total * 1.05) + 16
            self.rp_pool = self._pinned(self.cap_rp * 4, np.int32, self.cap_rp)
        if cells > self.cap_cells:
            self.cap_cells = int(cells * 1.1) + 1024
            self.cells = self._pinned(self.cap_cells * 8, np.dtype([("col", np.int32), ("val", np.float32)]), self.cap_cells)

    def matrix(self, a, b, lens):
        """(row_ptr, col, val) of ordered pair (a, b) as views into the pooled buffers."""
        s = a * self.n + b
        rp = self.rp_pool[self.rp_off[s]: self.rp_off[s] + int(lens[a]) + 2]
        c = self.cells[self.nz_off[s]: self.nz_off[s] + self.nz_cnt[s]]
        return rp, c["col"], c["val"]

    def nbytes(self):
        return self.n * self.n * 12 + self.rp_total * 4 + self.used * 8

    def close(self):
        for p in self._ptrs:
            load().mlp_free_pinned(p)
        self._ptrs = []


class PinnedPackedBuffers(PinnedCsrBuffers):
    """Page-locked host buffers for Engine.csr_packed: QuickProbs' own format (SparseEntry<uint16,uint16>: column in the low half, value code in the high half,
    uint16 row sizes) -- half the PCIe bytes of the {int32, float32} pool."""

    def ensure(self, n, rp_total, cells):
        if n != self.n:
            self.nz_off = self._pinned(n * n * 8, np.int64, n * n)
            self.nz_cnt = self._pinned(n * n * 4, np.int32, n * n)
            self.n = n
        if rp_total > self.cap_rp:
            self.cap_rp = int(rp_total * 1.05) + 16
            self.row_sizes = self._pinned(self.cap_rp * 2, np.uint16, self.cap_rp)
        if cells > self.cap_cells:
            self.cap_cells = int(cells * 1.1) + 1024
            self.cells = self._pinned(self.cap_cells * 4, np.uint32, self.cap_cells)

    def matrix(self, a, b, lens):
        """(row_ptr, col, val) of ordered pair (a, b), decoded (value = code / 65535 in float32, SparseEntry.h:31-32)."""
        s = a * self.n + b
        sizes = self.row_sizes[self.rp_off[s]: self.rp_off[s] + int(lens[a]) + 2].astype(np.int32)
        rp = np.zeros(int(lens[a]) + 2, np.int32)
        rp[1:] = np.cumsum(sizes)[:-1]
        c = self.cells[self.nz_off[s]: self.nz_off[s] + self.nz_cnt[s]]
        return rp, (c & 0xffff).astype(np.int32), (c >> 16).astype(np.float32) / np.float32(65535.0)

    def nbytes(self):
        return self.n * self.n * 12 + self.rp_total * 2 + self.used * 4


class Engine:
    """One GPU context (mlp_ctx)."""

    def __init__(self, device=0):
        self._lib = load()
        self._ctx = C.c_void_p()
        rc = self._lib.mlp_create(device, C.byref(self._ctx))
        if rc:
            raise MlpError(rc)
        self.n = 0
        self.lens = None

    def close(self):
        if self._ctx:
            self._lib.mlp_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise MlpError(rc, (self._lib.mlp_last_error(self._ctx) or b"").decode())

    def configure(self, scratch_bytes=0, cell_capacity=0):
        self._ck(self._lib.mlp_configure(self._ctx, scratch_bytes, cell_capacity))

    def set_tables(self, hmm, part):
        self._hmm, self._part = hmm, part
        self._ck(self._lib.mlp_set_tables(self._ctx, C.byref(hmm), C.byref(part)))

    def set_sequences(self, seqs):
        """seqs: list of bytes, upper-case letters."""
        self.lens = np.array([len(s) for s in seqs], np.int32)
        self.n = len(seqs)
        cat = b"".join(seqs)
        self._ck(self._lib.mlp_set_sequences(self._ctx, self.n, _ptr(self.lens), cat))

    def set_shard(self, rank, world):
        self._ck(self._lib.mlp_set_shard(self._ctx, rank, world))

    def posterior_all_pairs(self, flavour, mask, cutoff=0.01):
        self._ck(self._lib.mlp_posterior_all_pairs(self._ctx, flavour, mask, C.c_float(cutoff)))

    def viterbi_all_pairs(self):
        """(n_identical, align_len) per pair in row-major a<b order (ModelAdjustmentTest's pair loop)."""
        npairs = self.n * (self.n - 1) // 2
        ident = np.zeros(npairs, np.int32); ln = np.zeros(npairs, np.int32)
        self._ck(self._lib.mlp_viterbi_all_pairs(self._ctx, _ptr(ident), _ptr(ln)))
        return ident, ln

    def viterbi_alignments(self):
        """(n_identical, align_len, aln bytes, aln_off) -- the B/X/Y strings ComputeViterbiAlignment returns, for every pair."""
        npairs = self.n * (self.n - 1) // 2
        ident = np.zeros(npairs, np.int32); ln = np.zeros(npairs, np.int32)
        total = int(sum(int(self.lens[a]) + int(self.lens[b]) for a in range(self.n) for b in range(a + 1, self.n)))
        aln = np.zeros(total + 16, np.uint8); off = np.zeros(npairs + 1, np.int64)
        self._ck(self._lib.mlp_viterbi_all_pairs_ex(self._ctx, _ptr(ident), _ptr(ln), _ptr(aln), _ptr(off)))
        return ident, ln, aln, off

    def distances(self):
        d = np.zeros((self.n, self.n), np.float32)
        self._ck(self._lib.mlp_get_distances(self._ctx, _ptr(d)))
        return d

    def relax(self, flavour, weights=None, seldist=None, selectivity=200.0, selfweight=3.0, cutoff=0.01):
        w = np.ascontiguousarray(weights, np.float32) if weights is not None else None
        sd = np.ascontiguousarray(seldist, np.float32) if seldist is not None else None
        self._ck(self._lib.mlp_relax(self._ctx, flavour, _ptr(w), _ptr(sd), C.c_float(selectivity),
                                     C.c_float(selfweight), C.c_float(cutoff)))

    def qp_guide_tree_device(self, min_weight=1e-6, want_seldist=False):
        """QuickProbs' guide tree built on the device from the resident distance matrix (mlp_qp_guide_tree_device): the weights
        (saturated at min_weight) and the selectivity distances stay on the device for relax(QP) with weights = seldist = None."""
        n = self.n
        w = np.zeros(n, np.float32)
        par = np.zeros(2 * n - 1, np.int32); left = np.zeros(2 * n - 1, np.int32); right = np.zeros(2 * n - 1, np.int32)
        sd = np.zeros((n, n), np.float32) if want_seldist else None
        self._ck(self._lib.mlp_qp_guide_tree_device(self._ctx, C.c_float(min_weight), _ptr(w), _ptr(par), _ptr(left), _ptr(right), _ptr(sd)))
        return {"weights": w, "parent": par, "left": left, "right": right, "seldist": sd}

    def stream_begin(self, reps=1):
        """Streamed posterior stage (mlp_stream_begin): batches are finished, digested and dropped."""
        self._ck(self._lib.mlp_stream_begin(self._ctx, int(reps)))

    def stream_end(self, want_digest=True):
        out = np.zeros((self.n, self.n), np.uint64) if want_digest else None
        self._ck(self._lib.mlp_stream_end(self._ctx, _ptr(out)))
        return out

    def restrict_pairs(self, seldist, selectivity=200.0):
        sd = np.ascontiguousarray(seldist, np.float32)
        self._ck(self._lib.mlp_restrict_pairs(self._ctx, _ptr(sd), C.c_float(selectivity)))

    def debug_set_distances(self, d):
        d = np.ascontiguousarray(d, np.float32)
        assert d.shape == (self.n, self.n)
        self._ck(self._lib.mlp_debug_set_distances(self._ctx, _ptr(d)))

    def comm_init(self, id128: bytes, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        self._ck(self._lib.mlp_comm_init(self._ctx, buf, rank, world))

    def exchange(self):
        """All-gather the sparse posteriors + distances of every rank's shard (NCCL)."""
        self._ck(self._lib.mlp_exchange(self._ctx))

    def exchange_distances(self):
        """All-reduce of the distance matrix only (the guide tree needs all of it)."""
        self._ck(self._lib.mlp_exchange_distances(self._ctx))

    def exchange_needed(self, seldist, selectivity=200.0):
        """Import the matrices QuickProbs' consistency can read (seldist <= selectivity) from the ranks that own them."""
        sd = np.ascontiguousarray(seldist, np.float32)
        self._ck(self._lib.mlp_exchange_needed(self._ctx, _ptr(sd), C.c_float(selectivity)))

    def set_digest(self):
        """Per-matrix 64-bit digests (n*n, zeros for matrices other ranks own), computed on the device."""
        out = np.zeros(self.n * self.n, np.uint64)
        self._ck(self._lib.mlp_set_digest(self._ctx, _ptr(out)))
        return out

    def exchange_begin(self):
        """Start the exchange; distances() then only waits for the distance all-reduce (the tree overlaps the cell broadcasts)."""
        self._lib.mlp_exchange_begin.argtypes = [C.c_void_p]
        self._ck(self._lib.mlp_exchange_begin(self._ctx))

    def exchange_end(self):
        self._lib.mlp_exchange_end.argtypes = [C.c_void_p]
        self._ck(self._lib.mlp_exchange_end(self._ctx))

    def csr(self, a, b):
        nnz = C.c_int64(0)
        rp = np.zeros(int(self.lens[a]) + 2, np.int32)
        self._ck(self._lib.mlp_get_csr(self._ctx, a, b, _ptr(rp), None, None, C.byref(nnz)))
        col = np.zeros(nnz.value, np.int32)
        val = np.zeros(nnz.value, np.float32)
        if nnz.value:
            self._ck(self._lib.mlp_get_csr(self._ctx, a, b, None, _ptr(col), _ptr(val), C.byref(nnz)))
        return rp, col, val

    def csr_bulk(self):
        npairs = self.n * (self.n - 1) // 2
        nnz = np.zeros(npairs, np.int64)
        self._ck(self._lib.mlp_get_csr_bulk(self._ctx, _ptr(nnz), None, None, None))
        rows = int(sum(int(self.lens[a]) + 2 for a in range(self.n) for _ in range(a + 1, self.n)))
        rp = np.zeros(rows, np.int32)
        col = np.zeros(int(nnz.sum()), np.int32)
        val = np.zeros(int(nnz.sum()), np.float32)
        self._ck(self._lib.mlp_get_csr_bulk(self._ctx, None, _ptr(rp), _ptr(col), _ptr(val)))
        return nnz, rp, col, val

    def csr_layout(self):
        rp_off = np.zeros(self.n * self.n, np.int64)
        rp_total = C.c_int64(0); used = C.c_int64(0)
        self._ck(self._lib.mlp_csr_layout(self._ctx, _ptr(rp_off), C.byref(rp_total), C.byref(used)))
        return rp_off, rp_total.value, used.value

    def csr_raw(self, out=None):
        """Pooled read-back (no per-pair reshuffle). `out` = PinnedCsrBuffers to reuse page-locked host memory."""
        rp_off, rp_total, used = self.csr_layout()
        if out is None:
            out = PinnedCsrBuffers(self.n, rp_total, used)
        out.ensure(self.n, rp_total, used)
        self._ck(self._lib.mlp_get_csr_raw(self._ctx, _ptr(out.nz_off), _ptr(out.nz_cnt), _ptr(out.rp_pool), _ptr(out.cells)))
        out.rp_off, out.rp_total, out.used = rp_off, rp_total, used
        return out

    def qp_finish_alignment(self, weights, left, right, ref_iters=-1, ref_seed=0):
        """Progressive construction + column refinement over the sparse set resident on the device -> aligned rows (bytes)."""
        keep = [np.ascontiguousarray(weights, np.float32), np.ascontiguousarray(left, np.int32), np.ascontiguousarray(right, np.int32)]
        rows_p = C.c_void_p(0)
        alen = C.c_int32(0)
        self._lib.mlp_qp_finish_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
        self._ck(self._lib.mlp_qp_finish_alignment(self._ctx, *[_ptr(k) for k in keep], int(ref_iters), int(ref_seed),
                                                   C.byref(rows_p), C.byref(alen)))
        return _take_rows(self.n, rows_p, alen)

    def cpnp_finish_alignment(self, iweights, left, right, refine_reps=100, pid=0):
        """c_p_np_aln -p 0 tail over the device-resident set -> (rows in the reference's output order, input index of each row)."""
        keep = [np.ascontiguousarray(iweights, np.int32), np.ascontiguousarray(left, np.int32), np.ascontiguousarray(right, np.int32)]
        rows_p = C.c_void_p(0)
        alen = C.c_int32(0)
        order = np.zeros(self.n, np.int32)
        self._lib.mlp_cpnp_finish_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                        C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_void_p]
        self._ck(self._lib.mlp_cpnp_finish_alignment(self._ctx, *[_ptr(k) for k in keep], int(refine_reps), int(pid),
                                                     C.byref(rows_p), C.byref(alen), _ptr(order)))
        return _take_rows(self.n, rows_p, alen), order

    def cpnp_np_finish_alignment(self, refine_reps=100, seed=-1):
        """c_p_np_aln -p 1 tail on the resident set (graph on the host from a read-back, refinement on the device)."""
        rows_p = C.c_void_p(0)
        alen = C.c_int32(0)
        self._lib.mlp_cpnp_np_finish_alignment.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
        self._ck(self._lib.mlp_cpnp_np_finish_alignment(self._ctx, int(refine_reps), int(seed), C.byref(rows_p), C.byref(alen)))
        return _take_rows(self.n, rows_p, alen)

    def csr_packed(self, out=None):
        """Pooled read-back in QuickProbs' packed cell format (QP flavour only). `out` = PinnedPackedBuffers to reuse."""
        rp_off, rp_total, used = self.csr_layout()
        if out is None:
            out = PinnedPackedBuffers(self.n, rp_total, used)
        out.ensure(self.n, rp_total, used)
        self._ck(self._lib.mlp_get_csr_packed(self._ctx, _ptr(out.nz_off), _ptr(out.nz_cnt), _ptr(out.row_sizes), _ptr(out.cells)))
        out.rp_off, out.rp_total, out.used = rp_off, rp_total, used
        return out

    def csr_packed_begin(self, out):
        """Start the packed read-back into `out` (PinnedPackedBuffers); csr_packed_end() waits for it.  The next family's posterior
        stage may run in between."""
        key = self.lens.tobytes()
        if getattr(self, "_layout_key", None) != key:          # the row-pointer layout depends on the lengths only
            self._layout_rp_off, self._layout_rp_total, _ = self.csr_layout()
            self._layout_key = key
        rp_total = C.c_int64(0); used = C.c_int64(0)
        self._ck(self._lib.mlp_csr_layout(self._ctx, None, C.byref(rp_total), C.byref(used)))
        rp_off, rp_total, used = self._layout_rp_off, rp_total.value, used.value
        out.ensure(self.n, rp_total, used)
        self._ck(self._lib.mlp_get_csr_packed_begin(self._ctx, _ptr(out.nz_off), _ptr(out.nz_cnt), _ptr(out.row_sizes), _ptr(out.cells)))
        out.rp_off, out.rp_total, out.used = rp_off, rp_total, used
        return out

    def csr_packed_end(self):
        self._ck(self._lib.mlp_get_csr_packed_end(self._ctx))

    def total_cells(self):
        c = C.c_int64(0)
        self._ck(self._lib.mlp_total_cells(self._ctx, C.byref(c)))
        return c.value

    def debug_pair_dense(self, flavour, mask, a, b):
        L1, L2 = int(self.lens[a]), int(self.lens[b])
        outs = [np.zeros((L1 + 1, L2 + 1), np.float32) for _ in range(4)]
        dist = C.c_float(0)
        self._ck(self._lib.mlp_debug_pair_dense(self._ctx, flavour, mask, a, b, _ptr(outs[0]), _ptr(outs[1]),
                                                _ptr(outs[2]), _ptr(outs[3]), C.byref(dist)))
        return {"merged": outs[0], "hmm5": outs[1], "part": outs[2], "local": outs[3], "dist": dist.value}

    def stats(self):
        s = StageStats()
        self._ck(self._lib.mlp_last_stats(self._ctx, C.byref(s)))
        return {"ms_total": s.ms_total, "ms_kernel": {K_NAMES[i]: s.ms_kernel[i] for i in range(8)},
                "launches": s.launches, "cells": s.cells, "pairs": s.pairs, "nnz": s.nnz,
                "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes}
