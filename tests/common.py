"""Shared helpers for the test-suite (fixtures under tests/golden were written by oracle/gen_golden.py from the compiled reference)."""
import os
import zlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def split_seqs(d):
    res = d["residues"].tobytes()
    out, p = [], 0
    for L in d["lens"]:
        out.append(res[p:p + int(L)])
        p += int(L)
    return out


def pairs(n):
    return [(a, b) for a in range(n) for b in range(a + 1, n)]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def digest_of(getter, n, transposed=False):
    """getter(a, b) -> (rowptr, col, val). Returns (nnz, rowptr_crc, col_crc, val_crc) arrays in pair order."""
    P = pairs(n)
    nnz = np.zeros(len(P), np.int32)
    rc = np.zeros(len(P), np.uint32); cc = np.zeros(len(P), np.uint32); vc = np.zeros(len(P), np.uint32)
    for p, (a, b) in enumerate(P):
        rp, c, v = getter(b, a) if transposed else getter(a, b)
        nnz[p] = len(c); rc[p] = crc(rp.astype(np.int32)); cc[p] = crc(c.astype(np.int32)); vc[p] = crc(v.astype(np.float32))
    return nnz, rc, cc, vc


def assert_digest(d, tag, getter, n, transposed=False):
    nnz, rc, cc, vc = digest_of(getter, n, transposed)
    np.testing.assert_array_equal(nnz, d["digest.%s.nnz" % tag], err_msg="nnz per pair (%s)" % tag)
    np.testing.assert_array_equal(rc, d["digest.%s.rowptr_crc" % tag], err_msg="row pointers (%s)" % tag)
    np.testing.assert_array_equal(cc, d["digest.%s.col_crc" % tag], err_msg="column index sets (%s)" % tag)
    np.testing.assert_array_equal(vc, d["digest.%s.val_crc" % tag], err_msg="values (%s)" % tag)


def cpnp_mask(pid):
    pid = int(pid) % 10
    return 7 if pid <= 1 else (4 if pid == 2 else 2)


def tail_from_csrset(S, seqs, distances, ref_iters=-1):
    """QuickProbs tail on the host (mlp_qp_finish_alignment_host) from an oracle_lib.CsrSet: tree from `distances`
    (before the in-place update), final weights saturated at 1e-6 (ExtendedMSA.cpp:237-238)."""
    import mlprobs_b200 as M
    t = M.qp_guide_tree_ex(distances)
    w = np.maximum(t["weights"], np.float32(1e-6))
    cells = np.zeros(len(S.col), dtype=[("c", np.int32), ("v", np.float32)])
    cells["c"] = S.col
    cells["v"] = S.val
    return M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], S.rp_off, S.nz_off, S.rowptr, cells, ref_iters)


def cpnp_tail_from_csrset(S, seqs, distances, variance_mean, refine_reps=100):
    """c_p_np_aln -p 0 tail on the host (mlp_cpnp_guide_tree + mlp_cpnp_finish_alignment_host) from an oracle_lib.CsrSet."""
    import mlprobs_b200 as M
    t = M.cpnp_guide_tree(distances, int(variance_mean) // 10)
    cells = np.zeros(len(S.col), dtype=[("c", np.int32), ("v", np.float32)])
    cells["c"] = S.col
    cells["v"] = S.val
    return M.cpnp_finish_alignment_host(seqs, t["weights"], t["left"], t["right"], S.rp_off, S.nz_off, S.rowptr, cells,
                                        refine_reps, int(variance_mean) % 10)


def cpnp_p1_sparse_set(seqs, threads=4):
    """What `c_p_np_aln -p 1` holds before its alignment graph, computed with the ORACLE: Viterbi statistics -> model class and
    initDistrib[2] (MSA.cpp:775-882), all-pairs posteriors of that class with the -p 1 merge and distance
    (ArrangePosteriorProbs, MSA.cpp:1635-1766), two relaxations.  Returns (distances, CsrSet, variance_mean)."""
    import oracle_lib as O
    n = len(seqs)
    ht0 = O.hmm_tables()
    ids, lens = [], []
    for a, b in pairs(n):
        _, i, l, _ = O.viterbi(ht0, seqs[a], seqs[b])
        ids.append(i); lens.append(l)
    vm, _, _, i2 = O.model_adjustment(ids, lens)
    ht = O.hmm_tables(float(np.float32(i2))); pt = O.part_tables(O.CPNP_P0)
    dist, S, rc = O.posterior_stage(O.CPNP_P1, cpnp_mask(vm), ht, pt, seqs, threads=threads)
    assert rc == 0
    for _ in range(2):
        S = O.relax_cpnp(S, 0.01, threads=threads)
    return dist, S, vm


def cpnp_np_tail_from_csrset(S, seqs, distances, refine_reps=100, seed=-1):
    """c_p_np_aln -p 1 tail on the host (mlp_cpnp_np_finish_alignment_host) from an oracle_lib.CsrSet."""
    import mlprobs_b200 as M
    cells = np.zeros(len(S.col), dtype=[("c", np.int32), ("v", np.float32)])
    cells["c"] = S.col
    cells["v"] = S.val
    return M.cpnp_np_finish_alignment_host(seqs, distances, S.rp_off, S.nz_off, S.rowptr, cells, refine_reps, seed)
