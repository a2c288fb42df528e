"""QuickProbs tail (progressive construction + column refinement) on the host, fed with the REFERENCE's own final sparse
matrices from the full fixtures: checks mlp_qp_guide_tree_ex + mlp_qp_finish_alignment_host against the reference's
alignments with nothing of the oracle in between (ConstructionStage.cpp:51-127, ColumnRefinement.cpp, RefinementBase.cpp)."""
import numpy as np
import pytest
import mlprobs_b200 as M
from common import load_golden, split_seqs, pairs


def pooled_from_fixture(d, n, lens):
    rp_off = np.zeros(n * n, np.int64); nz_off = np.zeros(n * n, np.int64)
    rps, cols, vals = [], [], []
    rp_at = 0; nz_at = 0
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            tag = "pair.%d.%d.sF" % (a, b) if a < b else "pair.%d.%d.tF" % (b, a)
            rp = d[tag + ".rowptr"].astype(np.int32); c = d[tag + ".col"].astype(np.int32); v = d[tag + ".val"].astype(np.float32)
            assert len(rp) == lens[a] + 2
            rp_off[a * n + b] = rp_at; nz_off[a * n + b] = nz_at
            rps.append(rp); cols.append(c); vals.append(v)
            rp_at += len(rp); nz_at += len(c)
    cells = np.zeros(nz_at, dtype=[("c", np.int32), ("v", np.float32)])
    cells["c"] = np.concatenate(cols); cells["v"] = np.concatenate(vals)
    return rp_off, nz_off, np.concatenate(rps), cells


@pytest.mark.parametrize("name", ["qp_sup139", "qp_sup002"])
def test_tail_from_reference_matrices(name):
    d = load_golden(name)
    seqs = split_seqs(d); n = len(seqs)
    lens = [len(s) for s in seqs]
    t = M.qp_guide_tree_ex(d["distances"])
    np.testing.assert_array_equal(t["weights"], np.asarray(d["weights"]) if np.all(d["weights"] > 1e-6) else t["weights"])
    w = np.maximum(t["weights"], np.float32(1e-6))
    rp_off, nz_off, rp_pool, cells = pooled_from_fixture(d, n, lens)
    for key, it in (("msa_construct", -2), ("msa", -1)):
        rows = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], rp_off, nz_off, rp_pool, cells, it)
        assert rows == [r.tobytes() for r in d[key]], key
        for r, s in zip(rows, seqs):
            assert r.replace(b"-", b"") == s


def test_tree_children_are_consistent_with_parents():
    d = load_golden("qp_sup002")
    t = M.qp_guide_tree_ex(d["distances"])
    n = len(d["lens"])
    par, left, right = t["parent"], t["left"], t["right"]
    assert par[2 * n - 2] == -1 and np.all(left[:n] == -1) and np.all(right[:n] == -1)
    for v in range(n, 2 * n - 1):
        assert par[left[v]] == v and par[right[v]] == v and left[v] < v and right[v] < v and left[v] != right[v]
    w2, sd2, par2, _ = M.qp_guide_tree(d["distances"])
    np.testing.assert_array_equal(par2, par); np.testing.assert_array_equal(w2, t["weights"])


def test_single_sequence_and_refinement_options():
    rows = M.qp_finish_alignment_host([b"ACDEFG"], None, None, None, None, None, None, None)
    assert rows == [b"ACDEFG"]
    d = load_golden("qp_sup139")
    seqs = split_seqs(d); n = len(seqs)
    t = M.qp_guide_tree_ex(d["distances"])
    w = np.maximum(t["weights"], np.float32(1e-6))
    rp_off, nz_off, rp_pool, cells = pooled_from_fixture(d, n, [len(s) for s in seqs])
    a = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], rp_off, nz_off, rp_pool, cells, 5, 0)
    b = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], rp_off, nz_off, rp_pool, cells, 5, 12345)
    for rows in (a, b):
        assert len({len(r) for r in rows}) == 1
        for r, s in zip(rows, seqs):
            assert r.replace(b"-", b"") == s
    # `-r 0` means "reference default" there (RefinementBase.cpp:33): same result as -1
    assert M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], rp_off, nz_off, rp_pool, cells, 0) == \
        [r.tobytes() for r in d["msa"]]


def test_private_glibc_rand_replica_matches_libc():
    """DoIterativeRefinement (MSA.cpp:1545) draws from an unseeded rand(); the library carries its own copy of that stream."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    want = [libc.rand() for _ in range(5000)]
    assert M.debug_glibc_rand(5000).tolist() == want


@pytest.mark.parametrize("seed", [0, 1, 777, 1792000000, 2**31 - 1, 2**31 + 5, 2**32 - 1])
def test_private_glibc_rand_replica_matches_libc_after_srand(seed):
    """`c_p_np_aln -p 1` calls srand(time(0)) before every refinement sweep (MSA.cpp:1896)."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(ctypes.c_uint(seed))
    want = [libc.rand() for _ in range(2000)]
    assert M.debug_glibc_rand_seeded(seed, 2000).tolist() == want


def _quick_sort_reference(arr, ind, low, high):
    """AlignGraph::Quick_sort / Partition (AlignGraph.h:62-113), restated in Python for small inputs."""
    stack = [(low, high)]
    while stack:
        low, high = stack.pop()
        if low >= high:
            continue
        lo, hi = low, high
        pivot, ip = arr[lo], ind[lo]
        while hi > lo:
            while pivot <= arr[hi] and hi > lo:
                hi -= 1
            arr[lo], ind[lo] = arr[hi], ind[hi]
            while pivot >= arr[lo] and hi > lo:
                lo += 1
            arr[hi], ind[hi] = arr[lo], ind[lo]
        arr[lo], ind[lo] = pivot, ip
        stack.append((low, lo - 1)); stack.append((lo + 1, high))


def test_alignment_graph_sort_keeps_the_reference_order_of_ties():
    """The graph takes cells strongest first from an UNSTABLE quicksort, so equal posteriors come out in that algorithm's order;
    the production version runs large sub-ranges as OpenMP tasks and must give the same permutation."""
    import ctypes as C
    lib = M.load()
    lib.mlp_debug_reference_sort.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(3)
    def run(keys, tasks):
        out = np.zeros(len(keys), np.int32)
        assert lib.mlp_debug_reference_sort(len(keys), keys.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), tasks) == 0
        return out
    small = (rng.integers(0, 40, 3000) / 40).astype(np.float32)           # heavy ties
    arr, ind = small.tolist(), list(range(len(small)))
    _quick_sort_reference(arr, ind, 0, len(arr) - 1)
    assert run(small, 0).tolist() == ind and run(small, 1).tolist() == ind
    big = np.round(rng.random(600000), 3).astype(np.float32)              # 1000 distinct values: ties everywhere, task path
    a, b = run(big, 0), run(big, 1)
    assert np.array_equal(a, b) and np.all(np.diff(big[a]) >= 0) and sorted(a.tolist()) == list(range(len(big)))


def _fake_pooled_set(seqs, rng, per_row=3):
    """A pooled sparse set with `per_row` cells around the length-scaled diagonal of every ordered pair; values on a coarse grid so
    that many cells tie.  The pairs contradict each other on purpose: the graph has to refuse cells that would close a cycle."""
    n = len(seqs); lens = [len(s) for s in seqs]
    rp_off = np.zeros(n * n, np.int64); nz_off = np.zeros(n * n, np.int64)
    rps, cols, vals = [], [], []
    rp_at = nz_at = 0
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            la, lb = lens[a], lens[b]
            rp = np.zeros(la + 2, np.int32); c = []
            for i in range(1, la + 1):
                k = min(per_row, lb)
                j0 = max(1, min(lb - k + 1, int(i * lb / la) - 1))
                c += list(range(j0, j0 + k))
                rp[i + 1] = rp[i] + k
            rp_off[a * n + b] = rp_at; nz_off[a * n + b] = nz_at
            rps.append(rp); cols.append(np.array(c, np.int32)); vals.append((rng.integers(1, 21, len(c)) / 20).astype(np.float32))
            rp_at += len(rp); nz_at += len(c)
    cells = np.zeros(nz_at, dtype=[("c", np.int32), ("v", np.float32)])
    cells["c"] = np.concatenate(cols); cells["v"] = np.concatenate(vals)
    return rp_off, nz_off, np.concatenate(rps), cells


@pytest.mark.parametrize("n,length,seed", [(2, 1, 1), (3, 7, 2), (9, 60, 3), (25, 40, 4), (12, 300, 5)])
def test_alignment_graph_always_yields_a_valid_alignment(n, length, seed):
    """Whatever the cells say, AlignGraph's output is an alignment: equal-length rows that degap to the inputs, no all-gap column,
    and the same rows when the call is repeated (c_p_np_aln -p 1 without refinement is deterministic)."""
    from mlprobs_b200 import synth
    rng = np.random.default_rng(seed)
    seqs = [synth.family(1, max(1, length + int(rng.integers(-length // 3, length // 3 + 1))), seed=seed * 100 + k)[0] for k in range(n)]
    rp_off, nz_off, rp_pool, cells = _fake_pooled_set(seqs, rng)
    dist = rng.random((n, n)).astype(np.float32)
    rows = M.cpnp_np_finish_alignment_host(seqs, dist, rp_off, nz_off, rp_pool, cells, 0, 1)
    assert len({len(r) for r in rows}) == 1
    assert [r.replace(b"-", b"") for r in rows] == seqs
    assert all(any(r[c] != ord("-") for r in rows) for c in range(len(rows[0])))
    assert M.cpnp_np_finish_alignment_host(seqs, dist, rp_off, nz_off, rp_pool, cells, 0, 99) == rows
    # with refinement the result depends on the seed only through the visiting order; it stays a valid alignment
    ref = M.cpnp_np_finish_alignment_host(seqs, dist, rp_off, nz_off, rp_pool, cells, 20, 7)
    assert len({len(r) for r in ref}) == 1 and [r.replace(b"-", b"") for r in ref] == seqs
    assert M.cpnp_np_finish_alignment_host(seqs, dist, rp_off, nz_off, rp_pool, cells, 20, 7) == ref
