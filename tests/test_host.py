"""CPU: host-side logic of the product library and the C-ABI surface (no compute kernels are launched)."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import mlprobs_b200 as M
from mlprobs_b200 import _capi
import oracle_lib as O
from common import load_golden, HERE

ROOT = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mlprobs_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mlp_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    lib = C.CDLL(M.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in include/mlprobs_b200.h is not exported" % name
    assert set(_capi.EXPORTS) <= set(declared)


def test_default_tables_equal_oracle_tables():
    for flav, oflav, i2 in [(M.QP, O.QP, 0.700645), (M.CPNP_P0, O.CPNP_P0, 0.170705), (M.CPNP_P0, O.CPNP_P0, 0.100675)]:
        h, p = M.default_tables(flav, i2)
        oh = O.hmm_tables(i2 if flav != M.QP else 0.700645); op = O.part_tables(oflav)
        assert bytes(h) == bytes(oh)
        assert bytes(p) == bytes(op)


def test_default_tables_equal_reference_dump():
    d = load_golden("qp_sup139")
    h, p = M.default_tables(M.QP)
    np.testing.assert_array_equal(np.ctypeslib.as_array(h.match), d["hmm.match"])
    np.testing.assert_array_equal(np.ctypeslib.as_array(p.sub), d["part.sub"])


@pytest.mark.parametrize("name", ["qp_sup139", "qp_sup002", "qp_676s4", "qp_75t2"])
def test_guide_tree_weights_and_subtree_distances(name):
    d = load_golden(name)
    w, sd, par, after = M.qp_guide_tree(d["distances"])
    np.testing.assert_array_equal(np.maximum(w, np.float32(1e-6)), d["weights"])   # saturation, ExtendedMSA.cpp:169
    np.testing.assert_array_equal(sd, d["seldist"])
    np.testing.assert_array_equal(after, d["distances_after_tree"])
    assert (par >= -1).all() and (par == -1).sum() == 1


def _upgma_bruteforce(dist):
    """Literal restatement of the reference scan (ClusterTree.cpp:69-120): i ascending, j ascending, strict '<'."""
    d = dist.astype(np.float32).copy(); n = d.shape[0]
    alive = list(range(n)); node_of = list(range(n)); leaves = [1] * n + [0] * (n - 1)
    parent = [-1] * (2 * n - 1)
    for node in range(n, 2 * n - 1):
        best = np.float32(2.0); bi = bj = -1
        for i in alive:
            for j in alive:
                if j >= i:
                    break
                if d[i, j] < best:
                    best = d[i, j]; bi, bj = i, j
        ni, nj = node_of[bi], node_of[bj]
        parent[ni] = parent[nj] = node; leaves[node] = leaves[ni] + leaves[nj]
        alive.remove(bj)
        isz, jsz = np.float32(leaves[ni]), np.float32(leaves[nj])
        joins = {k: np.float32((d[bi, k] * isz + d[bj, k] * jsz) / (isz + jsz)) for k in alive}
        node_of[bi] = node
        for k in alive:
            d[bi, k] = d[k, bi] = joins[k]
    return np.array(parent, np.int32), d


def test_guide_tree_tie_breaking_matches_the_reference_scan_order():
    rng = np.random.default_rng(12)
    for trial in range(6):
        n = 37
        d = (rng.integers(1, 12, size=(n, n)) / np.float32(16)).astype(np.float32)   # coarse values: many exact ties
        d = np.maximum(d, d.T); np.fill_diagonal(d, 0)
        w, sd, par, after = M.qp_guide_tree(d)
        bpar, bafter = _upgma_bruteforce(d)
        np.testing.assert_array_equal(par, bpar)


def test_shard_pairs_partition_is_a_disjoint_cover():
    rng = np.random.default_rng(3)
    lens = rng.integers(20, 700, size=37)
    allp = set()
    sizes = []
    for r in range(4):
        p = M.shard_pairs(lens, r, 4)
        s = set(map(tuple, p.tolist()))
        assert not (s & allp)
        allp |= s
        sizes.append(sum(int(lens[a] + 1) * int(lens[b] + 1) for a, b in s))
    assert allp == {(a, b) for a in range(37) for b in range(a + 1, 37)}
    assert max(sizes) / min(sizes) < 1.15          # cost-balanced


@pytest.mark.parametrize("name", ["cpnp_sup002_ref", "cpnp_676s4_ref", "cpnp_sup139_mix"])
def test_model_adjustment_host_helper(name):
    d = load_golden(name)
    vm, ident, sig, i2 = M.cpnp_model_adjustment(d["vit.ident"], d["vit.len"])
    assert vm == int(d["variance_mean"][0])
    assert np.float32(i2) == d["initDistrib2"][0]
    assert (vm, ident, sig, i2) == O.model_adjustment(d["vit.ident"], d["vit.len"])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(M.MlpError) as e:
        M.Engine(0)
    assert e.value.code == -1


def test_argument_errors_are_reported_not_thrown():
    lib = M.load()
    assert lib.mlp_default_tables(9, C.c_float(0.5), None, None) == -3
    cnt = C.c_int64(0)
    lens = np.array([5], np.int32)
    lib.mlp_shard_pairs.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
    assert lib.mlp_shard_pairs(1, lens.ctypes.data_as(C.c_void_p), 0, 1, None, C.byref(cnt)) == -3


def test_column_scores_equal_the_reference_python_function():
    """mlp_column_scores vs utils/calculate_column_scores.py (fixture written by oracle/gen_golden_colscore.py): same doubles."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "colscore.npz"))
    for name in ("qp_sup139", "qp_sup002", "qp_676s4", "cpnp_676s4_ref"):
        d = load_golden(name)
        rows = [r.tobytes() for r in d["msa"]]
        col, mean, sd, ratio = M.column_scores(rows)
        assert np.array_equal(col, g[name + ".col"])
        assert [mean, sd, ratio] == g[name + ".stats"].tolist()


def test_suite_fixture_is_consistent():
    """tests/golden/suites: every family of the manifest has its input in one of the two archives (inputs_rest.tar.gz: the large
    families pinned in round 2) and a quickprobs digest; the round-1 families also have the c_p_np_aln one."""
    import json, tarfile
    base = os.path.join(os.path.dirname(__file__), "golden", "suites")
    man = json.load(open(os.path.join(base, "manifest.json")))["families"]
    names = set()
    for arc in ("inputs.tar.gz", "inputs_rest.tar.gz"):
        with tarfile.open(os.path.join(base, arc)) as t:
            names |= {m.name for m in t.getmembers() if m.isfile()}
    assert {m["suite"] + "/" + m["name"] for m in man} == names and len(man) > 1500
    assert all(m["qp_sha"] for m in man)
    assert all(m["cpnp_sha"] for m in man if not m.get("rest"))


def test_g_feature_line_host_code_on_suite_families():
    """`c_p_np_aln -G` (MSA::Alter_ModelAdjustmentTest, MSA.cpp:646-762) over a spread of bundled families: the product's host
    function mlp_cpnp_g_features, fed with the oracle's Viterbi alignments (what k_viterbi delivers on the device), prints the
    reference binary's line byte for byte (tests/golden/suites/manifest.json `cpnpG`, written by oracle/gen_suite_golden_G.py)."""
    import json, tarfile
    from common import HERE
    import oracle_lib as O
    suites = os.path.join(HERE, "golden", "suites")
    man = [m for m in json.load(open(os.path.join(suites, "manifest.json")))["families"] if m.get("cpnpG") and not m.get("rest")]   # the round-1 (small) families
    assert len(man) > 500
    texts = {}
    with tarfile.open(os.path.join(suites, "inputs.tar.gz")) as tar:
        for ti in tar.getmembers():
            if ti.isfile():
                texts[ti.name] = tar.extractfile(ti).read().decode()
    ht = O.hmm_tables()
    standard = set(b"ARNDCQEGHILKMFPSTWYV")
    checked = other = 0
    for m in sorted(man, key=lambda e: e["cpnp_s"])[::12]:                    # every 12th family, cheapest first
        seqs = []
        for rec in texts["%s/%s" % (m["suite"], m["name"])].split(">")[1:]:
            body = "".join(rec.split("\n")[1:])
            seqs.append("".join(ch for ch in body if ch.isalpha()).upper().encode())
        if len(seqs) < 2 or len(seqs) > 40:
            continue
        alns, offs = [], [0]
        for a in range(len(seqs)):
            for b in range(a + 1, len(seqs)):
                aln = O.viterbi(ht, seqs[a], seqs[b])[3]
                alns.append(aln); offs.append(offs[-1] + len(aln))
        line = M.cpnp_g_features(seqs, np.frombuffer(b"".join(alns), np.uint8), np.array(offs, np.int64)).decode()
        if all(set(s) <= standard for s in seqs):
            assert line == m["cpnpG"], (m["suite"], m["name"])
            checked += 1
        else:
            # B/J/O/U/X/Z: the reference adds out-of-bounds reads to fields 5 and 6 (see include/mlprobs_b200.h); the other
            # fields are exact, and fields 5-6 stay close whenever the reference's own number is sane
            got, want = line.split("\t"), m["cpnpG"].split("\t")
            assert [got[k] for k in (0, 1, 2, 3, 6)] == [want[k] for k in (0, 1, 2, 3, 6)], (m["suite"], m["name"])
            if abs(float(want[4])) < 100:
                assert abs(float(got[4]) - float(want[4])) <= 5e-3 * max(1.0, abs(float(want[4]))) and abs(float(got[5]) - float(want[5])) < 0.02
            other += 1
    assert checked >= 40 and other >= 3


def test_float_facts_the_device_shortcuts_rest_on():
    """Pure IEEE-754 facts behind two exact shortcuts of the CUDA path, checked in numpy float32:
    final_c.cu mode 7 treats a merged posterior p as 0 when sq = v5^2 + vp^2 + vl^2 < 2^-122 and the MEA score it is added to is
    >= 2^-30: then p = sqrt(sq / 3) < 2^-60 and p + s == s; part_sc.cu applies the scale shift with one multiply by a power of two,
    which is exact for normal results."""
    rng = np.random.default_rng(7)
    sq = (rng.random(20000).astype(np.float32) * np.float32(2.0 ** -122)).astype(np.float32)
    p = np.sqrt((sq / np.float32(3.0)).astype(np.float32)).astype(np.float32)
    assert float(p.max()) < 2.0 ** -60
    s = np.exp2(rng.uniform(-30, 8, 20000)).astype(np.float32)
    s[:4] = np.float32(2.0 ** -30)
    assert np.array_equal((p + s).astype(np.float32), s)
    q = rng.random(20000) * np.exp2(rng.integers(-200, 200, 20000).astype(np.float64))
    k = rng.integers(-300, 300, 20000)
    assert np.array_equal(q * np.exp2(k.astype(np.float64)), np.ldexp(q, k))
