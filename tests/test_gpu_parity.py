"""GPU parity: the CUDA path (through the C ABI) against golden vectors of the compiled reference, against the
oracle on seeded synthetic inputs, and through size-independent properties at larger sizes.
Bit-exact for everything except cpnp's 80-bit partition function (FP64 on the device): tolerance 1e-5 relative
(BASELINE.json north_star), index sets must still be identical on the fixtures."""
import numpy as np
import pytest
import mlprobs_b200 as M
from mlprobs_b200 import synth
import oracle_lib as O
from common import load_golden, split_seqs, pairs, assert_digest, cpnp_mask

pytestmark = pytest.mark.gpu
REL_TOL_PARTITION = 1e-5


def engine(flavour, seqs, init2=0.700645):
    eng = M.Engine(0)
    h, p = M.default_tables(flavour, init2)
    eng.set_tables(h, p)
    eng.set_sequences(seqs)
    return eng


# ------------------------------------------------------------------ golden vectors (reference-produced)
@pytest.mark.parametrize("name", ["qp_sup139", "qp_sup002", "qp_676s4", "qp_75t2"])
def test_qp_against_reference_fixture(name):
    d = load_golden(name)
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    np.testing.assert_array_equal(eng.distances(), d["distances"])
    assert_digest(d, "s0", eng.csr, n)
    assert_digest(d, "t0", eng.csr, n, transposed=True)
    w, sd, par, _ = M.qp_guide_tree(eng.distances())
    w = np.maximum(w, np.float32(1e-6))
    iters = 1 if n > 50 else 2
    for it in range(iters):
        cutoff = float(np.float32(0.01)) if it < iters - 1 else float(np.float32(1e-5))
        eng.relax(M.QP, w, sd, 200.0, 3.0, cutoff)
    assert_digest(d, "sF", eng.csr, n)
    assert_digest(d, "tF", eng.csr, n, transposed=True)
    # the tail on the device-resident set: progressive construction, then column refinement (reference alignments)
    t = M.qp_guide_tree_ex(d["distances"])
    np.testing.assert_array_equal(t["weights"], M.qp_guide_tree(d["distances"])[0])
    for key, ref_iters in (("msa_construct", -2), ("msa", -1)):
        rows = eng.qp_finish_alignment(w, t["left"], t["right"], ref_iters)
        assert rows == [r.tobytes() for r in d[key]], key
    st = eng.stats()
    assert st["launches"] >= n - 1 and st["ms_total"] > 0
    # the host tail over the read-back pooled set gives the same alignment
    raw = eng.csr_raw()
    rows_h = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], raw.rp_off, raw.nz_off, raw.rp_pool, raw.cells)
    assert rows_h == rows
    raw.close()
    eng.close()


def test_qp_tail_long_rows_and_wide_profiles_vs_host():
    """Device profile posterior vs the host provider on a diffuse family (sparse rows longer than the staging cap)."""
    seqs = synth.family(24, 180, seed=77, p_sub=0.7)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    t = M.qp_guide_tree_ex(eng.distances())
    w = np.maximum(t["weights"], np.float32(1e-6))
    eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, 0.01)
    eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, float(np.float32(1e-5)))
    raw = eng.csr_raw()
    rp = raw.rp_pool[:raw.rp_total]
    assert int(np.max(np.diff(rp.astype(np.int64)))) > 8            # at least one sparse row beyond the staged 8 cells
    dev = eng.qp_finish_alignment(w, t["left"], t["right"], 12, 7)
    host = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], raw.rp_off, raw.nz_off, raw.rp_pool, raw.cells, 12, 7)
    assert dev == host
    for r, s in zip(dev, seqs):
        assert r.replace(b"-", b"") == s
    raw.close()
    eng.close()


def test_qp_tail_profiles_wider_than_the_device_wavefront_fall_back_to_the_host_dp():
    """Three unrelated sequences of 8300 residues: every profile is wider than the 8192 columns the MEA wavefront kernel
    handles, so the dense matrix comes back and the host walks it; device tail == host tail."""
    rng = np.random.default_rng(11)
    al = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    seqs = [al[rng.integers(0, 20, 8300)].tobytes() for _ in range(3)]
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    t = M.qp_guide_tree_ex(eng.distances())
    w = np.maximum(t["weights"], np.float32(1e-6))
    for it in range(2):
        eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, 0.01 if it == 0 else float(np.float32(1e-5)))
    dev = eng.qp_finish_alignment(w, t["left"], t["right"], 3)
    raw = eng.csr_raw()
    host = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], raw.rp_off, raw.nz_off, raw.rp_pool, raw.cells, 3)
    assert len(dev[0]) > 8192
    assert dev == host
    raw.close()
    eng.close()


def test_qp_dense_posteriors_fixture():
    d = load_golden("qp_sup139")
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.QP, seqs)
    for a, b in pairs(n):
        g = eng.debug_pair_dense(M.QP, 3, a, b)
        t = "pair.%d.%d" % (a, b)
        np.testing.assert_array_equal(g["hmm5"], d[t + ".post5"])
        np.testing.assert_array_equal(g["part"], d[t + ".postP"])
        np.testing.assert_array_equal(g["merged"], d[t + ".post"])
    eng.close()


@pytest.mark.parametrize("name", ["cpnp_sup139_local", "cpnp_sup139_mix", "cpnp_sup139_part", "cpnp_sup002_ref", "cpnp_676s4_ref"])
def test_cpnp_against_reference_fixture(name):
    d = load_golden(name)
    seqs = split_seqs(d); n = len(seqs)
    mask = cpnp_mask(d["pid"][0])
    eng = engine(M.CPNP_P0, seqs, float(d["initDistrib2"][0]))
    eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)
    if mask & 2:
        # FP64 partition function vs the reference's long double: values within 1e-5, index sets identical here
        np.testing.assert_allclose(eng.distances(), d["distances"], rtol=REL_TOL_PARTITION, atol=1e-6)
        np.testing.assert_array_equal([len(eng.csr(a, b)[1]) for a, b in pairs(n)], d["digest.s0.nnz"])
        from common import digest_of
        _, _, cc, _ = digest_of(eng.csr, n)
        np.testing.assert_array_equal(cc, d["digest.s0.col_crc"])
    else:
        np.testing.assert_array_equal(eng.distances(), d["distances"])
        assert_digest(d, "s0", eng.csr, n)
        for r in range(int(d["reps"][0])):
            eng.relax(M.CPNP_P0, cutoff=0.01)
            assert_digest(d, "s%d" % (r + 1), eng.csr, n)
    eng.close()


@pytest.mark.parametrize("name", ["cpnp_sup002_ref", "cpnp_676s4_ref"])
def test_cpnp_tail_against_reference_alignment(name):
    """`c_p_np_aln -p 0` to the end on the device-resident set: tree, weighted progressive alignment, random-bipartition
    refinement -> the reference's one-thread output, rows in its order (MSA.cpp:1369-1623)."""
    d = load_golden(name)
    seqs = split_seqs(d); n = len(seqs)
    vm = int(d["variance_mean"][0])
    eng = engine(M.CPNP_P0, seqs, float(d["initDistrib2"][0]))
    eng.posterior_all_pairs(M.CPNP_P0, cpnp_mask(vm % 10), 0.01)
    t = M.cpnp_guide_tree(eng.distances(), vm // 10)
    for r in range(2):
        eng.relax(M.CPNP_P0, cutoff=0.01)
    for key, ir in (("msa_ir0", 0), ("msa", 100)):
        rows, order = eng.cpnp_finish_alignment(t["weights"], t["left"], t["right"], ir, vm % 10)
        np.testing.assert_array_equal(order, d[key + "_order"], err_msg=key)
        assert rows == [r.tobytes() for r in d[key]], key
    raw = eng.csr_raw()
    rows_h, order_h = M.cpnp_finish_alignment_host(seqs, t["weights"], t["left"], t["right"], raw.rp_off, raw.nz_off, raw.rp_pool, raw.cells, 100, vm % 10)
    assert rows_h == rows and np.array_equal(order_h, order)
    raw.close()
    eng.close()


def test_cpnp_dense_models_fixture():
    d = load_golden("cpnp_sup139_mix")
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.CPNP_P0, seqs, float(d["initDistrib2"][0]))
    for a, b in pairs(n):
        g = eng.debug_pair_dense(M.CPNP_P0, 7, a, b)
        t = "pair.%d.%d" % (a, b)
        np.testing.assert_array_equal(g["hmm5"], d[t + ".post5"])
        np.testing.assert_array_equal(g["local"], d[t + ".postL"])
        ref = d[t + ".postP"]
        np.testing.assert_allclose(g["part"], ref, rtol=REL_TOL_PARTITION, atol=1e-30)
        np.testing.assert_allclose(g["merged"], d[t + ".post"], rtol=REL_TOL_PARTITION, atol=1e-30)
    eng.close()


# ------------------------------------------------------------------ oracle on seeded synthetic inputs
def _cmp_sets(eng, S, n):
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            rp, c, v = eng.csr(a, b)
            orp, oc, ov = S.get(a, b)
            np.testing.assert_array_equal(rp, orp)
            np.testing.assert_array_equal(c, oc)
            np.testing.assert_array_equal(v, ov)


@pytest.mark.parametrize("lens", [(1, 1, 2, 3), (31, 32, 33, 63, 64, 65), (510, 511, 512, 513), (40, 700, 1100)])
def test_qp_ragged_and_multiblock_vs_oracle(lens):
    seqs = [synth.family(1, L, seed=100 + L)[0][:L].ljust(L, b"A") for L in lens]
    n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    dist, S, _ = O.posterior_stage(O.QP, 3, O.hmm_tables(), O.part_tables(O.QP), seqs, threads=8)
    np.testing.assert_array_equal(eng.distances(), dist)
    _cmp_sets(eng, S, n)
    eng.close()


@pytest.mark.parametrize("mask", [1, 4])
def test_cpnp_models_vs_oracle_family(mask):
    seqs = synth.family(9, 150, seed=7)
    n = len(seqs)
    i2 = 0.100675
    eng = engine(M.CPNP_P0, seqs, i2)
    eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)
    dist, S, _ = O.posterior_stage(O.CPNP_P0, mask, O.hmm_tables(i2), O.part_tables(O.CPNP_P0), seqs, threads=8)
    np.testing.assert_array_equal(eng.distances(), dist)
    _cmp_sets(eng, S, n)
    for _ in range(2):
        eng.relax(M.CPNP_P0, cutoff=0.01)
        S = O.relax_cpnp(S, 0.01, threads=8)
        _cmp_sets(eng, S, n)
    eng.close()


@pytest.mark.parametrize("mask", [1, 4])
def test_cpnp_p1_distance_uses_the_traceback_match_count(mask):
    # -p 1 (ArrangePosteriorProbs, MSA.cpp:1744-1752): distance = MEA score / number of matched columns
    seqs = synth.family(7, 90, seed=31) + [synth.family(1, 600, seed=32)[0]]
    n = len(seqs)
    i2 = 0.170705
    eng = engine(M.CPNP_P1, seqs, i2)
    eng.posterior_all_pairs(M.CPNP_P1, mask, 0.01)
    dist, S, _ = O.posterior_stage(O.CPNP_P1, mask, O.hmm_tables(i2), O.part_tables(O.CPNP_P0), seqs, threads=8)
    np.testing.assert_array_equal(eng.distances(), dist)
    _cmp_sets(eng, S, n)
    eng.close()


def test_cpnp_p1_mix_against_reference_fixture():
    d = load_golden("cpnp_sup139_p1mix")
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.CPNP_P1, seqs, float(d["initDistrib2"][0]))
    eng.posterior_all_pairs(M.CPNP_P1, 7, 0.01)
    np.testing.assert_allclose(eng.distances(), d["distances"], rtol=REL_TOL_PARTITION, atol=1e-6)
    from common import digest_of
    nnz, _, cc, _ = digest_of(eng.csr, n)
    np.testing.assert_array_equal(nnz, d["digest.s0.nnz"])
    np.testing.assert_array_equal(cc, d["digest.s0.col_crc"])
    eng.close()


@pytest.mark.parametrize("name", ["cpnp_p1_sup139", "cpnp_p1_BB12003", "cpnp_p1_676s4"])
def test_cpnp_non_progressive_program_on_the_device(name):
    """`c_p_np_aln -p 1` end to end through the C ABI (Viterbi statistics, -p 1 posteriors, two relaxations, alignment graph from
    the read-back set, similar-set refinement on the resident set) against the reference program's output."""
    d = load_golden(name)
    seqs = split_seqs(d)
    eng = engine(M.CPNP_P0, seqs, 0.700645)
    ident, ln = eng.viterbi_all_pairs()
    vm, _, _, i2 = M.cpnp_model_adjustment(ident, ln)
    h, p = M.default_tables(M.CPNP_P0, i2)
    eng.set_tables(h, p)
    eng.posterior_all_pairs(M.CPNP_P1, cpnp_mask(vm), 0.01)
    for _ in range(2):
        eng.relax(M.CPNP_P0, cutoff=0.01)
    seed = int(d["fixtime"][0])
    for key, ir in (("msa_ir0", 0), ("msa", 100)):
        assert eng.cpnp_np_finish_alignment(ir, seed) == [r.tobytes() for r in d[key]], key
    eng.close()


def test_cpnp_non_progressive_tail_device_equals_host_on_a_synthetic_family():
    """40 sequences x 120: the device refinement (k_profile_posterior + k_mea_wavefront) and the host provider walk the same
    set and must give the same rows; every row degaps to its input."""
    from common import cpnp_np_tail_from_csrset
    seqs = synth.family(40, 120, seed=91)
    n = len(seqs)
    eng = engine(M.CPNP_P1, seqs, 0.4)
    eng.posterior_all_pairs(M.CPNP_P1, 4, 0.01)
    eng.relax(M.CPNP_P0, cutoff=0.01)
    dist, S, _ = O.posterior_stage(O.CPNP_P1, 4, O.hmm_tables(0.4), O.part_tables(O.CPNP_P0), seqs, threads=8)
    S = O.relax_cpnp(S, 0.01, threads=8)
    np.testing.assert_array_equal(eng.distances(), dist)
    got = eng.cpnp_np_finish_alignment(100, 12345)
    assert got == cpnp_np_tail_from_csrset(S, seqs, dist, 100, 12345)
    assert len({len(r) for r in got}) == 1 and [r.replace(b"-", b"") for r in got] == seqs
    eng.close()


def test_cpnp_partition_beyond_fp64_range_is_rescaled():
    """Similar sequences of length 1500: Z ~ 1e600, beyond FP64 but inside the reference's 80-bit range.  The device runs
    FP64 with per-row power-of-two rescaling and must agree with the long-double oracle to 1e-5 relative."""
    seqs = synth.family(3, 1500, seed=77, p_sub=0.05, p_del=0.01, p_ins=0.01)
    n = len(seqs)
    eng = engine(M.CPNP_P0, seqs, 0.25)
    ht, pt = O.hmm_tables(0.25), O.part_tables(O.CPNP_P0)
    g = eng.debug_pair_dense(M.CPNP_P0, 2, 0, 1)
    ref, _ = O.model_posterior("part_cpnp", ht, pt, seqs[0], seqs[1])
    assert np.isfinite(g["part"]).all()
    big = ref > 1e-12
    np.testing.assert_allclose(g["part"][big], ref[big], rtol=REL_TOL_PARTITION)
    assert np.abs(g["part"][~big]).max() < 2e-12
    assert ref.max() > 0.9                                  # a real alignment, not a degenerate all-zero posterior
    eng.posterior_all_pairs(M.CPNP_P0, 2, 0.01)
    dist, S, rc = O.posterior_stage(O.CPNP_P0, 2, ht, pt, seqs, threads=4)
    assert rc == 0
    np.testing.assert_allclose(eng.distances(), dist, rtol=REL_TOL_PARTITION, atol=1e-6)
    for a, b in pairs(n):
        np.testing.assert_array_equal(eng.csr(a, b)[1], S.get(a, b)[1])
    eng.close()


@pytest.mark.parametrize("name", ["cpnp_sup002_ref", "cpnp_676s4_ref", "cpnp_sup139_mix"])
def test_viterbi_all_pairs_against_reference_fixture(name):
    d = load_golden(name)
    seqs = split_seqs(d)
    eng = engine(M.CPNP_P0, seqs, 0.700645)
    ident, ln = eng.viterbi_all_pairs()
    np.testing.assert_array_equal(ident, d["vit.ident"])
    np.testing.assert_array_equal(ln, d["vit.len"])
    vm, _, _, i2 = M.cpnp_model_adjustment(ident, ln)
    assert vm == int(d["variance_mean"][0]) and np.float32(i2) == d["initDistrib2"][0]
    eng.close()


@pytest.mark.parametrize("name", ["cpnp_sup002_ref", "cpnp_sup139_mix"])
def test_g_feature_line_from_gpu_alignments(name):
    d = load_golden(name)
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.CPNP_P0, seqs, 0.700645)
    ident, ln, aln, off = eng.viterbi_alignments()
    ht = O.hmm_tables()
    for p, (a, b) in enumerate(pairs(n)):
        assert aln[off[p]:off[p + 1]].tobytes() == O.viterbi(ht, seqs[a], seqs[b])[3]
    assert M.cpnp_g_features(seqs, aln, off) == d["gline"].tobytes()
    eng.close()


def test_viterbi_ragged_and_multiblock_vs_oracle():
    seqs = [synth.family(1, L, seed=300 + L)[0] for L in (1, 2, 33, 64, 200)] + synth.family(2, 620, seed=4, p_sub=0.3)
    n = len(seqs)
    eng = engine(M.CPNP_P0, seqs, 0.700645)
    ident, ln = eng.viterbi_all_pairs()
    ht = O.hmm_tables()
    want = [O.viterbi(ht, seqs[a], seqs[b])[1:3] for a, b in pairs(n)]
    np.testing.assert_array_equal(ident, [w[0] for w in want])
    np.testing.assert_array_equal(ln, [w[1] for w in want])
    eng.close()


def test_unknown_letters_and_identical_sequences():
    seqs = [b"ACDEFGHIKLMNPQRSTVWYBZX" * 3, b"ACDEFGHIKLMNPQRSTVWYBZX" * 3, b"XXBZACDWWWWWYYHHKKLMNP", b"MKV"]
    n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    dist, S, _ = O.posterior_stage(O.QP, 3, O.hmm_tables(), O.part_tables(O.QP), seqs, threads=4)
    np.testing.assert_array_equal(eng.distances(), dist)
    _cmp_sets(eng, S, n)
    eng.close()


def test_cpnp_rejects_letters_the_reference_cannot_score():
    eng = engine(M.CPNP_P0, [b"ACDJKL", b"ACDKLM"], 0.17)
    with pytest.raises(M.MlpError) as e:
        eng.posterior_all_pairs(M.CPNP_P0, 2, 0.01)
    assert e.value.code == -7
    eng.close()


# ------------------------------------------------------------------ size-independent properties at a larger size
def test_properties_on_a_large_family():
    seqs = synth.family_fast(120, 300, seed=11, p_sub=0.65)
    n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    d1 = eng.distances()
    assert np.array_equal(d1, d1.T) and (np.diag(d1) == 0).all() and (d1[np.triu_indices(n, 1)] > 0).all()
    nnz, rp, col, val = eng.csr_bulk()
    # determinism: a second run gives the same bytes
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    nnz2, rp2, col2, val2 = eng.csr_bulk()
    assert np.array_equal(nnz, nnz2) and np.array_equal(rp, rp2) and np.array_equal(col, col2) and np.array_equal(val, val2)
    assert np.array_equal(d1, eng.distances())
    # every stored value is a uint16 code / 65535 and >= the cut-off code
    codes = np.round(val.astype(np.float64) * 65535).astype(np.int64)
    assert np.array_equal((codes.astype(np.float32) / np.float32(65535)), val)
    assert (val >= np.float32(0.01) - np.float32(1.6e-5)).all() and (val <= 1).all()
    rng = np.random.default_rng(0)
    for _ in range(40):
        a, b = sorted(rng.choice(n, 2, replace=False))
        rp_ab, c_ab, v_ab = eng.csr(a, b)
        rp_ba, c_ba, v_ba = eng.csr(b, a)
        La, Lb = len(seqs[a]), len(seqs[b])
        dense = np.zeros((La + 1, Lb + 1), np.float32)
        for i in range(1, La + 1):
            cols = c_ab[rp_ab[i]:rp_ab[i + 1]]
            assert (np.diff(cols) > 0).all() and (cols >= 1).all() and (cols <= Lb).all()      # sorted, in range
            dense[i, cols] = v_ab[rp_ab[i]:rp_ab[i + 1]]
        denseT = np.zeros((Lb + 1, La + 1), np.float32)
        for j in range(1, Lb + 1):
            denseT[j, c_ba[rp_ba[j]:rp_ba[j + 1]]] = v_ba[rp_ba[j]:rp_ba[j + 1]]
        assert np.array_equal(dense.T, denseT)                                                  # stored transpose
        assert dense.sum(axis=1).max() < 1.5
    # pooled pinned read-back == per-pair read-back
    raw = eng.csr_raw()
    for a, b in [(0, 1), (5, 3), (n - 1, 2)]:
        rp1, c1, v1 = eng.csr(a, b)
        rp2, c2, v2 = raw.matrix(a, b, eng.lens)
        assert np.array_equal(rp1, rp2) and np.array_equal(c1, c2) and np.array_equal(v1, v2)
    raw.close()
    # the packed read-back (QuickProbs' uint16|uint16 cells + uint16 row sizes) decodes to the same matrices
    pk = eng.csr_packed()
    assert pk.nbytes() < 0.55 * (n * n * 12 + pk.rp_total * 4 + pk.used * 8)
    for a, b in [(0, 1), (5, 3), (n - 1, 2), (n - 2, n - 1)]:
        rp1, c1, v1 = eng.csr(a, b)
        rp2, c2, v2 = pk.matrix(a, b, eng.lens)
        assert np.array_equal(rp1, rp2) and np.array_equal(c1, c2) and np.array_equal(v1, v2)
    pk.close()
    # oracle spot-check of a few pairs at this size
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    for a, b in [(0, 1), (17, 93), (118, 119)]:
        post, dist, _ = O.pair_posterior(O.QP, 3, ht, pt, seqs[a], seqs[b])
        assert dist == d1[a, b]
        rp_ab, c_ab, v_ab = eng.csr(a, b)
        keep = np.argwhere(post >= np.float32(0.01))
        assert len(keep) == len(c_ab) and np.array_equal(keep[:, 1], c_ab)
    # relaxation: pattern can only shrink, result deterministic
    w, sd, _, _ = M.qp_guide_tree(d1)
    eng.relax(M.QP, np.maximum(w, np.float32(1e-6)), sd, 200.0, 3.0, float(np.float32(1e-5)))
    nnz3, rp3, col3, val3 = eng.csr_bulk()
    assert (nnz3 <= nnz).all() and nnz3.sum() > 0
    eng.close()


def test_full_size_family_properties():
    """BASELINE config A at full size (1,000 sequences x 300, the family bench.py times; 499,500 pairs, 4.7e10 cells per model):
    size-independent properties plus oracle spot checks.  Pairs are independent in the posterior stage, so any sub-family
    processed on its own must give the same bytes as the same pairs inside the full run."""
    seqs = synth.family_fast(1000, 300, seed=20220148 + 2)
    n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    d = eng.distances()
    assert d.shape == (n, n) and np.array_equal(d, d.T) and (np.diag(d) == 0).all()
    off = d[np.triu_indices(n, 1)]
    assert np.isfinite(off).all() and (off > 0).all() and (off <= 1).all()
    # a sub-family on its own context: same distances, same matrices (both orientations)
    pick = [0, 1, 2, 3, 499, 500, 501, 997, 998, 999]
    sub = engine(M.QP, [seqs[k] for k in pick])
    sub.posterior_all_pairs(M.QP, 3, 0.01)
    assert np.array_equal(sub.distances(), d[np.ix_(pick, pick)])
    for x in range(len(pick)):
        for y in range(len(pick)):
            if x != y:
                r1, c1, v1 = sub.csr(x, y)
                r2, c2, v2 = eng.csr(pick[x], pick[y])
                assert np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.array_equal(v1, v2), (x, y)
    sub.close()
    # oracle on three pairs spread over the (cost-sorted, batched) task list
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    for a, b in [(0, 999), (123, 456), (998, 999)]:
        post, dist, _ = O.pair_posterior(O.QP, 3, ht, pt, seqs[a], seqs[b])
        assert dist == d[a, b]
        rp, c, v = eng.csr(a, b)
        keep = np.argwhere(post >= np.float32(0.01))
        assert len(keep) == len(c) and np.array_equal(keep[:, 1], c)
        codes = np.floor(post[keep[:, 0], keep[:, 1]].astype(np.float32) * np.float32(65535)).astype(np.float32)
        assert np.array_equal(codes / np.float32(65535), v)
    cells_before = eng.total_cells()
    before = {p: eng.csr(*p) for p in [(0, 999), (999, 0), (123, 456), (700, 40)]}
    # consistency (one repetition with the final cut-off, as QuickProbs runs it above 50 sequences): patterns only shrink,
    # the stored transpose stays the transpose, values stay uint16 codes
    w, sd, _, _ = M.qp_guide_tree(d)
    eng.relax(M.QP, np.maximum(w, np.float32(1e-6)), sd, 200.0, 3.0, float(np.float32(1e-5)))
    assert 0 < eng.total_cells() <= cells_before
    for (a, b), (rp0, c0, v0) in before.items():
        rp1, c1, v1 = eng.csr(a, b)
        La = len(seqs[a])
        for i in range(1, La + 1):
            assert set(c1[rp1[i]:rp1[i + 1]].tolist()) <= set(c0[rp0[i]:rp0[i + 1]].tolist())
        assert np.array_equal(np.round(v1.astype(np.float64) * 65535).astype(np.float32) / np.float32(65535), v1)
    rp_ab, c_ab, v_ab = eng.csr(123, 456)
    rp_ba, c_ba, v_ba = eng.csr(456, 123)
    fwd = {(i, int(c)): float(v) for i in range(1, len(seqs[123]) + 1) for c, v in zip(c_ab[rp_ab[i]:rp_ab[i + 1]], v_ab[rp_ab[i]:rp_ab[i + 1]])}
    rev = {(int(c), j): float(v) for j in range(1, len(seqs[456]) + 1) for c, v in zip(c_ba[rp_ba[j]:rp_ba[j + 1]], v_ba[rp_ba[j]:rp_ba[j + 1]])}
    assert fwd == rev and len(fwd) > 0
    eng.close()


# ------------------------------------------------------------------ capacity misses, reused contexts
def _same_sets(eng, S, n):
    for a in range(n):
        for b in range(n):
            if a != b:
                rp, c, v = eng.csr(a, b)
                orp, oc, ov = S.get(a, b)
                assert np.array_equal(rp, orp) and np.array_equal(c, oc) and np.array_equal(v, ov), (a, b)


def test_capacity_misses_are_retried_and_leave_earlier_families_intact():
    """cutoff = 0 keeps every inner cell, so a pair stages far more hits than the per-warp staging buffer holds (error bit 1)
    and the cell pool, configured tiny, overflows as well (error bit 2): k_final must publish nothing for the pairs it could
    not write, k_transpose must not index with stale columns, and the re-run after growing must equal the oracle -- on a
    context that has already served another family (directory mode re-uses the pools)."""
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    first = synth.family(5, 60, seed=5)
    eng = engine(M.QP, first)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    d1, S1, _ = O.posterior_stage(O.QP, 3, ht, pt, first)
    _same_sets(eng, S1, len(first))
    seqs = synth.family(7, 90, seed=9)
    n = len(seqs)
    eng.configure(0, 64)                       # 64 cells: every batch overflows the pool at least once
    eng.set_sequences(seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.0)
    all_cells = 2 * sum(len(a) * len(b) for i, a in enumerate(seqs) for b in seqs[i + 1:]) + 1024
    dist, S, rc = O.posterior_stage(O.QP, 3, ht, pt, seqs, cutoff=0.0, nz_cap=all_cells)
    assert rc == 0
    np.testing.assert_array_equal(eng.distances(), dist)
    _same_sets(eng, S, n)
    # and the normal cutoff afterwards on the same context (pools now larger than needed)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    dist, S, _ = O.posterior_stage(O.QP, 3, ht, pt, seqs)
    _same_sets(eng, S, n)
    eng.close()


def test_debug_pair_dense_leaves_a_relaxed_set_untouched():
    seqs = synth.family(6, 70, seed=21)
    n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    t = M.qp_guide_tree_ex(eng.distances())
    w = np.maximum(t["weights"], np.float32(1e-6))
    eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, 0.01)
    before = [[eng.csr(a, b) for b in range(n) if b != a] for a in range(n)]
    dist_before = eng.distances().copy()
    eng.debug_pair_dense(M.QP, 3, 1, 4)
    after = [[eng.csr(a, b) for b in range(n) if b != a] for a in range(n)]
    for ra, rb in zip(before, after):
        for x, y in zip(ra, rb):
            assert all(np.array_equal(p, q) for p, q in zip(x, y))
    np.testing.assert_array_equal(eng.distances(), dist_before)
    eng.close()


def test_qp_selectivity_fixture_400_sequences():
    """Reference-pinned relaxation where QuickProbs' selectivity really rejects third sequences (ConsistencyStage.cpp:181-216):
    400 synthetic sequences in eight sub-families, ref_qp dump reduced to per-pair nnz + CRC32(row pointers | columns | values)."""
    import zlib
    d = load_golden("qp_syn400")
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    np.testing.assert_array_equal(eng.distances(), d["distances"].reshape(n, n))
    t = M.qp_guide_tree_ex(eng.distances())
    np.testing.assert_array_equal(t["weights"], d["weights"])
    np.testing.assert_array_equal(t["seldist"].reshape(-1), d["seldist"].reshape(-1))
    sd = t["seldist"].reshape(n, n)
    accepted = np.array([(np.maximum(sd[a], sd[b]) <= 200).sum() - 2 for a, b in pairs(n)])
    assert (accepted < n - 2).mean() > 0.5

    def digest(tag, transposed):
        raw = eng.csr_raw()
        try:
            for p, (a, b) in enumerate(pairs(n)):
                x, y = (b, a) if transposed else (a, b)
                slot = x * n + y
                cnt = int(raw.nz_cnt[slot]); off = int(raw.nz_off[slot]); ro = int(raw.rp_off[slot])
                assert cnt == int(d["digest.%s.nnz" % tag][p]), (tag, a, b)
                cells = raw.cells[off:off + cnt]
                c = zlib.crc32(np.ascontiguousarray(raw.rp_pool[ro:ro + int(eng.lens[x]) + 2]).tobytes())
                c = zlib.crc32(np.ascontiguousarray(cells["col"].astype(np.int32)).tobytes(), c)
                c = zlib.crc32(np.ascontiguousarray(cells["val"]).tobytes(), c) & 0xffffffff
                assert c == int(d["digest.%s.crc" % tag][p]), (tag, a, b)
        finally:
            raw.close()
    digest("s0", False)
    w = np.maximum(t["weights"], np.float32(1e-6))
    eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, float(np.float32(1e-5)))    # N > 50: one repetition
    digest("sF", False)
    digest("tF", True)
    eng.close()


def test_split_read_back_overlapping_the_next_posterior_stage():
    """mlp_get_csr_packed_begin / _end: the packed copy of a relaxed set runs on its own stream while the same family is
    submitted again and its posterior stage runs; what arrives equals the synchronous read-back, and the second pass is
    unharmed."""
    seqs = synth.family(9, 110, seed=44)
    eng = engine(M.QP, seqs)

    def stage():
        eng.posterior_all_pairs(M.QP, 3, 0.01)
        t = M.qp_guide_tree_ex(eng.distances())
        w = np.maximum(t["weights"], np.float32(1e-6))
        return w, t["seldist"]

    w, sd = stage()
    eng.relax(M.QP, w, sd, 200.0, 3.0, 0.01)
    ref = eng.csr_packed()
    want = (ref.nz_off.copy(), ref.nz_cnt.copy(), ref.row_sizes[:ref.rp_total].copy(), ref.cells[:ref.used].copy())
    lay = eng.csr_layout()
    out = M.PinnedPackedBuffers(len(seqs), lay[1], lay[2])
    eng.csr_packed_begin(out)
    eng.set_sequences(seqs)                      # same shape: the pools stay, the read-back goes on
    w2, sd2 = stage()                            # posterior of the "next family" beside the copy
    eng.relax(M.QP, w2, sd2, 200.0, 3.0, 0.01)   # ends the read-back before it overwrites the set
    got = (out.nz_off, out.nz_cnt, out.row_sizes[:out.rp_total], out.cells[:out.used])
    for x, y in zip(want, got):
        np.testing.assert_array_equal(x, y)
    again = eng.csr_packed()                     # the second pass: same matrices (their places in the pool may differ)
    np.testing.assert_array_equal(again.nz_cnt, want[1])
    n = len(seqs)
    for a in range(n):
        for b in range(n):
            if a != b:
                for x, y in zip(ref.matrix(a, b, eng.lens), again.matrix(a, b, eng.lens)):
                    np.testing.assert_array_equal(x, y)
    eng.csr_packed_end()
    eng.close()


@pytest.mark.parametrize("lens", [[40, 37, 52], [300, 290, 310, 305], [20, 170, 33, 400], [700, 650], [5, 31, 33, 64, 257, 511]])
def test_register_band_kernels_equal_the_round1_kernels(lens, monkeypatch):
    """Every register-band kernel (sweep_c.cuh: k_part_fwd_c, k_part_rev_c, k_hmm_fwd_c, k_hmm_bwd_c, k_final_c) in isolation and all
    together against the round-1 shared-memory-band kernels they replaced (MLP_OLD_SWEEP = bit mask of kernels that fall back,
    read at every launch): distances and every matrix bit for bit, on ragged families that mix columns-per-lane groups, one
    and several column blocks, and sequences shorter than a warp."""
    rng = np.random.default_rng(sum(lens))
    al = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    base = al[rng.integers(0, 20, max(lens))]
    seqs = []
    for L in lens:
        s = base[:L].copy(); m = rng.random(L) < 0.4; s[m] = al[rng.integers(0, 20, int(m.sum()))]; seqs.append(s.tobytes())
    n = len(seqs)

    def run(mask):
        monkeypatch.setenv("MLP_OLD_SWEEP", str(mask))
        eng = engine(M.QP, seqs)
        eng.posterior_all_pairs(M.QP, 3, 0.01)
        out = (eng.distances().copy(), [eng.csr(a, b) for a in range(n) for b in range(n) if a != b])
        eng.close()
        return out

    ref = run(31)
    for name, mask in (("part_fwd", 30), ("part_rev", 29), ("hmm_fwd", 27), ("hmm_bwd", 23), ("final", 15), ("all", 0)):
        got = run(mask)
        np.testing.assert_array_equal(ref[0], got[0], err_msg=name)
        for r, g in zip(ref[1], got[1]):
            for x, y in zip(r, g):
                np.testing.assert_array_equal(x, y, err_msg=name)
    # and against the oracle, for the new kernels alone
    dist, S, _ = O.posterior_stage(O.QP, 3, O.hmm_tables(), O.part_tables(O.QP), seqs)
    np.testing.assert_array_equal(run(0)[0], dist)


@pytest.mark.parametrize("lens", [[40, 37, 52], [300, 290, 310, 305], [20, 170, 33, 400], [700, 650], [5, 31, 33, 64, 257, 511], [1, 1, 2, 3]])
@pytest.mark.parametrize("mask", [1, 2, 4, 7])
def test_local_model_register_band_kernels_equal_the_round1_kernels(lens, mask, monkeypatch):
    """cpnp's 3-state local model on the register-band sweeps with the filtered, thread-per-pair Z chain (loc_c.cu: k_loc_fwd_c,
    k_loc_bwd_c, k_loc_cand_c, k_loc_replay) and the C-specialised merge kernel for c_p_np_aln's model mixes (final_c.cu, modes 1, 2, 4, 7)
    and the rescaled FP64 partition sweeps on the same skeleton (part_sc.cu)
    against the round-1 kernels (MLP_OLD_SWEEP bits 32, 16, 2 and 1: warp-per-pair replay of the complete row-major layer without a
    candidate filter; general merge kernel; shared-memory-band partition sweeps): distances and every matrix bit for bit, every model alone and all three merged (mask
    7, where the Z terms and the candidate lists alias the partition layer); mask 4 also against the oracle."""
    rng = np.random.default_rng(sum(lens) + mask)
    al = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    base = al[rng.integers(0, 20, max(lens))]
    seqs = []
    for L in lens:
        s = base[:L].copy(); m = rng.random(L) < 0.4; s[m] = al[rng.integers(0, 20, int(m.sum()))]; seqs.append(s.tobytes())
    n = len(seqs)

    def run(old):
        monkeypatch.setenv("MLP_OLD_SWEEP", str(old))
        eng = engine(M.CPNP_P0, seqs)
        eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)
        out = (eng.distances().copy(), [eng.csr(a, b) for a in range(n) for b in range(a + 1, n)])
        eng.close()
        return out

    ref, got = run(51), run(0)
    np.testing.assert_array_equal(ref[0], got[0])
    for r, g in zip(ref[1], got[1]):
        for x, y in zip(r, g):
            np.testing.assert_array_equal(x, y)
    if mask == 4:
        dist, S, _ = O.posterior_stage(O.CPNP_P0, 4, O.hmm_tables(), O.part_tables(O.CPNP_P0), seqs)
        np.testing.assert_array_equal(got[0], dist)


@pytest.mark.parametrize("case", ["posterior_9", "posterior_dups", "random_1000", "ties_1000", "ties_257", "two", "three"])
def test_device_guide_tree_equals_the_host_tree(case):
    """QuickProbs' guide tree built on the device (tree_dev.cu: single-CTA UPGMA, weights, subtree distances) against the host
    restatement (host_tree.cpp, itself pinned on the reference's trees): parent / children, weights and selectivity distances bit
    for bit -- on distances a posterior stage produced (with duplicated sequences: exact ties) and on injected matrices
    quantised to a handful of values, where almost every merge has to break ties the way the reference's row-major scan does.
    A consistency repetition fed from the resident tree equals one fed with the host arrays."""
    rng = np.random.default_rng(len(case))
    if case.startswith("posterior"):
        seqs = synth.family(9, 60, seed=11)
        if case == "posterior_dups":
            seqs = [seqs[i % 4] if i % 3 else seqs[i] for i in range(9)] + seqs[:3]
        eng = engine(M.QP, seqs)
        eng.posterior_all_pairs(M.QP, 3, 0.01)
    else:
        n = {"random_1000": 1000, "ties_1000": 1000, "ties_257": 257, "two": 2, "three": 3}[case]
        seqs = synth.family(n, 8, seed=5)
        eng = engine(M.QP, seqs)
        eng.posterior_all_pairs(M.QP, 3, 0.01)          # allocates the sets and the distance matrix
        d = rng.random((n, n)).astype(np.float32)
        if case.startswith("ties"):
            d = (np.floor(d * 6) / 8 + 0.125).astype(np.float32)
        d = np.triu(d, 1); d = d + d.T
        eng.debug_set_distances(d)
    n = len(seqs)
    d = eng.distances()
    host = M.qp_guide_tree_ex(d.copy())
    dev = eng.qp_guide_tree_device(min_weight=0.0, want_seldist=True)
    for k in ("parent", "left", "right", "weights", "seldist"):
        np.testing.assert_array_equal(np.asarray(host[k]).reshape(-1), np.asarray(dev[k]).reshape(-1), err_msg=k)
    np.testing.assert_array_equal(eng.distances(), d)   # the resident matrix is not the one the clustering consumed
    if case.startswith("posterior"):
        w = np.maximum(host["weights"], np.float32(1e-6))
        eng.qp_guide_tree_device(min_weight=1e-6)
        eng.relax(M.QP, None, None, 200.0, 3.0, 0.01)
        got = [eng.csr(a, b) for a in range(n) for b in range(n) if a != b]
        eng.posterior_all_pairs(M.QP, 3, 0.01)
        eng.relax(M.QP, w, host["seldist"], 200.0, 3.0, 0.01)
        ref = [eng.csr(a, b) for a in range(n) for b in range(n) if a != b]
        for r, g in zip(ref, got):
            for x, y in zip(r, g):
                np.testing.assert_array_equal(x, y)
    eng.close()


def _two_pass_digest(eng, n, scratch=None):
    """The streamed flow for families whose set does not fit HBM: streamed stage over all pairs (finished + digested + dropped batch
    by batch), guide tree, ordinary stage + consistency restricted to the pairs that accept third sequences."""
    if scratch:
        eng.configure(scratch, 0)
    eng.stream_begin(1)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    streamed = eng.stream_end().reshape(n, n)
    tree = eng.qp_guide_tree_device(1e-6, want_seldist=True)
    sd = tree["seldist"].reshape(n, n)
    eng.restrict_pairs(sd, 200.0)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    eng.relax(M.QP, None, None, 200.0, 3.0, float(np.float32(1e-5)))
    relaxed = eng.set_digest().reshape(n, n)
    eng.set_shard(0, 1)
    ingroup = sd <= 200.0
    np.fill_diagonal(ingroup, False)
    return np.where(ingroup, relaxed, streamed), ingroup


def test_streamed_two_pass_flow_equals_the_ordinary_flow():
    """BASELINE config #5's flow (mlp_stream_begin / mlp_restrict_pairs: cell pool recycled batch by batch, only the pairs inside a
    <= 200-leaf subtree are kept, exchanged and relaxed) on the reference-pinned 400-sequence family where QuickProbs' selectivity
    rejects third sequences: every matrix's digest after the consistency repetition equals the ordinary flow's (which
    test_qp_selectivity_fixture_400_sequences pins on the reference), with the dense scratch capped so that the streamed stage
    really runs many batches."""
    d = load_golden("qp_syn400")
    seqs = split_seqs(d); n = len(seqs)
    eng = engine(M.QP, seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    t = eng.qp_guide_tree_device(1e-6)
    eng.relax(M.QP, None, None, 200.0, 3.0, float(np.float32(1e-5)))
    ref = eng.set_digest().reshape(n, n)
    eng.close()
    eng = engine(M.QP, seqs)
    got, ingroup = _two_pass_digest(eng, n, scratch=64 << 20)
    eng.close()
    assert 0.05 < ingroup.mean() < 0.6          # both kinds of pairs are present
    np.testing.assert_array_equal(got, ref)


def test_local_model_bound_check_failure_falls_back_to_the_unfiltered_kernels(monkeypatch):
    """loc_c.cu's filtered Z chain checks its one assumption (running sum >= bound - 0.5) at every candidate; a failed check sets bit
    16 of the error word and the batch is redone by the round-1 kernels (no filter, no assumption).  MLP_LOC_FORCE_FALLBACK makes the
    chain report a failure: the stage must give the same distances and matrices as an undisturbed run, on the same context twice
    (the context stays on the round-1 kernels afterwards)."""
    seqs = synth.family(7, 90, seed=21)
    n = len(seqs)

    def run(eng):
        eng.posterior_all_pairs(M.CPNP_P0, 4, 0.01)
        return eng.distances().copy(), [eng.csr(a, b) for a in range(n) for b in range(a + 1, n)]

    eng = engine(M.CPNP_P0, seqs)
    ref = run(eng)
    eng.close()
    monkeypatch.setenv("MLP_LOC_FORCE_FALLBACK", "1")
    lib_eng = engine(M.CPNP_P0, seqs)
    for _ in range(2):
        got = run(lib_eng)
        np.testing.assert_array_equal(ref[0], got[0])
        for r, g in zip(ref[1], got[1]):
            for x, y in zip(r, g):
                np.testing.assert_array_equal(x, y)
    lib_eng.close()
