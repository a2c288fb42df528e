"""Developer GPU check (not a pytest): CUDA path vs oracle, verbose diagnostics."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as O
import mlprobs_b200 as M
from mlprobs_b200 import synth


def cmp_dense(name, g, o):
    eq = np.array_equal(g, o)
    if eq:
        print("   %-8s exact" % name)
        return True
    d = np.abs(g.astype(np.float64) - o.astype(np.float64))
    bad = np.argwhere(g != o)
    rel = d / np.maximum(np.abs(o), 1e-30)
    print("   %-8s MISMATCH cells=%d/%d maxabs=%.3e maxrel(where o>1e-6)=%.3e first=%s g=%r o=%r" % (
        name, len(bad), g.size, d.max(), (rel * (np.abs(o) > 1e-6)).max(), bad[0], g[tuple(bad[0])], o[tuple(bad[0])]))
    return False


def cmp_sets(eng, S, n, tag):
    ok = True
    nb = 0
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            rp, c, v = eng.csr(a, b)
            orp, oc, ov = S.get(a, b)
            if not (np.array_equal(rp, orp) and np.array_equal(c, oc) and np.array_equal(v, ov)):
                nb += 1
                if ok:
                    print("   %s first mismatch pair (%d,%d): nnz %d vs %d rp_eq=%s col_eq=%s val_eq=%s" % (
                        tag, a, b, len(c), len(oc), np.array_equal(rp, orp), np.array_equal(c, oc),
                        len(v) == len(ov) and np.array_equal(v, ov)))
                    if len(v) == len(ov) and np.array_equal(c, oc):
                        k = np.argwhere(v != ov)[0][0]
                        print("      val[%d] %r vs %r" % (k, v[k], ov[k]))
                ok = False
    print("   %s sets %s (%d mismatching matrices)" % (tag, "EXACT" if ok else "DIFFER", nb))
    return ok


def run(flavour, mask, seqs, init2=0.170705, dense=True, relax_reps=2, label=""):
    n = len(seqs)
    print("== %s flavour=%d mask=%d n=%d lens=%s" % (label, flavour, mask, n, [len(s) for s in seqs][:8]))
    oflav = {M.QP: O.QP, M.CPNP_P0: O.CPNP_P0}[flavour]
    oht = O.hmm_tables(init2 if flavour != M.QP else 0.700645)
    opt = O.part_tables(oflav)
    h, p = M.default_tables(flavour, init2)
    eng = M.Engine(0)
    eng.set_tables(h, p)
    eng.set_sequences(seqs)
    allok = True
    if dense:
        for (a, b) in [(0, 1), (n - 2, n - 1)]:
            g = eng.debug_pair_dense(flavour, mask, a, b)
            print("  pair", a, b)
            if mask & 1:
                o5, _ = O.model_posterior("hmm5", oht, opt, seqs[a], seqs[b], 1 if flavour == M.QP else 0)
                allok &= cmp_dense("hmm5", g["hmm5"], o5)
            if mask & 2:
                op_, _ = O.model_posterior("part_qp" if flavour == M.QP else "part_cpnp", oht, opt, seqs[a], seqs[b])
                if flavour == M.QP:
                    allok &= cmp_dense("part", g["part"], op_)
                else:
                    d = np.abs(g["part"].astype(np.float64) - op_)
                    rel = (d / np.maximum(op_, 1e-30))[op_ > 1e-10]
                    print("   part(cpnp, fp64 vs long double) maxrel=%.3e nexact=%d/%d" % (rel.max() if rel.size else 0, (g["part"] == op_).sum(), op_.size))
            if mask & 4:
                ol, _ = O.model_posterior("local", oht, opt, seqs[a], seqs[b])
                allok &= cmp_dense("local", g["local"], ol)
            om, od, _ = O.pair_posterior(oflav, mask, oht, opt, seqs[a], seqs[b])
            if flavour == M.QP or not (mask & 2):
                allok &= cmp_dense("merged", g["merged"], om)
                print("   dist gpu=%r oracle=%r %s" % (g["dist"], od, "ok" if g["dist"] == od else "DIFF"))
            else:
                d = np.abs(g["merged"].astype(np.float64) - om)
                print("   merged maxabs=%.3e dist gpu=%r oracle=%r" % (d.max(), g["dist"], od))
    t0 = time.time()
    eng.posterior_all_pairs(flavour, mask, 0.01)
    t1 = time.time()
    st = eng.stats()
    print("  posterior stage: wall %.3fs device %.3f ms cells=%d kernels=%s" % (t1 - t0, st["ms_total"], st["cells"], {k: round(v, 3) for k, v in st["ms_kernel"].items() if v}))
    odist, S, rc = O.posterior_stage(oflav, mask, oht, opt, seqs, threads=8)
    gd = eng.distances()
    exact_expected = (flavour == M.QP) or not (mask & 2)
    if exact_expected:
        print("  distances", "EXACT" if np.array_equal(gd, odist) else "DIFFER maxabs=%.3e" % np.abs(gd - odist).max())
        allok &= np.array_equal(gd, odist)
        allok &= cmp_sets(eng, S, n, "posterior")
    else:
        print("  distances maxabs=%.3e" % np.abs(gd - odist).max())
        # index sets
        same = sum(np.array_equal(eng.csr(a, b)[1], S.get(a, b)[1]) for a in range(n) for b in range(a + 1, n))
        print("  index sets identical for %d/%d pairs" % (same, n * (n - 1) // 2))
    if exact_expected:
        cur = S
        for r in range(relax_reps):
            if flavour == M.QP:
                rng = np.random.default_rng(5)
                w = (rng.random(n) * 0.5 + 0.5).astype(np.float32)
                sd = rng.integers(2, 400, size=(n, n)).astype(np.float32)
                sd = np.maximum(sd, sd.T)
                cutoff = 0.01 if r < relax_reps - 1 else float(np.float32(1e-5))
                eng.relax(M.QP, w, sd, 200.0, 3.0, cutoff)
                cur = O.relax_qp(cur, w, sd, cutoff, 200.0, 3.0, threads=8)
            else:
                eng.relax(flavour, cutoff=0.01)
                cur = O.relax_cpnp(cur, 0.01, threads=8)
            st = eng.stats()
            print("  relax rep %d: device %.3f ms" % (r, st["ms_total"]))
            allok &= cmp_sets(eng, cur, n, "relax%d" % r)
    eng.close()
    return allok


if __name__ == "__main__":
    ok = True
    small = synth.family(6, 60, seed=1)
    ok &= run(M.QP, 3, small, label="QP small")
    ok &= run(M.CPNP_P0, 4, small, label="cpnp local small")
    ok &= run(M.CPNP_P0, 1, small, label="cpnp hmm5-only small")
    ok &= run(M.CPNP_P0, 2, small, label="cpnp partition small")
    ok &= run(M.CPNP_P0, 7, small, label="cpnp mix small")
    mid = synth.family(5, 300, seed=2)
    ok &= run(M.QP, 3, mid, label="QP L300")
    ok &= run(M.CPNP_P0, 4, mid, label="cpnp local L300")
    ragged = [synth.family(1, L, seed=10 + L)[0] for L in (1, 2, 31, 33, 64, 530, 700)]
    ok &= run(M.QP, 3, ragged, label="QP ragged/multiblock", dense=True)
    ok &= run(M.CPNP_P0, 4, ragged, label="cpnp local ragged/multiblock", dense=True)
    print("ALL OK" if ok else "SOME CHECKS FAILED")
    sys.exit(0 if ok else 1)
