"""Developer A/B: every register-band kernel (sweep_c.cuh) in isolation against the round-1 kernels (MLP_OLD_SWEEP bit mask)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlprobs_b200 as M
from mlprobs_b200 import synth

def run(seqs, mask):
    os.environ["MLP_OLD_SWEEP"] = str(mask)
    eng = M.Engine(0)
    h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    d = eng.distances().copy()
    raw = eng.csr_raw()
    out = (d, raw.nz_cnt.copy(), [eng.csr(a, b) for a in range(len(seqs)) for b in range(len(seqs)) if a != b])
    raw.close(); eng.close()
    return out

for lens in ([40, 37, 52], [300, 290, 310, 305], [20, 170, 33, 400], [700, 650]):
    rng = np.random.default_rng(sum(lens))
    al = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    base = al[rng.integers(0, 20, max(lens))]
    seqs = []
    for L in lens:
        s = base[:L].copy(); m = rng.random(L) < 0.4; s[m] = al[rng.integers(0, 20, int(m.sum()))]; seqs.append(s.tobytes())
    ref = run(seqs, 31)
    for name, mask in (("part_fwd", 31 - 1), ("part_rev", 31 - 2), ("hmm_fwd", 31 - 4), ("hmm_bwd", 31 - 8), ("final", 31 - 16), ("all", 0)):
        got = run(seqs, mask)
        okd = np.array_equal(ref[0], got[0]); okn = np.array_equal(ref[1], got[1])
        okc = all(np.array_equal(x, y) for r, g in zip(ref[2], got[2]) for x, y in zip(r, g))
        print(lens, name, "dist", okd, "nnz", okn, "cells", okc, "" if okd else "maxdiff %.3g" % np.abs(ref[0] - got[0]).max(), flush=True)
