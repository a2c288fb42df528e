#!/usr/bin/env python3
"""Wall-clock breakdown of one end-to-end step at the bench workload (tools/: developer script)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlprobs_b200 as M
from mlprobs_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
seqs = synth.family_fast(n, L, seed=20220150)
eng = M.Engine(0)
h, p = M.default_tables(M.QP); eng.set_tables(h, p)
host_out = None
for rep in range(3):
    T = [time.perf_counter()]
    def tick(): T.append(time.perf_counter())
    eng.set_sequences(seqs); tick()
    eng.posterior_all_pairs(M.QP, 3, 0.01); tick(); dev_post = eng.stats()["ms_total"]
    d = eng.distances(); tick()
    w, sd, _, _ = M.qp_guide_tree(d); tick()
    eng.relax(M.QP, np.maximum(w, np.float32(1e-6)), sd, 200.0, 3.0, float(np.float32(1e-5))); tick(); dev_relax = eng.stats()["ms_total"]
    if host_out is None:
        lay = eng.csr_layout(); host_out = M.PinnedPackedBuffers(n, lay[1], int(lay[2] * 1.05)); T[-1] = time.perf_counter()
    out = eng.csr_packed(host_out); tick()
    names = ["set_sequences", "posterior", "distances", "tree", "relax", "csr_packed"]
    ms = [(T[i + 1] - T[i]) * 1e3 for i in range(len(names))]
    print("rep %d total %.0f ms | " % (rep, sum(ms)) + ", ".join("%s %.0f" % (a, b) for a, b in zip(names, ms)) +
          " | device: posterior %.0f relax %.0f | d2h %.2f GB" % (dev_post, dev_relax, out.nbytes() / 1e9))
eng.close()
