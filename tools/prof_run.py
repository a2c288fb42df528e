#!/usr/bin/env python3
"""One small pass of the hot path for profiling under ncu (tools/: developer scripts, not product)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlprobs_b200 as M
from mlprobs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
flav = sys.argv[3] if len(sys.argv) > 3 else "qp"
seqs = synth.family_fast(n, L, seed=20220150)
eng = M.Engine(0)
if flav == "qp":
    h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(seqs)
    import time
    for rep in range(2):
        t0 = time.perf_counter(); eng.posterior_all_pairs(M.QP, 3, 0.01); t1 = time.perf_counter()
        st = eng.stats()
        print("posterior wall %.1f ms device %.1f ms" % ((t1 - t0) * 1e3, st["ms_total"]), {k: round(v, 2) for k, v in st["ms_kernel"].items() if v})
        t0 = time.perf_counter(); d = eng.distances(); t1 = time.perf_counter(); w, sd, _, _ = M.qp_guide_tree(d); t2 = time.perf_counter()
        print("distances d2h %.1f ms, tree %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
        t0 = time.perf_counter(); eng.relax(M.QP, np.maximum(w, np.float32(1e-6)), sd, 200.0, 3.0, float(np.float32(1e-5))); t1 = time.perf_counter()
        print("relax wall %.1f ms device %.1f ms kernel %.1f" % ((t1 - t0) * 1e3, eng.stats()["ms_total"], eng.stats()["ms_kernel"]["relax"]), "cells", eng.total_cells())
else:
    h, p = M.default_tables(M.CPNP_P0, 0.100675); eng.set_tables(h, p); eng.set_sequences(seqs)
    eng.posterior_all_pairs(M.CPNP_P0, 4, 0.01)
    print("posterior", eng.stats()["ms_kernel"])
    eng.relax(M.CPNP_P0, cutoff=0.01)
    print("relax", eng.stats()["ms_kernel"]["relax"], "cells", eng.total_cells())
eng.close()
