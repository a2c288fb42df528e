#!/usr/bin/env python3
"""Time the QuickProbs tail (construction + refinement) on the device-resident set (tools/: developer script)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlprobs_b200 as M
from mlprobs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
host_too = len(sys.argv) > 3 and sys.argv[3] == "host"
seqs = synth.family_fast(n, L, seed=20220150)
eng = M.Engine(0)
h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(seqs)
t0 = time.perf_counter(); eng.posterior_all_pairs(M.QP, 3, 0.01); t1 = time.perf_counter()
t = M.qp_guide_tree_ex(eng.distances())
w = np.maximum(t["weights"], np.float32(1e-6))
iters = 1 if n > 50 else 2
for it in range(iters):
    eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, 0.01 if it < iters - 1 else float(np.float32(1e-5)))
t2 = time.perf_counter()
print("n=%d L=%d posterior %.1f ms, tree+relax %.1f ms, cells %d" % (n, L, (t1 - t0) * 1e3, (t2 - t1) * 1e3, eng.total_cells()))
for ref_iters, tag in ((-2, "warm-up (construction only)"), (-2, "construction only"), (-1, "construction + refinement"), (-1, "construction + refinement (again)")):
    t3 = time.perf_counter(); rows = eng.qp_finish_alignment(w, t["left"], t["right"], ref_iters); t4 = time.perf_counter()
    st = eng.stats()
    print("%s: wall %.1f ms, kernels %.1f ms over %d launches, pair-matrices summed %d, h2d %.1f MB d2h %.1f MB, columns %d"
          % (tag, (t4 - t3) * 1e3, st["ms_total"], st["launches"], st["pairs"], st["h2d_bytes"] / 1e6, st["d2h_bytes"] / 1e6, len(rows[0])))
if host_too:
    raw = eng.csr_raw()
    t5 = time.perf_counter()
    rows_h = M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], raw.rp_off, raw.nz_off, raw.rp_pool, raw.cells)
    t6 = time.perf_counter()
    print("host tail (1 thread): %.1f ms, identical=%s" % ((t6 - t5) * 1e3, rows_h == rows))
eng.close()
