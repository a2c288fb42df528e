#!/usr/bin/env python3
"""A/B of the local model's kernels (loc_c.cu against the round-1 k_loc_*): per-kernel times and equality of the distances on one
synthetic family (developer script).  Usage: loc_ab.py [n] [L] [mask]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlprobs_b200 as M
from mlprobs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
mask = int(sys.argv[3]) if len(sys.argv) > 3 else 4
seqs = synth.family_fast(n, L, seed=20220150)
out = {}
for old in (32, 0):
    os.environ["MLP_OLD_SWEEP"] = str(old)
    eng = M.Engine(0)
    h, p = M.default_tables(M.CPNP_P0, 0.100675); eng.set_tables(h, p); eng.set_sequences(seqs)
    for rep in range(2):
        eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)
    st = eng.stats()
    out[old] = (eng.distances().copy(), eng.total_cells())
    if old == 0 and os.environ.get("MLP_LOC_SPLIT"):
        import ctypes
        c4 = (ctypes.c_ulonglong * 4)()
        eng._lib.mlp_debug_loc_counters(eng._ctx, c4)
        cells = sum(len(a) * len(b) for i, a in enumerate(seqs) for b in seqs[i + 1:]) * 2   # two repetitions
        print("forward chain: candidates %.3f of the cells, firing %.3f; backward chain: candidates %.3f, firing %.3f" % (c4[0] / cells, c4[1] / cells, c4[2] / cells, c4[3] / cells))
    print("MLP_OLD_SWEEP=%d posterior device %.1f ms" % (old, st["ms_total"]), {k: round(v, 2) for k, v in st["ms_kernel"].items() if v}, "cells", out[old][1], flush=True)
    eng.close()
print("distances equal:", np.array_equal(out[32][0], out[0][0]), "cells equal:", out[32][1] == out[0][1])
