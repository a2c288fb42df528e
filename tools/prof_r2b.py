#!/usr/bin/env python3
"""One small pass of the kernels added in the second half of round 2, for ncu (developer script): c_p_np_aln's three models on
loc_c.cu / part_sc.cu / final_c.cu mode 7, then QuickProbs' posterior stage and the device guide tree."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlprobs_b200 as M
from mlprobs_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
seqs = synth.family_fast(n, 300, seed=20220150)
eng = M.Engine(0)
h, p = M.default_tables(M.CPNP_P0, 0.100675); eng.set_tables(h, p); eng.set_sequences(seqs)
eng.posterior_all_pairs(M.CPNP_P0, 7, 0.01)
print("cpnp three models", {k: round(v, 2) for k, v in eng.stats()["ms_kernel"].items() if v})
eng.close()
seqs = synth.family_fast(1000, 40, seed=20220150)
eng = M.Engine(0)
h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(seqs)
eng.posterior_all_pairs(M.QP, 3, 0.01)
t = eng.qp_guide_tree_device(1e-6)
print("tree ok", t["left"][-1], t["right"][-1])
eng.close()
