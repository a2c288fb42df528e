import sys, time, numpy as np
sys.path.insert(0, '.')
import mlprobs_b200 as M
from mlprobs_b200 import synth
for n in (1000, 2000, 4000):
    seqs = synth.family_fast(n, 12, seed=3)
    eng = M.Engine(0); h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(seqs)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    rng = np.random.default_rng(n)
    d = rng.random((n, n)).astype(np.float32); d = np.triu(d, 1); d = d + d.T
    eng.debug_set_distances(d)
    for rep in range(3):
        t0 = time.perf_counter(); dev = eng.qp_guide_tree_device(1e-6); t1 = time.perf_counter()
    dd = eng.distances()
    t2 = time.perf_counter(); host = M.qp_guide_tree_ex(dd); t3 = time.perf_counter()
    print("n=%d device tree %.2f ms (weights + children to the host), host tree %.2f ms (+ %.2f ms distances d2h), equal: %s" % (n, (t1 - t0) * 1e3, (t3 - t2) * 1e3, 0.0, np.array_equal(host["left"], dev["left"]) and np.array_equal(np.maximum(host["weights"], np.float32(1e-6)), dev["weights"])), flush=True)
    eng.close()
