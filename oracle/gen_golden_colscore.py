#!/usr/bin/env python3
"""Golden vectors for mlp_column_scores: runs the REFERENCE's own Python function (utils/calculate_column_scores.py,
calculateColScore) on the reference alignments stored in the QuickProbs / c_p_np_aln fixtures.  Build container only."""
import os, sys, io, contextlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/utils")
import calculate_column_scores as ref   # pure Python, importable as is

out = {}
for name in ("qp_sup139", "qp_sup002", "qp_676s4", "cpnp_676s4_ref"):
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    rows = [r.tobytes().decode() for r in d["msa"]]
    text = "\n".join(">s%04d\n%s" % (i, r) for i, r in enumerate(rows))
    with contextlib.redirect_stdout(io.StringIO()):
        _, col, mean, lens, nseq, sd, ratio = ref.calculateColScore(text)
    assert lens == len(rows[0]) and nseq == len(rows)
    out[name + ".col"] = np.array(col, np.float64)
    out[name + ".stats"] = np.array([mean, sd, ratio], np.float64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "colscore.npz"), **out)
print("wrote colscore.npz", {k: v.shape for k, v in out.items()})
