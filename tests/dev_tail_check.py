#!/usr/bin/env python3
"""Developer check (CPU): host tail fed by the ORACLE's sparse set vs the reference harness' final MSA."""
import os, sys, subprocess, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import mlprobs_b200 as M
from mlprobs_b200 import synth
import oracle_lib as O


def read_fasta(path):
    hs, ss = [], []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"): hs.append(line[1:]); ss.append("")
        elif line: ss[-1] += line
    return hs, ss


def oracle_tail(seqs, iters=None, ref_iters=-1, threads=8):
    if iters is None:
        iters = 2 if len(seqs) <= 50 else 1          # ConsistencyStage.cpp:73-76, Configuration.cpp:100-103
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    dist, s, rc = O.posterior_stage(0, 3, ht, pt, seqs, 0.01, threads)
    assert rc == 0
    t = M.qp_guide_tree_ex(dist)
    w = np.maximum(t["weights"], np.float32(1e-6))
    for it in range(iters):
        s = O.relax_qp(s, w, t["seldist"], float(np.float32(1e-5)) if it == iters - 1 else 0.01, threads=threads)
    cells = np.zeros(len(s.col), dtype=[("c", np.int32), ("v", np.float32)])
    cells["c"] = s.col; cells["v"] = s.val
    return M.qp_finish_alignment_host(seqs, w, t["left"], t["right"], s.rp_off, s.nz_off, s.rowptr, cells, ref_iters)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    seqs = synth.family(n, L, seed=seed)
    seqs = [x if isinstance(x, bytes) else x.encode() for x in seqs]
    d = tempfile.mkdtemp()
    fa = os.path.join(d, "in.fa")
    with open(fa, "w") as f:
        for i, x in enumerate(seqs): f.write(">s%d\n%s\n" % (i, x.decode()))
    out = os.path.join(d, "ref.fa")
    r = subprocess.run([os.path.join(HERE, "..", "oracle", "_ref", "ref_qp"), "msa", fa, out, "--threads", "4"], capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1])
    _, ref = read_fasta(out)
    mine = [x.decode() for x in oracle_tail(seqs)]
    same = mine == ref
    print("n=%d L=%d  ref len %d mine len %d  identical=%s" % (n, L, len(ref[0]), len(mine[0]), same))
    if not same:
        _, rc = read_fasta(out + ".construct")
        print("construct-only ref len", len(rc[0]))
        for i in range(n):
            if mine[i] != ref[i]: print(i, "\n ", mine[i], "\n ", ref[i]); break
    sys.exit(0 if same else 1)
