"""The realignment regime (BASELINE config #4): the short region files MLProbs' driver hands to `quickprobs`
(tests/golden/regions, captured by oracle/gen_region_golden.py from runs of the unmodified driver).  CPU check of the FLOW
quickprobs_b200 executes (csrc/quickprobs_main.cpp): the oracle stands in for the device stages, the tree and the tail are the
product's host code; the FASTA text must hash to what the reference `quickprobs` printed.  The device stages on inputs this
small are covered by the ragged-length GPU tests; tools/region_parity.py runs the executable itself over the same files."""
import hashlib
import io
import json
import os
import tarfile
import numpy as np
import pytest
import mlprobs_b200 as M
import oracle_lib as O
from common import HERE, tail_from_csrset

REG = os.path.join(HERE, "golden", "regions")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REG, "manifest.json")), reason="no region fixtures")


def _regions():
    man = json.load(open(os.path.join(REG, "manifest.json")))["regions"]
    data = {}
    with tarfile.open(os.path.join(REG, "inputs.tar.gz")) as tar:
        for ti in tar.getmembers():
            data[ti.name] = tar.extractfile(ti).read().decode()
    return man, data


def _fasta(headers, rows):
    out = io.StringIO()
    for h, r in zip(headers, rows):
        out.write(">" + h + "\n")
        for p in range(0, len(r), 60):
            out.write(r[p:p + 60] + "\n")
    return out.getvalue().encode()


def test_region_files_flow_matches_the_reference_quickprobs():
    man, data = _regions()
    assert len(man) >= 20
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    checked = 0
    for m in sorted(man, key=lambda e: e["nseq"] * e["max_len"]):                # all of them: seconds on the CPU
        if m["sha"] is None:
            continue
        lines = data[m["name"]].split("\n")
        headers = [l[1:].strip() for l in lines if l.startswith(">")]
        seqs = [l.strip().upper().encode() for l in lines if l and not l.startswith(">")]
        assert len(headers) == len(seqs) == m["nseq"]
        n = len(seqs)
        if n == 1:
            rows = [seqs[0].decode()]
        else:
            dist, S, rc = O.posterior_stage(O.QP, 3, ht, pt, seqs, threads=2)
            assert rc == 0
            t = M.qp_guide_tree_ex(dist)
            w = np.maximum(t["weights"], np.float32(1e-6))
            iters = 1 if n > 50 else 2
            for it in range(iters):
                S = O.relax_qp(S, w, t["seldist"], float(np.float32(1e-5)) if it == iters - 1 else float(np.float32(0.01)), 200.0, 3.0)
            rows = [r.decode() for r in tail_from_csrset(S, seqs, dist)]
        assert hashlib.sha256(_fasta(headers, rows)).hexdigest() == m["sha"], m["name"]
        checked += 1
    assert checked == sum(1 for m in man if m["sha"] is not None)
