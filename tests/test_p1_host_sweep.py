"""`c_p_np_aln -p 1` FASTA to FASTA without a GPU on a sample of the bundled benchmark families: the oracle stands in for the device stages, the
alignment graph and the refinement are the product's host code, and the text must hash to what the reference program wrote with its clock
pinned (tests/golden/suites/manifest.json, `cpnp1_sha`).  tools/p1_host_sweep.py runs every pinned family this way
(profiles/r2c_p1_host_sweep.txt); the device flow is compared by tests/test_suites_gpu.py and tools/suite_parity.py."""
import json
import os
import sys
import tarfile
import pytest
from common import HERE

sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
SUITES = os.path.join(HERE, "golden", "suites")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(SUITES, "manifest.json")), reason="no suite fixtures")


def test_sample_of_pinned_families_is_byte_identical(tmp_path):
    import p1_host_sweep as P
    manifest = json.load(open(os.path.join(SUITES, "manifest.json")))
    seed = int(manifest.get("p1_fixtime", 777))
    fams = [m for m in manifest["families"] if m.get("cpnp1_sha") and 0.3 <= float(m.get("cpnp1_s") or 0) <= 1.5]
    sample = []
    for suite in ("bali3", "ox", "oxx", "sabre"):                       # three per suite, spread over the list
        s = [m for m in fams if m["suite"] == suite]
        sample += [s[k] for k in sorted({0, len(s) // 2, len(s) - 1})] if s else []
    assert len(sample) >= 8
    want = {m["suite"] + "/" + m["name"] for m in sample}
    for arc in ("inputs.tar.gz", "inputs_rest.tar.gz"):
        p = os.path.join(SUITES, arc)
        if os.path.exists(p):
            with tarfile.open(p) as tar:
                tar.extractall(tmp_path, members=[ti for ti in tar.getmembers() if ti.name in want], filter="data")
    for m in sample:
        _, n, verdict, _, _ = P.one((str(tmp_path / m["suite"] / m["name"]), m, seed))
        assert verdict == "MATCH", (m["suite"], m["name"], n, verdict)
