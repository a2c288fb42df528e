"""Suite-level parity of the two drop-in executables: byte-identical FASTA output on bundled benchmark families
(tests/golden/suites, written by oracle/gen_suite_golden.py from the reference programs).  The ox and oxx suites run here
(546 families, quickprobs and c_p_np_aln -p 0, about a minute and a half) plus c_p_np_aln -p 1 on ox; tools/suite_parity.py
runs all four suites (1276 families)."""
import os
import sys
import pytest
from common import HERE

sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
pytestmark = pytest.mark.gpu


def test_ox_and_oxx_families_are_byte_identical_to_the_reference_programs():
    import suite_parity
    rep = suite_parity.run(None, suites=("ox", "oxx"), tools=("qp_sha", "cpnp_sha"))
    assert rep["mismatches"] == [] and rep["failures"] == []
    for key, v in rep["suites"].items():
        assert v["identical"] == v["families"] - v["reference_failed"] > 0, (key, v)


def test_ox_families_non_progressive_program_is_byte_identical():
    """c_p_np_aln -p 1 (alignment graph + similar-set refinement) with the clock value the reference outputs were pinned to."""
    import suite_parity
    rep = suite_parity.run(None, suites=("ox",), tools=("cpnp1_sha",))
    assert rep["suites"], "no -p 1 reference outputs in the manifest"
    assert rep["mismatches"] == [] and rep["failures"] == []
    for key, v in rep["suites"].items():
        assert v["identical"] == v["families"] - v["reference_failed"] > 0, (key, v)
