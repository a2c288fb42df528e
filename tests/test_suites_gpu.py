"""Suite-level parity of the two drop-in executables: byte-identical FASTA output on bundled benchmark families
(tests/golden/suites, written by oracle/gen_suite_golden*.py from the reference programs).  The ox and oxx suites run here
(790 families, quickprobs and c_p_np_aln -p 0, about two and a half minutes on a B200) plus c_p_np_aln -p 1 on ox, the -G line and the
realignment regions; tools/suite_parity.py runs all four suites (1,599 families; profiles/r2b_suite_parity_all_1599.txt)."""
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
    rep = suite_parity.run(None, suites=("ox",), tools=("cpnp1_sha",), gpu_verified_only=True)
    assert rep["suites"], "no -p 1 reference outputs in the manifest"
    assert rep["mismatches"] == [] and rep["failures"] == []
    for key, v in rep["suites"].items():
        assert v["identical"] == v["families"] - v["reference_failed"] > 0, (key, v)


def test_feature_line_G_is_byte_identical_on_the_standard_letter_families():
    """c_p_np_aln -G (the line classifier 1 consumes, MSA.cpp:646-762) against the one-thread reference harness, ox + oxx."""
    import suite_parity
    rep = suite_parity.run(None, suites=("ox", "oxx"), tools=("cpnpG_sha",))
    assert rep["suites"], "no -G reference lines in the manifest"
    assert rep["mismatches"] == [] and rep["failures"] == []
    for key, v in rep["suites"].items():
        assert v["identical"] == v["families"] - v["reference_failed"] > 0, (key, v)


def test_realignment_regions_are_byte_identical_to_the_reference_quickprobs():
    """BASELINE config #4: the 112 region files MLProbs' unmodified driver handed to quickprobs (utils/do_realign.py:52-63),
    realigned by quickprobs_b200 in directory mode (one CUDA context)."""
    import region_parity
    rep = region_parity.run(None)
    assert rep["different"] == [] and rep["no_output"] == [], rep
    assert rep["identical"] == rep["regions"] - rep["reference_failed"] > 0
