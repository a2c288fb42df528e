"""MLProbs' unmodified Python driver on top of tools/mlprobs_overlay (working directory assembled from symlinks + a joblib
stand-in that evaluates the scikit-learn 0.21.3 random forests without scikit-learn).  Build container only: it needs the
MLProbs checkout for MLProbs.py, utils/, the classifier pickles and -- for this CPU check -- the checkout's own CPU programs.
With `--binaries b200` the same tool puts this repository's executables at the two relative paths the driver hard-codes."""
import os
import sys
import numpy as np
import pytest
from common import HERE

MLPROBS = "/root/reference"
OVERLAY = os.path.join(os.path.dirname(HERE), "tools", "mlprobs_overlay")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(MLPROBS, "MLProbs.py")), reason="no MLProbs checkout here")


def fasta(path):
    out = {}
    name = None
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            name = line; out[name] = ""
        elif line:
            out[name] += line
    return out


@pytest.mark.parametrize("suite,name", [("bali3", "BB11001"), ("sabre", "sup_139"), ("ox", "104s10"), ("sabre", "sup_200")])
def test_driver_runs_and_reproduces_the_published_alignment(tmp_path, monkeypatch, suite, name):
    """Families the first classifier sends to the progressive strategy (deterministic on one thread): the driver's final
    alignment has the rows of the checkout's published result (output4evaluation/; the row order of c_p_np_aln's refinement is
    not stable there, so rows are compared by header).  sup_200 takes the `Realign Incredible Regions` branch, which also
    exercises the second and third classifier."""
    import shutil
    if not shutil.which("taskset"):
        pytest.skip("taskset is needed to make the reference programs repeatable")
    sys.path.insert(0, OVERLAY)
    try:
        import run as overlay_run
    finally:
        sys.path.remove(OVERLAY)
    out = tmp_path / "out.msa"
    # pinned to one core: the reference c_p_np_aln takes every core it may use and is not repeatable with more than one
    rc, log = overlay_run.run(MLPROBS, "reference", os.path.join(MLPROBS, "TEST", suite, "in", name), str(out), quiet=True, one_core=True)
    assert rc == 0 and "Got the final MSA" in log
    assert fasta(out) == fasta(os.path.join(MLPROBS, "output4evaluation", suite, name))


def test_forest_shim_loads_all_three_classifiers():
    """The stand-in reads the joblib files (100 trees each) and votes like RandomForestClassifier.predict: mean of the per-tree
    leaf distributions, first maximum."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mlprobs_overlay_joblib", os.path.join(OVERLAY, "joblib.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    for sub, nfeat, nclass in (("branch", 5, 2), ("regions", 4, 2), ("seq_lens", 5, 4)):
        f = shim.load(os.path.join(MLPROBS, "classifier", "model", sub, "randomforest.joblib"))
        assert len(f.trees) == 100 and len(f.classes_) == nclass
        x = np.linspace(0.1, 0.9, nfeat)[None, :]
        p = f.predict_proba(x)
        assert p.shape == (1, nclass) and abs(p.sum() - 1) < 1e-9
        assert f.predict(x)[0] == f.classes_[int(np.argmax(p[0]))]
