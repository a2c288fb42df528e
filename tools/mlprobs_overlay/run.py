#!/usr/bin/env python3
"""Run MLProbs' UNMODIFIED Python driver (MLProbs.py + utils/, SURVEY 8b: it calls `./baseMSA/C_P_NP_Aln/c_p_np_aln` and
`./realign/QuickProbs/bin/quickprobs` relative to its working directory) on top of a chosen pair of executables.

A scratch working directory is assembled from symlinks: MLProbs.py, utils/ and classifier/ of the MLProbs checkout, and the
two executables at the relative paths the driver hard-codes -- either this repository's drop-ins (`--binaries b200`, needs a
GPU) or the checkout's own CPU programs (`--binaries reference`, for comparison).  Nothing of the checkout is copied or
modified.  PYTHONPATH puts tools/mlprobs_overlay first so that `from joblib import load` resolves to the shim in joblib.py
(the classifiers were pickled with scikit-learn 0.21.3, which a current scikit-learn cannot read).

usage: run.py --mlprobs /path/to/MLProbs [--binaries b200|reference] [--seed S] <family.fasta> <out.msa>"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def assemble(mlprobs, binaries):
    work = tempfile.mkdtemp(prefix="mlprobs_overlay_")
    for name in ("MLProbs.py", "utils", "classifier"):
        os.symlink(os.path.join(mlprobs, name), os.path.join(work, name))
    os.makedirs(os.path.join(work, "baseMSA", "C_P_NP_Aln"))
    os.makedirs(os.path.join(work, "realign", "QuickProbs", "bin"))
    if binaries == "b200":
        cpnp = os.path.join(ROOT, "mlprobs_b200", "bin", "c_p_np_aln_b200")
        qp = os.path.join(ROOT, "mlprobs_b200", "bin", "quickprobs_b200")
    else:
        cpnp = os.path.join(mlprobs, "baseMSA", "C_P_NP_Aln", "c_p_np_aln")
        qp = os.path.join(mlprobs, "realign", "QuickProbs", "bin", "quickprobs")
    for exe in (cpnp, qp):
        if not os.path.exists(exe):
            raise SystemExit("missing executable: " + exe)
    os.symlink(cpnp, os.path.join(work, "baseMSA", "C_P_NP_Aln", "c_p_np_aln"))
    os.symlink(qp, os.path.join(work, "realign", "QuickProbs", "bin", "quickprobs"))
    return work


def run(mlprobs, binaries, fasta, out, seed=None, keep=False, quiet=False, one_core=False, server=False):
    work = assemble(os.path.abspath(mlprobs), binaries)
    env = dict(os.environ)
    env["PYTHONPATH"] = HERE + os.pathsep + env.get("PYTHONPATH", "")
    if server:
        env["MLP_B200_SERVER"] = "1"              # persistent-process mode of the drop-ins (csrc/serve.h): one CUDA context for all the driver's calls
    if seed is not None:
        env["MLP_CPNP_SEED"] = str(seed)          # c_p_np_aln_b200 -p 1: stands in for the clock the reference seeds with
    try:
        # the driver passes file names through a shell unquoted: use plain absolute paths
        # one_core: the reference c_p_np_aln sizes its OpenMP team from the cores it may run on (it overrides OMP_NUM_THREADS)
        # and both its -G line and its refinement change from run to run with more than one thread; taskset makes it repeatable
        pin = ["taskset", "-c", "0"] if one_core and shutil.which("taskset") else []
        r = subprocess.run(pin + [sys.executable, "MLProbs.py", os.path.abspath(fasta), os.path.abspath(out)], cwd=work, env=env,
                           capture_output=quiet, text=True)
        return r.returncode, (r.stdout if quiet else "")
    finally:
        if not keep:
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mlprobs", required=True)
    ap.add_argument("--binaries", choices=("b200", "reference"), default="b200")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--server", action="store_true", help="b200 binaries in persistent-process mode (MLP_B200_SERVER=1)")
    ap.add_argument("--one-core", action="store_true", help="pin the run to one core (repeatable output with the reference programs)")
    ap.add_argument("fasta")
    ap.add_argument("out")
    a = ap.parse_args()
    rc, _ = run(a.mlprobs, a.binaries, a.fasta, a.out, a.seed, a.keep, one_core=a.one_core, server=a.server)
    sys.exit(rc)
