#!/usr/bin/env python3
"""Stage times of `c_p_np_aln_b200 -p 1` on a synthetic family next to the reference harness (`oracle/_ref/ref_cpnp msa --p1`,
all host threads; the reference's graph and refinement are serial).  Usage: p1_demo.py [n] [length] [reference timeout s]"""
import os, sys, subprocess, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mlprobs_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
tmo = int(sys.argv[3]) if len(sys.argv) > 3 else 150
seqs = synth.family(n, L, seed=11)
td = tempfile.mkdtemp()
fa = os.path.join(td, "in.fa")
with open(fa, "w") as f:
    for i, s in enumerate(seqs):
        f.write(">s%d\n%s\n" % (i, s.decode()))
exe = os.path.join(ROOT, "mlprobs_b200", "bin", "c_p_np_aln_b200")
for rep in range(2):        # the second run has a warm driver
    t0 = time.time()
    r = subprocess.run([exe, "-p", "1", "--seed", "1", "-v", "-o", os.path.join(td, "ours.fa"), fa], capture_output=True, text=True)
    print("ours rc %d wall %.2f s | %s" % (r.returncode, time.time() - t0, r.stderr.strip()), flush=True)
ref = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
if os.path.exists(ref):
    t0 = time.time()
    try:
        r = subprocess.run([ref, "msa", fa, os.path.join(td, "ref.fa"), "--p1", "--threads", str(os.cpu_count()), "--fixtime", "1"], capture_output=True, timeout=tmo)
        print("reference (%d threads) rc %d wall %.2f s" % (os.cpu_count(), r.returncode, time.time() - t0), flush=True)
    except subprocess.TimeoutExpired:
        print("reference (%d threads): not finished after %d s" % (os.cpu_count(), tmo), flush=True)
