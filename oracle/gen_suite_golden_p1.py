#!/usr/bin/env python3
"""Adds the reference's `c_p_np_aln -p 1` outputs to tests/golden/suites/manifest.json (key `cpnp1_sha`, SHA-256 of the FASTA).
Build container only.  The program is run as oracle/_ref/ref_cpnp msa --p1 --threads 1 --fixtime 777: one OpenMP thread
(its refinement races on the shared posterior otherwise) and the harness' pinned time() instead of the wall clock the
program seeds every refinement sweep with.  Only families whose `-p 0` reference run took at most --max-seconds are run
(the alignment graph of the reference copies its whole child table per candidate cell and is slow on large families).
Usage: gen_suite_golden_p1.py [--min-seconds S] [--max-seconds S] [--workers W] [--timeout S]"""
import os, sys, json, hashlib, subprocess, tempfile, time, argparse
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/TEST"
CPNP = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
MAN = os.path.join(ROOT, "tests", "golden", "suites", "manifest.json")
FIXTIME = 777
TIMEOUT = 900


def run_one(job):
    m, tmp = job
    path = os.path.join(REF, m["suite"], "in", m["name"])
    out = os.path.join(tmp, "%s_%s.p1" % (m["suite"], m["name"]))
    t0 = time.time()
    try:
        c = subprocess.run([CPNP, "msa", path, out, "--p1", "--threads", "1", "--fixtime", str(FIXTIME)], capture_output=True, timeout=TIMEOUT)
        ok = c.returncode == 0 and os.path.exists(out) and os.path.getsize(out) > 0
    except subprocess.TimeoutExpired:
        ok = False
    return (m["suite"], m["name"]), (hashlib.sha256(open(out, "rb").read()).hexdigest() if ok else None), round(time.time() - t0, 2)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-seconds", type=float, default=5.0)
    ap.add_argument("--workers", type=int, default=6)
    ap.add_argument("--min-seconds", type=float, default=0.0)
    ap.add_argument("--timeout", type=int, default=900)
    a = ap.parse_args()
    TIMEOUT = a.timeout
    man = json.load(open(MAN))
    tmp = tempfile.mkdtemp()
    todo = [(m, tmp) for m in man["families"] if m.get("cpnp_s") is not None and m["cpnp_sha"] and a.min_seconds <= m["cpnp_s"] <= a.max_seconds and "cpnp1_sha" not in m]
    todo.sort(key=lambda j: -j[0]["cpnp_s"])
    print("to run:", len(todo), flush=True)
    t0 = time.time()
    with ThreadPoolExecutor(a.workers) as ex:
        res = {k: (sha, s) for k, sha, s in ex.map(run_one, todo)}
    for m in man["families"]:
        k = (m["suite"], m["name"])
        if k in res:
            m["cpnp1_sha"], m["cpnp1_s"] = res[k]
    man["note"] = ("sha256 of the reference outputs: quickprobs <file> (stdout), c_p_np_aln -p 0 on one OpenMP thread, and (cpnp1_sha, where present) "
                   "c_p_np_aln -p 1 on one thread with time() pinned to %d" % FIXTIME)
    man["p1_fixtime"] = FIXTIME
    json.dump(man, open(MAN, "w"), indent=0)
    print("done: %d families, %d failed, %.0f s" % (len(res), sum(1 for v in res.values() if v[0] is None), time.time() - t0))
