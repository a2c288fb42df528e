#!/usr/bin/env python3
"""Suite-level golden data: the reference programs' final alignments on the bundled benchmark families (TEST/bali3, ox,
oxx, sabre).  Build container only (needs /root/reference and oracle/_ref).

For every selected family:  quickprobs (the reference's prebuilt binary; its output does not depend on the thread count)
and c_p_np_aln -p 0 (oracle/_ref/ref_cpnp msa = the unmodified sources' whole-program flow on ONE OpenMP thread, the only
deterministic setting of that program) are run; the SHA-256 of each output goes into tests/golden/suites/manifest.json and
the input files into tests/golden/suites/inputs.tar.gz, so that the GPU box can check the drop-in executables without the
reference tree.  Usage: gen_suite_golden.py [--budget CELLS] (families are taken smallest first per suite)."""
import os, sys, json, hashlib, subprocess, tarfile, io, time, argparse
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/TEST"
QP = "/root/reference/realign/QuickProbs/bin/quickprobs"
CPNP = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
OUT = os.path.join(ROOT, "tests", "golden", "suites")


def stats(path):
    n = 0; L = 0
    for line in open(path):
        if line.startswith(">"): n += 1
        else: L += len(line.strip())
    return n, L / max(n, 1)


def run_one(job):
    suite, name, path, tmp = job
    res = {"suite": suite, "name": name}
    t0 = time.time()
    q = subprocess.run([QP, path, "-t", "1"], capture_output=True)
    res["qp_rc"] = q.returncode
    res["qp_sha"] = hashlib.sha256(q.stdout).hexdigest() if q.returncode == 0 and q.stdout else None
    res["qp_s"] = round(time.time() - t0, 2)
    t0 = time.time()
    out = os.path.join(tmp, "%s_%s.cpnp" % (suite, name))
    c = subprocess.run([CPNP, "msa", path, out, "--threads", "1"], capture_output=True)
    res["cpnp_rc"] = c.returncode
    res["cpnp_sha"] = hashlib.sha256(open(out, "rb").read()).hexdigest() if c.returncode == 0 and os.path.exists(out) and os.path.getsize(out) else None
    res["cpnp_s"] = round(time.time() - t0, 2)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, nargs=4, default=[3e9, 1e12, 1.5e9, 1e12], help="N^2/2*L^2 cell budget per suite: bali3 ox oxx sabre")
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--append", action="store_true", help="keep the families already in the manifest, run only the new ones")
    a = ap.parse_args()
    import tempfile
    tmp = tempfile.mkdtemp()
    jobs = []
    have = {}
    if a.append and os.path.exists(os.path.join(OUT, "manifest.json")):
        for m in json.load(open(os.path.join(OUT, "manifest.json")))["families"]:
            have[(m["suite"], m["name"])] = m
    for suite, budget in zip(("bali3", "ox", "oxx", "sabre"), a.budget):
        fams = []
        for f in sorted(os.listdir(os.path.join(REF, suite, "in"))):
            p = os.path.join(REF, suite, "in", f)
            n, L = stats(p)
            if n >= 2: fams.append((n * n / 2 * L * L, f, p))
        fams.sort()
        used = 0
        for cost, f, p in fams:
            if used + cost > budget: break
            used += cost
            jobs.append((suite, f, p, tmp))
        print(suite, "selected", sum(1 for j in jobs if j[0] == suite), "of", len(fams), "families, %.2e cells" % used, flush=True)
    t0 = time.time()
    todo = [j for j in jobs if (j[0], j[1]) not in have]
    todo.sort(key=lambda j: -os.path.getsize(j[2]))           # big families first: better packing of the worker pool
    print("to run:", len(todo), "already have:", len(jobs) - len(todo), flush=True)
    with ThreadPoolExecutor(a.workers) as ex:
        fresh = {(r["suite"], r["name"]): r for r in ex.map(run_one, todo)}
    results = [have.get((j[0], j[1])) or fresh[(j[0], j[1])] for j in jobs]
    print("reference runs: %.0f s" % (time.time() - t0))
    os.makedirs(OUT, exist_ok=True)
    with tarfile.open(os.path.join(OUT, "inputs.tar.gz"), "w:gz") as tar:
        for suite, f, p, _ in jobs:
            tar.add(p, arcname="%s/%s" % (suite, f))
    json.dump({"families": results, "note": "sha256 of the reference outputs: quickprobs <file> (stdout) and c_p_np_aln -p 0 on one OpenMP thread"},
              open(os.path.join(OUT, "manifest.json"), "w"), indent=0)
    bad = [r for r in results if r["qp_sha"] is None or r["cpnp_sha"] is None]
    print("families:", len(results), "reference failures:", len(bad), [(r["suite"], r["name"], r["qp_rc"], r["cpnp_rc"]) for r in bad][:10])
    print("sizes:", os.path.getsize(os.path.join(OUT, "inputs.tar.gz")), os.path.getsize(os.path.join(OUT, "manifest.json")))
