#!/usr/bin/env python3
"""Pins the benchmark families oracle/gen_suite_golden.py's cell budget left out (the largest bali3 and oxx families).
Build container only (needs /root/reference and oracle/_ref).

Every run appends one JSON line per (family, program) to tests/golden/suites/rest_runs.jsonl as soon as it finishes, so a
multi-hour job can be interrupted and resumed; `--merge` folds the lines into manifest.json and adds the inputs to
inputs_rest.tar.gz.  Programs, cheapest first: quickprobs (prebuilt reference binary, -t 1), c_p_np_aln -G and
c_p_np_aln -p 0 (oracle/_ref/ref_cpnp msa, one OpenMP thread), then c_p_np_aln -p 1 with time() pinned to the manifest's
p1_fixtime.  All jobs go through one worker pool in that order and each has its own timeout;
a timed-out job is recorded as such and is not retried.
Usage: gen_suite_golden_rest.py [--workers W] [--phases qp,G,p0,p1] [--timeout S] [--merge]"""
import os, sys, json, hashlib, subprocess, tarfile, tempfile, time, argparse, threading
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/TEST"
QP = "/root/reference/realign/QuickProbs/bin/quickprobs"
CPNP = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
OUT = os.path.join(ROOT, "tests", "golden", "suites")
MAN = os.path.join(OUT, "manifest.json")
LOG = os.path.join(OUT, "rest_runs.jsonl")
lock = threading.Lock()


def stats(path):
    n = 0; L = 0
    for line in open(path):
        if line.startswith(">"): n += 1
        else: L += len(line.strip())
    return n, L / max(n, 1)


def sha_file(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest() if os.path.exists(p) and os.path.getsize(p) else None


def run_one(job):
    suite, name, path, phase, tmp, timeout, fixtime = job
    t0 = time.time()
    rec = {"suite": suite, "name": name, "phase": phase}
    env = dict(os.environ); env["OMP_NUM_THREADS"] = "1"
    try:
        if phase == "qp":
            r = subprocess.run([QP, path, "-t", "1"], capture_output=True, timeout=timeout, env=env)
            rec["rc"] = r.returncode
            rec["sha"] = hashlib.sha256(r.stdout).hexdigest() if r.returncode == 0 and r.stdout else None
        elif phase == "G":
            r = subprocess.run([CPNP, "msa", path, "-", "--G", "--threads", "1"], capture_output=True, timeout=timeout, env=env)
            ok = r.returncode == 0 and r.stdout.strip()
            rec["rc"] = r.returncode
            rec["sha"] = hashlib.sha256(r.stdout).hexdigest() if ok else None
            rec["line"] = r.stdout.decode().strip() if ok else None
        else:
            out = os.path.join(tmp, "%s_%s.%s" % (suite, name, phase))
            args = [CPNP, "msa", path, out, "--threads", "1"] + (["--p1", "--fixtime", str(fixtime)] if phase == "p1" else [])
            r = subprocess.run(args, capture_output=True, timeout=timeout, env=env)
            rec["rc"] = r.returncode
            rec["sha"] = sha_file(out) if r.returncode == 0 else None
            if os.path.exists(out): os.remove(out)
    except subprocess.TimeoutExpired:
        rec["rc"] = None; rec["sha"] = None; rec["timeout"] = timeout
    rec["s"] = round(time.time() - t0, 2)
    with lock:
        with open(LOG, "a") as f:
            f.write(json.dumps(rec) + "\n")
    return rec


def missing_families(man):
    have = {(m["suite"], m["name"]) for m in man["families"]}
    fams = []
    for suite in ("bali3", "ox", "oxx", "sabre"):
        for f in sorted(os.listdir(os.path.join(REF, suite, "in"))):
            p = os.path.join(REF, suite, "in", f)
            n, L = stats(p)
            if n >= 2 and (suite, f) not in have:
                fams.append((suite, f, p, n, L))
    return fams


def merge():
    man = json.load(open(MAN))
    by = {(m["suite"], m["name"]): m for m in man["families"]}
    key = {"qp": "qp", "p0": "cpnp", "p1": "cpnp1", "G": "cpnpG"}
    recs = [json.loads(l) for l in open(LOG)] if os.path.exists(LOG) else []
    for r in recs:
        k = (r["suite"], r["name"])
        m = by.get(k)
        if m is None:
            m = by[k] = {"suite": r["suite"], "name": r["name"], "rest": True}
            man["families"].append(m)
        p = key[r["phase"]]
        m[p + "_sha"] = r["sha"]
        m[p + "_s"] = r["s"]
        if r["phase"] in ("qp", "p0"):
            m[p + "_rc"] = r["rc"]
        if r.get("timeout"):
            m[p + "_timeout"] = r["timeout"]
        if r["phase"] == "G":
            m["cpnpG"] = r.get("line")
            path = os.path.join(REF, r["suite"], "in", r["name"])
            letters = set(ch.upper() for line in open(path) if not line.startswith(">") for ch in line if ch.isalpha())
            m["cpnpG_exact"] = bool(letters <= set("ARNDCQEGHILKMFPSTWYV"))
    rest = [m for m in man["families"] if m.get("rest")]
    # a family joins the manifest as soon as one program has a reference answer (or a recorded failure); tools/suite_parity.py
    # runs every tool on the families that have that tool's key
    man["families"] = [m for m in man["families"] if not m.get("rest") or ("qp_sha" in m or "cpnp_sha" in m)]
    with tarfile.open(os.path.join(OUT, "inputs_rest.tar.gz"), "w:gz") as tar:
        for m in man["families"]:
            if m.get("rest"):
                tar.add(os.path.join(REF, m["suite"], "in", m["name"]), arcname="%s/%s" % (m["suite"], m["name"]))
    json.dump(man, open(MAN, "w"), indent=0)
    n_rest = sum(1 for m in man["families"] if m.get("rest"))
    print("manifest: %d families (%d from this script, %d still without both main programs)" % (len(man["families"]), n_rest, len(rest) - n_rest))
    for p in ("qp", "cpnp", "cpnp1", "cpnpG"):
        print("  %-6s pinned on %d" % (p, sum(1 for m in man["families"] if m.get(p + "_sha"))))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=6)
    ap.add_argument("--phases", default="qp,G,p0,p1")
    ap.add_argument("--timeout", type=float, default=4 * 3600)
    ap.add_argument("--p1-timeout", type=float, default=3600)
    ap.add_argument("--merge", action="store_true")
    a = ap.parse_args()
    if a.merge:
        merge(); sys.exit(0)
    man = json.load(open(MAN))
    fixtime = man.get("p1_fixtime", 777)
    done = set()
    if os.path.exists(LOG):
        for l in open(LOG):
            r = json.loads(l); done.add((r["suite"], r["name"], r["phase"]))
    fams = missing_families(man)
    # already-merged "rest" families may still lack phases
    for m in man["families"]:
        if m.get("rest"):
            p = os.path.join(REF, m["suite"], "in", m["name"]); n, L = stats(p)
            fams.append((m["suite"], m["name"], p, n, L))
    tmp = tempfile.mkdtemp()
    cost = lambda x: x[3] ** 2 * x[4] ** 2 + x[3] ** 3 * x[4] * 0.3
    jobs = []
    for phase in a.phases.split(","):
        # one queue, no barrier between the phases: quickprobs and -G first (cheap), -p 0 largest first (packs the pool),
        # -p 1 smallest first (the reference's alignment graph is the slow part; whatever fits the time gets pinned)
        order = sorted(fams, key=cost, reverse=(phase == "p0"))
        jobs += [(s, f, p, phase, tmp, a.p1_timeout if phase == "p1" else a.timeout, fixtime) for s, f, p, n, L in order if (s, f, phase) not in done]
    print("jobs:", len(jobs), flush=True)
    t0 = time.time()
    with ThreadPoolExecutor(a.workers) as ex:
        res = list(ex.map(run_one, jobs))
    print("done in %.0f s, failed/timed out: %d" % (time.time() - t0, sum(1 for r in res if r["sha"] is None)), flush=True)
