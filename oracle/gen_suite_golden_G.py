#!/usr/bin/env python3
"""Adds the reference's `c_p_np_aln -G <file>` output (the feature line MLProbs' first classifier reads,
utils/prepare_features_4_classifier_1.py:12-24) to tests/golden/suites/manifest.json: key `cpnpG_sha` = SHA-256 of stdout,
`cpnpG` = the line itself.  Build container only.  The bundled binary always uses every core (MSA.cpp:146-151 overrides
OMP_NUM_THREADS) and its line then varies from run to run (the fifth field most), so the line comes from the harness around
the unmodified sources with the team pinned to one thread: oracle/_ref/ref_cpnp msa <fasta> - --G --threads 1.  Families
whose `-p 0` reference run took more than --max-seconds are skipped.  Usage: gen_suite_golden_G.py [--max-seconds S] [--workers W]"""
import os, json, hashlib, subprocess, time, argparse
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/TEST"
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
MAN = os.path.join(ROOT, "tests", "golden", "suites", "manifest.json")


def run_one(m):
    env = dict(os.environ); env["OMP_NUM_THREADS"] = "1"
    try:
        r = subprocess.run([EXE, "msa", os.path.join(REF, m["suite"], "in", m["name"]), "-", "--G", "--threads", "1"], capture_output=True, env=env, timeout=600)
    except subprocess.TimeoutExpired:
        return (m["suite"], m["name"]), None, None
    ok = r.returncode == 0 and r.stdout.strip()
    return (m["suite"], m["name"]), (hashlib.sha256(r.stdout).hexdigest() if ok else None), (r.stdout.decode().strip() if ok else None)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-seconds", type=float, default=5.0)
    ap.add_argument("--workers", type=int, default=6)
    a = ap.parse_args()
    man = json.load(open(MAN))
    todo = [m for m in man["families"] if m.get("cpnp_s") is not None and m["cpnp_s"] <= a.max_seconds]
    t0 = time.time()
    with ThreadPoolExecutor(a.workers) as ex:
        res = {k: (sha, line) for k, sha, line in ex.map(run_one, todo)}
    for m in man["families"]:
        k = (m["suite"], m["name"])
        if k in res:
            m["cpnpG_sha"], m["cpnpG"] = res[k]
            # exact = only the 20 standard letters occur (otherwise fields 5-6 of the reference's line are out-of-bounds reads)
            letters = set(ch.upper() for line in open(os.path.join(REF, m["suite"], "in", m["name"])) if not line.startswith(">") for ch in line if ch.isalpha())
            m["cpnpG_exact"] = bool(letters <= set("ARNDCQEGHILKMFPSTWYV"))
    json.dump(man, open(MAN, "w"), indent=0)
    print("done: %d families, %d without a line, %.0f s" % (len(res), sum(1 for v in res.values() if v[0] is None), time.time() - t0))
