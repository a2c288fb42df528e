#!/usr/bin/env python3
"""Golden data for the realignment regime (BASELINE config #4): the short region files MLProbs' driver feeds to `quickprobs`
(utils/do_realign.py:52-63 writes ./tmp/qp_tmp/<from>-<to>.unreliable and runs `quickprobs <file> > ...`).
Build container only.  For a spread of bundled families the UNMODIFIED driver is run through tools/mlprobs_overlay with the
checkout's own CPU programs; every region file it produced is kept together with the SHA-256 of what the reference
`quickprobs` prints for it (its output does not depend on the thread count).  Output: tests/golden/regions/inputs.tar.gz +
manifest.json, consumed by tools/region_parity.py on the GPU box.
Usage: gen_region_golden.py [--per-suite K]"""
import argparse, glob, hashlib, io, json, os, subprocess, sys, tarfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "mlprobs_overlay"))
import run as overlay
MLPROBS = "/root/reference"
QP = os.path.join(MLPROBS, "realign", "QuickProbs", "bin", "quickprobs")
OUT = os.path.join(ROOT, "tests", "golden", "regions")

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-suite", type=int, default=30)
    a = ap.parse_args()
    os.environ["OMP_NUM_THREADS"] = "1"
    os.makedirs(OUT, exist_ok=True)
    entries, blobs = [], {}
    for suite in ("bali3", "ox", "oxx", "sabre"):
        fams = sorted(os.listdir(os.path.join(MLPROBS, "TEST", suite, "in")), key=lambda f: os.path.getsize(os.path.join(MLPROBS, "TEST", suite, "in", f)))
        fams = fams[:: max(1, len(fams) // (3 * a.per_suite))][: a.per_suite]          # smaller two thirds, evenly spread
        for fam in fams:
            work = overlay.assemble(MLPROBS, "reference")
            env = dict(os.environ); env["PYTHONPATH"] = overlay.HERE
            try:
                subprocess.run([sys.executable, "MLProbs.py", os.path.join(MLPROBS, "TEST", suite, "in", fam), os.path.join(work, "o.msa")],
                               cwd=work, env=env, capture_output=True, timeout=600)
            except subprocess.TimeoutExpired:
                continue
            for f in sorted(glob.glob(os.path.join(work, "tmp", "qp_tmp", "*.unreliable"))):
                data = open(f, "rb").read()
                if not data.strip():
                    continue
                name = "%s__%s__%s" % (suite, fam, os.path.basename(f))
                r = subprocess.run([QP, f], capture_output=True)
                n = data.count(b">")
                lens = [len(x) for x in data.decode().split("\n")[1::2]]
                entries.append({"name": name, "nseq": n, "min_len": min(lens) if lens else 0, "max_len": max(lens) if lens else 0, "rc": r.returncode,
                                "sha": hashlib.sha256(r.stdout).hexdigest() if r.returncode == 0 else None, "out_bytes": len(r.stdout)})
                blobs[name] = data
            subprocess.run(["rm", "-rf", work])
        print(suite, "families", len(fams), "regions so far", len(entries), flush=True)
    with tarfile.open(os.path.join(OUT, "inputs.tar.gz"), "w:gz") as tar:
        for name, data in sorted(blobs.items()):
            ti = tarfile.TarInfo(name); ti.size = len(data)
            tar.addfile(ti, io.BytesIO(data))
    json.dump({"note": "region files written by MLProbs' driver (utils/do_realign.py) and sha256 of the reference quickprobs' stdout for each", "regions": entries},
              open(os.path.join(OUT, "manifest.json"), "w"), indent=0)
    print("regions:", len(entries), "reference failures:", sum(1 for e in entries if e["rc"] != 0),
          "bytes:", os.path.getsize(os.path.join(OUT, "inputs.tar.gz")))
