#!/usr/bin/env python3
"""quickprobs_b200 on the short region files MLProbs' driver realigns (tests/golden/regions, written by
oracle/gen_region_golden.py): directory mode (one CUDA context for all of them), every output compared byte for byte with the
reference quickprobs' (SHA-256 in the manifest).  Prints a summary and writes gpurun_out/region_parity.json.
Not part of the pytest run yet: the fixtures were generated at the end of round 1, after the GPU budget was spent."""
import hashlib, json, os, subprocess, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REG = os.path.join(ROOT, "tests", "golden", "regions")
EXE = os.path.join(ROOT, "mlprobs_b200", "bin", "quickprobs_b200")


def run(report_path=None):
    man = json.load(open(os.path.join(REG, "manifest.json")))["regions"]
    tmp = tempfile.mkdtemp()
    indir, outdir = os.path.join(tmp, "in"), os.path.join(tmp, "out")
    os.makedirs(indir); os.makedirs(outdir)
    with tarfile.open(os.path.join(REG, "inputs.tar.gz")) as tar:
        tar.extractall(indir, filter="data")
    t0 = time.time()
    r = subprocess.run([EXE, indir, "-o", outdir], capture_output=True, text=True)
    dt = time.time() - t0
    same, diff, missing, skipped = 0, [], [], 0
    for m in man:
        if m["sha"] is None:
            skipped += 1
            continue
        p = os.path.join(outdir, m["name"])
        data = open(p, "rb").read() if os.path.exists(p) else b""
        if not data and m["out_bytes"]:
            missing.append(m["name"])
        elif hashlib.sha256(data).hexdigest() == m["sha"]:
            same += 1
        else:
            diff.append(m["name"])
    rep = {"regions": len(man), "identical": same, "different": diff, "no_output": missing, "reference_failed": skipped,
           "seconds": round(dt, 2), "rc": r.returncode, "stderr_tail": r.stderr[-400:]}
    print("regions %d  identical %d  different %d  no output %d  (reference failed %d)  %.2f s = %.1f ms per region"
          % (len(man), same, len(diff), len(missing), skipped, dt, 1e3 * dt / max(len(man), 1)))
    if report_path:
        os.makedirs(os.path.dirname(report_path), exist_ok=True)
        json.dump(rep, open(report_path, "w"), indent=1)
    return rep


if __name__ == "__main__":
    rep = run(os.path.join(ROOT, "gpurun_out", "region_parity.json"))
    sys.exit(0 if not rep["different"] and not rep["no_output"] else 1)
