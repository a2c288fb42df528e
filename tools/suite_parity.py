#!/usr/bin/env python3
"""Drop-in executables vs the reference programs on the bundled benchmark families (tests/golden/suites: inputs + SHA-256 of
the reference outputs, written by oracle/gen_suite_golden.py).  Runs quickprobs_b200 and c_p_np_aln_b200 -p 0 in directory
mode (one CUDA context per suite and tool) and compares every output byte for byte.  Prints one summary line per suite/tool
and writes the full report as JSON (default gpurun_out/suite_parity.json)."""
import os, sys, json, hashlib, subprocess, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = os.path.join(ROOT, "tests", "golden", "suites")
BIN = os.path.join(ROOT, "mlprobs_b200", "bin")


def run(report_path=None, suites=None):
    man = json.load(open(os.path.join(SUITES, "manifest.json")))["families"]
    tmp = tempfile.mkdtemp()
    with tarfile.open(os.path.join(SUITES, "inputs.tar.gz")) as tar:
        tar.extractall(tmp, filter="data")
    report = {"suites": {}, "mismatches": [], "failures": []}
    for suite in sorted({m["suite"] for m in man}):
        if suites and suite not in suites:
            continue
        fams = [m for m in man if m["suite"] == suite]
        for tool, exe, args, key in (("quickprobs", "quickprobs_b200", [], "qp_sha"), ("c_p_np_aln -p 0", "c_p_np_aln_b200", ["-p", "0"], "cpnp_sha")):
            outdir = os.path.join(tmp, "out_%s_%s" % (suite, key))
            os.makedirs(outdir)
            t0 = time.time()
            r = subprocess.run([os.path.join(BIN, exe)] + args + [os.path.join(tmp, suite), "-o", outdir], capture_output=True, text=True)
            dt = time.time() - t0
            same = diff = missing = skipped = 0
            for m in fams:
                if m[key] is None:
                    skipped += 1          # the reference itself failed on this family
                    continue
                p = os.path.join(outdir, m["name"])
                data = open(p, "rb").read() if os.path.exists(p) else b""
                if not data:
                    missing += 1
                    report["failures"].append({"suite": suite, "tool": tool, "name": m["name"]})
                elif hashlib.sha256(data).hexdigest() == m[key]:
                    same += 1
                else:
                    diff += 1
                    report["mismatches"].append({"suite": suite, "tool": tool, "name": m["name"]})
            report["suites"]["%s / %s" % (suite, tool)] = {"families": len(fams), "identical": same, "different": diff, "no_output": missing,
                                                           "reference_failed": skipped, "seconds": round(dt, 1), "rc": r.returncode,
                                                           "stderr_tail": r.stderr[-400:]}
            print("%-6s %-16s families %4d  identical %4d  different %3d  no output %3d  (reference failed %d)  %.1f s"
                  % (suite, tool, len(fams), same, diff, missing, skipped, dt), flush=True)
    if report_path:
        os.makedirs(os.path.dirname(report_path), exist_ok=True)
        json.dump(report, open(report_path, "w"), indent=1)
    return report


if __name__ == "__main__":
    run(os.path.join(ROOT, "gpurun_out", "suite_parity.json"), sys.argv[1:] or None)
