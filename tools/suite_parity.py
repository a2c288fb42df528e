#!/usr/bin/env python3
"""Drop-in executables vs the reference programs on the bundled benchmark families (tests/golden/suites: inputs + SHA-256 of
the reference outputs, written by oracle/gen_suite_golden.py and gen_suite_golden_p1.py).  Runs quickprobs_b200,
c_p_np_aln_b200 -p 0 and c_p_np_aln_b200 -p 1 --seed <the clock value the reference was pinned to> in directory mode (one CUDA
context per suite and tool) and compares every output byte for byte.  Prints one summary line per suite/tool
and writes the full report as JSON (default gpurun_out/suite_parity.json)."""
import os, sys, json, hashlib, subprocess, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = os.path.join(ROOT, "tests", "golden", "suites")
BIN = os.path.join(ROOT, "mlprobs_b200", "bin")


def run(report_path=None, suites=None, tools=None, gpu_verified_only=False):
    manifest = json.load(open(os.path.join(SUITES, "manifest.json")))
    man = manifest["families"]
    seed = str(manifest.get("p1_fixtime", 777))
    tmp = tempfile.mkdtemp()
    for arc in ("inputs.tar.gz", "inputs_rest.tar.gz"):          # the second archive: the large families pinned in round 2
        if os.path.exists(os.path.join(SUITES, arc)):
            with tarfile.open(os.path.join(SUITES, arc)) as tar:
                tar.extractall(tmp, filter="data")
    report = {"suites": {}, "mismatches": [], "failures": []}
    for suite in sorted({m["suite"] for m in man}):
        if suites and suite not in suites:
            continue
        fams = [m for m in man if m["suite"] == suite]
        for tool, exe, args, key in (("quickprobs", "quickprobs_b200", [], "qp_sha"), ("c_p_np_aln -p 0", "c_p_np_aln_b200", ["-p", "0"], "cpnp_sha"),
                                     ("c_p_np_aln -p 1", "c_p_np_aln_b200", ["-p", "1", "--seed", seed], "cpnp1_sha"),
                                     ("c_p_np_aln -G", "c_p_np_aln_b200", ["-G"], "cpnpG_sha")):
            if tools and key not in tools:
                continue
            if key == "cpnpG_sha" and not (tools and key in tools):
                continue                  # the feature-line run is opt-in (added at the end of round 1, not yet run on a GPU)
            fams = [m for m in man if m["suite"] == suite and key in m and (key != "cpnpG_sha" or m.get("cpnpG_exact"))]
            max_s = float(os.environ.get("MLP_SUITE_MAX_REF_S", "0"))      # GPU-time budget: skip families whose REFERENCE run of this tool took longer
            if max_s > 0:
                fams = [m for m in fams if float(m.get(key[:-4] + "_s", 0) or 0) <= max_s]
            if gpu_verified_only:         # the GPU test-suite: families pinned after the last GPU run of a round (checked on the CPU only, tools/p1_host_sweep.py) stay out
                fams = [m for m in fams if key not in m.get("cpu_checked_only", [])]
            if not fams:
                continue
            outdir = os.path.join(tmp, "out_%s_%s" % (suite, key))
            os.makedirs(outdir)
            indir = os.path.join(tmp, suite)
            if len(fams) != sum(1 for m in man if m["suite"] == suite):      # this tool has reference outputs for a subset only
                indir = os.path.join(tmp, "in_%s_%s" % (suite, key))
                os.makedirs(indir)
                for m in fams:
                    os.symlink(os.path.join(tmp, suite, m["name"]), os.path.join(indir, m["name"]))
            t0 = time.time()
            r = subprocess.run([os.path.join(BIN, exe)] + args + [indir, "-o", outdir], capture_output=True, text=True)
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
    # arguments: suite names and/or tool keys (qp_sha, cpnp_sha, cpnp1_sha)
    keys = [x for x in sys.argv[1:] if x.endswith("_sha")]
    names = [x for x in sys.argv[1:] if not x.endswith("_sha")]
    run(os.path.join(ROOT, "gpurun_out", "suite_parity%s.json" % ("_" + "_".join(keys) if keys else "")), names or None, keys or None)
