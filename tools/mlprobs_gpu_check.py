#!/usr/bin/env python3
"""MLProbs' UNMODIFIED Python driver on top of this repository's executables, on a GPU (BASELINE config #1).
The MLProbs checkout is not part of this repository: a scratch copy of MLProbs.py, utils/, classifier/, the two reference
programs and a few TEST families travels to the GPU box as untracked files (gpurun_scratch/mlprobs, git-ignored, deleted after
the run).  For every family the driver runs twice through tools/mlprobs_overlay: over quickprobs_b200 + c_p_np_aln_b200, and over
the checkout's own CPU programs pinned to one core (their only repeatable setting), and once more over the drop-ins in
persistent-process mode (MLP_B200_SERVER=1: one CUDA context for all the driver's calls); the results are compared, by header, with
each other and with the checkout's published alignment (output4evaluation/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "mlprobs_overlay"))
import run as overlay_run

MLPROBS = os.path.join(ROOT, "gpurun_scratch", "mlprobs")
FAMILIES = [("bali3", "BB11001"), ("sabre", "sup_139"), ("ox", "104s10"), ("sabre", "sup_200"), ("bali3", "BB12003"), ("oxx", "_676s4")]


def fasta(path):
    out, name = {}, None
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            name = line; out[name] = ""
        elif line:
            out[name] += line
    return out


rep = []
for suite, name in FAMILIES:
    src = os.path.join(MLPROBS, "TEST", suite, "in", name)
    if not os.path.exists(src):
        continue
    row = {"suite": suite, "name": name}
    for arm in ("b200", "b200_server", "reference"):
        out = "/tmp/mlprobs_%s_%s.msa" % (arm, name)
        if os.path.exists(out): os.remove(out)
        t0 = time.time()
        rc, log = overlay_run.run(MLPROBS, "reference" if arm == "reference" else "b200", src, out, seed=777, quiet=True,
                                  one_core=(arm == "reference"), server=(arm == "b200_server"))
        row[arm] = {"rc": rc, "seconds": round(time.time() - t0, 2), "final": "Got the final MSA" in (log or ""),
                    "branch": [l for l in (log or "").splitlines() if "egion" in l or "rogressive" in l][:3]}
        row[arm + "_rows"] = fasta(out) if os.path.exists(out) else None
    pub = os.path.join(MLPROBS, "output4evaluation", suite, name)
    published = fasta(pub) if os.path.exists(pub) else None
    row["b200_equals_reference_driver"] = row["b200_rows"] is not None and row["b200_rows"] == row["reference_rows"] and row["b200_server_rows"] == row["reference_rows"]
    row["b200_equals_published"] = (row["b200_rows"] == published) if published else None
    row["reference_equals_published"] = (row["reference_rows"] == published) if published else None
    del row["b200_rows"], row["reference_rows"], row["b200_server_rows"]
    rep.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "mlprobs_driver_gpu.json"), "w"), indent=1)
ok = all(r["b200_equals_reference_driver"] for r in rep)
print("MLProbs.py over the b200 executables: %d families, identical to the driver over the reference programs: %s" % (len(rep), ok))
sys.exit(0 if ok else 1)
