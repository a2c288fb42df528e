#!/usr/bin/env python3
"""`c_p_np_aln -p 1` on the CPU against the pinned reference outputs of the bundled benchmark families (no GPU, no reference run).

For every family of tests/golden/suites/manifest.json that carries `cpnp1_sha` (SHA-256 of what the unmodified reference program wrote
with its clock pinned, oracle/gen_suite_golden_p1.py / gen_suite_golden_rest.py): the ORACLE computes what the device stages compute
(Viterbi statistics -> model class, -p 1 posteriors, two relaxations; tests/common.py::cpnp_p1_sparse_set), the PRODUCT's host tail
(alignment graph + similar-set refinement, mlp_cpnp_np_finish_alignment_host, csrc/cpnp_graph.cpp / qp_tail.cpp) builds the alignment,
and the FASTA text -- input order, trimmed headers, 60 residues per line, as csrc/cpnp_main.cpp writes it -- must hash to the pinned value.
This is test infrastructure (it imports the oracle); the device flow itself is compared by tools/suite_parity.py cpnp1_sha on a GPU.

Usage: p1_host_sweep.py [--min-ref-s S] [--max-ref-s S] [--minutes M] [--procs P] [--log FILE] [--cpu-checked-only] [--manifest FILE] [suite ...]
--cpu-checked-only: just the families pinned after the last GPU run (manifest key `cpu_checked_only`).
Families run largest reference time first, one process per family; the sweep stops handing out work after M minutes."""
import hashlib, io, json, multiprocessing as mp, os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SUITES = os.path.join(ROOT, "tests", "golden", "suites")


def load_mfa(text):
    """csrc/cpnp_main.cpp::load_mfa (Sequence.h:96-112): trimmed headers, gaps stripped, upper case."""
    headers, seqs = [], []
    for rec in text.split(">")[1:]:
        head, _, body = rec.partition("\n")
        data = "".join(ch for ch in body if not ch.isspace() and ch not in ".-").upper()
        if not data:
            break
        headers.append(head.strip())
        seqs.append(data.encode())
    return headers, seqs


def fasta(headers, rows):
    out = io.StringIO()
    for h, r in zip(headers, rows):
        out.write(">" + h + "\n")
        for p in range(0, len(r), 60):
            out.write(r[p:p + 60] + "\n")
    return out.getvalue().encode()


def one(job):
    path, m, seed = job
    from common import cpnp_p1_sparse_set, cpnp_np_tail_from_csrset
    try:
        headers, seqs = load_mfa(open(path).read())
        t0 = time.time()
        dist, S, vm = cpnp_p1_sparse_set(seqs, threads=1)
        t1 = time.time()
        rows = cpnp_np_tail_from_csrset(S, seqs, dist, 100, seed)
        t2 = time.time()
        same = hashlib.sha256(fasta(headers, [r.decode() for r in rows])).hexdigest() == m["cpnp1_sha"]
        return m, len(seqs), "MATCH" if same else "DIFFERENT", t1 - t0, t2 - t1
    except Exception as e:                                    # a failure is a result, not the end of the sweep
        return m, -1, "ERROR %s" % (repr(e)[:200],), 0.0, 0.0


def main():
    args = sys.argv[1:]
    def opt(name, default):
        if name in args:
            i = args.index(name); v = args[i + 1]; del args[i:i + 2]; return v
        return default
    lo = float(opt("--min-ref-s", "0")); hi = float(opt("--max-ref-s", "1e30")); minutes = float(opt("--minutes", "60"))
    procs = int(opt("--procs", str(os.cpu_count()))); log = opt("--log", os.path.join(ROOT, "gpurun_out", "p1_host_sweep.txt"))
    only_new = "--cpu-checked-only" in args
    if only_new:
        args.remove("--cpu-checked-only")
    manifest = json.load(open(opt("--manifest", os.path.join(SUITES, "manifest.json"))))
    seed = int(manifest.get("p1_fixtime", 777))
    tmp = tempfile.mkdtemp()
    for arc in ("inputs.tar.gz", "inputs_rest.tar.gz"):
        if os.path.exists(os.path.join(SUITES, arc)):
            with tarfile.open(os.path.join(SUITES, arc)) as tar:
                tar.extractall(tmp, filter="data")
    fams = [m for m in manifest["families"] if m.get("cpnp1_sha") and lo <= float(m.get("cpnp1_s") or 0) <= hi and (not args or m["suite"] in args)
            and (not only_new or "cpnp1_sha" in m.get("cpu_checked_only", []))]
    fams.sort(key=lambda m: -float(m.get("cpnp1_s") or 0))
    jobs = [(os.path.join(tmp, m["suite"], m["name"]), m, seed) for m in fams]
    os.makedirs(os.path.dirname(log), exist_ok=True)
    t_start = time.time()
    done = match = 0
    bad = []
    with open(log, "a") as f, mp.Pool(procs) as pool:
        f.write("# %d families with a pinned -p 1 reference output, reference time in [%g, %g] s, %d processes\n" % (len(jobs), lo, hi, procs)); f.flush()
        it = pool.imap_unordered(one, jobs)
        while done < len(jobs):
            left = minutes * 60 - (time.time() - t_start)
            try:
                m, n, verdict, t_or, t_tail = it.next(timeout=max(1.0, left))
            except mp.TimeoutError:
                f.write("# stopped at the %g-minute limit\n" % minutes)
                break
            done += 1
            match += verdict == "MATCH"
            if verdict != "MATCH":
                bad.append((m["suite"], m["name"], verdict))
            f.write("%-6s %-12s n %4d %s  reference %.1f s, oracle stages %.1f s, host tail %.2f s\n" % (m["suite"], m["name"], n, verdict, float(m.get("cpnp1_s") or 0), t_or, t_tail))
            f.flush()
        f.write("# %d run, %d byte-identical, %d not: %s (%.1f minutes)\n" % (done, match, done - match, bad, (time.time() - t_start) / 60))
        pool.terminate()
    print("%d run, %d byte-identical, not identical: %s" % (done, match, bad))


if __name__ == "__main__":
    main()
