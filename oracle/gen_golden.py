#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/ref_cpnp, oracle/_ref/ref_qp).

Runs only in the build container (needs /root/reference for the input families and oracle/_ref built by
`make -C oracle ref`).  Two fixture kinds:
  full   : sequences, tables, distances, every CSR matrix (and dense posteriors when small) -- tiny families
  digest : sequences, distances, tree weights, and per-pair nnz + CRC32 of the column / value arrays -- mid-size
           families (a checksum of checksums keeps N>200 cases a few hundred KB)
"""
import os, subprocess, sys, zlib, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _dumpfmt import read_dump

REF = "/root/reference/TEST"
OUT = os.path.join(ROOT, "tests", "golden")
CPNP = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
QP = os.path.join(ROOT, "oracle", "_ref", "ref_qp")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def pairs(n):
    return [(a, b) for a in range(n) for b in range(a + 1, n)]


def pack(d, tags, full, dense):
    n = int(d["n"][0])
    out = {"n": d["n"], "lens": d["lens"], "residues": d["residues"], "distances": d["distances"]}
    for k in ("pid", "pid_ref", "initDistrib2", "weights", "seldist", "distances_after_tree", "cons.iterations",
              "cons.selfweight", "reps", "p1", "vit.ident", "vit.len", "variance_mean", "gline"):
        if k in d:
            out[k] = d[k]
    for k in d:
        if k.startswith("hmm.") or k.startswith("part."):
            out[k] = d[k]
    for tag in tags:
        nnz = np.zeros(len(pairs(n)), np.int32); ccrc = np.zeros_like(nnz, dtype=np.uint32); vcrc = np.zeros_like(ccrc)
        rcrc = np.zeros_like(ccrc)
        for p, (a, b) in enumerate(pairs(n)):
            t = "pair.%d.%d.%s" % (a, b, tag)
            nnz[p] = len(d[t + ".col"]); ccrc[p] = crc(d[t + ".col"].astype(np.int32)); vcrc[p] = crc(d[t + ".val"])
            rcrc[p] = crc(d[t + ".rowptr"])
            if full:
                out[t + ".rowptr"] = d[t + ".rowptr"]; out[t + ".col"] = d[t + ".col"].astype(np.int32); out[t + ".val"] = d[t + ".val"]
        out["digest.%s.nnz" % tag] = nnz; out["digest.%s.col_crc" % tag] = ccrc; out["digest.%s.val_crc" % tag] = vcrc
        out["digest.%s.rowptr_crc" % tag] = rcrc
    if dense:
        for k in d:
            if k.endswith((".post", ".post5", ".postP", ".postL")):
                out[k] = d[k]
    return out


def read_rows(path):
    """FASTA alignment -> (n, columns) uint8 matrix"""
    rows = []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"): rows.append("")
        elif line: rows[-1] += line
    return np.frombuffer("".join(rows).encode(), np.uint8).reshape(len(rows), -1).copy()


def run_cpnp(name, fasta, full, dense, pid=None, p1=False, reps=2):
    with tempfile.TemporaryDirectory() as td:
        dump = os.path.join(td, "d.bin")
        cmd = ["taskset", "-c", "0", CPNP, "dump", fasta, dump, "--reps", str(reps)]
        if pid is not None: cmd += ["--pid", str(pid)]
        if p1: cmd += ["--p1"]
        if not dense: cmd += ["--nodense"]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        d = read_dump(dump)
        extra = {}
        if pid is None and not p1:
            # the reference's whole `-p 0` program (tree, progressive alignment, iterative refinement) on one OpenMP thread:
            # with more threads its refinement races on the shared posterior and the output changes from run to run
            tmpfa = os.path.join(td, "in.fa")
            res = d["residues"].tobytes().decode(); at = 0
            with open(tmpfa, "w") as f:
                for i, L in enumerate(d["lens"]):
                    f.write(">s%d\n%s\n" % (i, res[at:at + int(L)])); at += int(L)
            for key, ir in (("msa", None), ("msa_ir0", 0)):
                out = os.path.join(td, key + ".fa")
                subprocess.check_call([CPNP, "msa", tmpfa, out, "--threads", "1"] + ([] if ir is None else ["--ir", str(ir)]), stdout=subprocess.DEVNULL)
                extra[key] = read_rows(out)
                extra[key + "_order"] = np.array([int(l[2:]) for l in open(out) if l.startswith(">")], np.int32)
    tags = ["s%d" % r for r in range(reps + 1)]
    out = pack(d, tags, full, dense)
    out.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".npz")))


def run_cpnp_p1(name, fasta, fixtime):
    """The reference's whole `c_p_np_aln -p 1` program (non-progressive: alignment graph + similar-set refinement) on one
    OpenMP thread.  It reseeds rand() from time(0) before every refinement sweep; `--fixtime` makes the harness' own time()
    return a constant, so the output is reproducible without touching the reference sources."""
    def read_fasta(path):
        rows = []
        for line in open(path):
            line = line.strip()
            if line.startswith(">"): rows.append("")
            elif line: rows[-1] += line
        return rows
    seqs = [r.replace("-", "").replace(".", "").upper() for r in read_fasta(fasta)]
    out = {"n": np.array([len(seqs)], np.int32), "lens": np.array([len(x) for x in seqs], np.int32),
           "residues": np.frombuffer("".join(seqs).encode(), np.uint8).copy(), "fixtime": np.array([fixtime], np.int64)}
    with tempfile.TemporaryDirectory() as td:
        tmpfa = os.path.join(td, "in.fa")
        with open(tmpfa, "w") as f:
            for i, x in enumerate(seqs): f.write(">s%d\n%s\n" % (i, x))
        for key, ir in (("msa", None), ("msa_ir0", 0)):
            o = os.path.join(td, key + ".fa")
            subprocess.check_call([CPNP, "msa", tmpfa, o, "--p1", "--threads", "1", "--fixtime", str(fixtime)] + ([] if ir is None else ["--ir", str(ir)]),
                                  stdout=subprocess.DEVNULL)
            out[key] = read_rows(o)
            assert [l.strip() for l in open(o) if l.startswith(">")] == [">s%d" % i for i in range(len(seqs))]     # input order
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".npz")))


def run_qp(name, fasta, full, dense):
    with tempfile.TemporaryDirectory() as td:
        dump = os.path.join(td, "d.bin")
        cmd = [QP, "dump", fasta, dump, "--threads", "8"]   # quickprobs is thread-count independent
        if not dense: cmd += ["--nodense"]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        d = read_dump(dump)
        # the reference's final alignment (ConstructionStage + ColumnRefinement) and the one before refinement
        msa = os.path.join(td, "msa.fa")
        subprocess.check_call([QP, "msa", fasta, msa, "--threads", "8"], stdout=subprocess.DEVNULL)
        extra = {"msa": read_rows(msa), "msa_construct": read_rows(msa + ".construct")}
    out = pack(d, ["s0", "t0", "sF", "tF"], full, dense)
    out.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".npz")))


def run_qp_compact(name, seqs):
    """Large synthetic family (mlprobs_b200.synth.family_clustered) where QuickProbs' selectivity really filters third
    sequences: the dump (gigabytes) is streamed once and only per-pair nnz + ONE combined CRC32 (row pointers | columns |
    values, in that order) per tag is kept, for the tags s0 (after the posterior stage), sF and tF (after consistency,
    both orientations), plus distances, weights and subtree distances."""
    import struct
    from _dumpfmt import _DT
    n = len(seqs)
    P = {ab: p for p, ab in enumerate(pairs(n))}
    out = {"n": np.array([n], np.int32), "lens": np.array([len(x) for x in seqs], np.int32),
           "residues": np.frombuffer(b"".join(seqs), np.uint8).copy()}
    tags = ("s0", "sF", "tF")
    nnz = {t: np.zeros(len(P), np.int32) for t in tags}
    part = {t: {} for t in tags}
    comb = {t: np.zeros(len(P), np.uint32) for t in tags}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "in.fa"); dump = os.path.join(td, "d.bin")
        with open(fa, "w") as f:
            for i, x in enumerate(seqs): f.write(">s%05d\n%s\n" % (i, x.decode()))
        subprocess.check_call([QP, "dump", fa, dump, "--threads", "8", "--nodense"], stdout=subprocess.DEVNULL)
        with open(dump, "rb") as f:
            while True:
                h = f.read(4)
                if len(h) < 4: break
                (nl,) = struct.unpack("<I", h)
                key = f.read(nl).decode(); dt = _DT[f.read(1)]
                (nd,) = struct.unpack("<I", f.read(4))
                dims = struct.unpack("<%dQ" % nd, f.read(8 * nd))
                cnt = int(np.prod(dims)) if nd else 1
                arr = np.frombuffer(f.read(cnt * np.dtype(dt).itemsize), dtype=dt).reshape(dims)
                if key in ("distances", "weights", "seldist", "cons.iterations", "cons.selfweight"):
                    out[key] = arr.copy()
                elif key.startswith("pair."):
                    parts = key.split(".")
                    if len(parts) != 5: continue                       # pair.a.b.dist: already in `distances`
                    _, a, b, tag, what = parts
                    if tag not in tags or what == "code": continue     # the value already is code / 65535
                    p = P[(int(a), int(b))]
                    part[tag].setdefault(p, {})[what] = arr.astype(np.int32) if what == "col" else arr
                    got = part[tag][p]
                    if len(got) == 3:
                        nnz[tag][p] = len(got["col"])
                        c = zlib.crc32(np.ascontiguousarray(got["rowptr"]).tobytes())
                        c = zlib.crc32(np.ascontiguousarray(got["col"]).tobytes(), c)
                        comb[tag][p] = zlib.crc32(np.ascontiguousarray(got["val"]).tobytes(), c) & 0xffffffff
                        del part[tag][p]
    for t in tags:
        out["digest.%s.nnz" % t] = nnz[t]; out["digest.%s.crc" % t] = comb[t]
    sd = out["seldist"].reshape(n, n)
    acc = [(np.maximum(sd[a], sd[b]) <= 200).sum() - 2 for a, b in list(P)[::97]]
    print("accepted third sequences per pair (sample): min %d max %d mean %.1f of %d" % (min(acc), max(acc), float(np.mean(acc)), n - 2))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["qp_syn400"]:
        # N = 400 x L ~ 120, eight sub-families of 50: most third sequences are rejected by the selectivity filter
        sys.path.insert(0, ROOT)
        from mlprobs_b200 import synth
        run_qp_compact("qp_syn400", synth.family_clustered(8, 50, 120, seed=20220148 + 7))
        sys.exit(0)
    f = lambda b, x: os.path.join(REF, b, "in", x)
    # tiny, everything stored (dense posteriors of all three models)
    run_cpnp("cpnp_sup139_mix", f("sabre", "sup_139"), True, True, pid=0)
    run_cpnp("cpnp_sup139_local", f("sabre", "sup_139"), True, True, pid=2)
    run_cpnp("cpnp_sup139_part", f("sabre", "sup_139"), True, True, pid=3)
    run_cpnp("cpnp_sup139_p1mix", f("sabre", "sup_139"), True, False, pid=0, p1=True)
    run_cpnp("cpnp_sup002_ref", f("sabre", "sup_002"), True, False)           # model chosen by the reference itself
    run_qp("qp_sup139", f("sabre", "sup_139"), True, True)
    run_qp("qp_sup002", f("sabre", "sup_002"), True, False)
    # mid-size, digests only
    run_cpnp("cpnp_676s4_ref", f("oxx", "_676s4"), False, False)
    run_qp("qp_676s4", f("oxx", "_676s4"), False, False)                       # N=51 -> 1 consistency iteration, cutoff 1e-5
    run_qp("qp_75t2", f("oxx", "__75t2"), False, False)                        # N=204 -> selectivity excludes some z
    # whole-program `-p 1` outputs (inputs + final alignments only)
    run_cpnp_p1("cpnp_p1_sup139", f("sabre", "sup_139"), 777)
    run_cpnp_p1("cpnp_p1_BB12003", f("bali3", "BB12003"), 777)               # the refinement order changes the result here
    run_cpnp_p1("cpnp_p1_676s4", f("oxx", "_676s4"), 1792000000)             # N=51, non-standard letters
