#!/usr/bin/env python3
"""bench.py -- all-pairs posterior + consistency throughput (BASELINE.json metric) on N B200s of one node.

A step = one pass of the hot path over one synthetic protein family: posterior stage (QuickProbs flavour:
5-state pair-HMM + FP64 partition function per cell -> merged posterior -> MEA distance -> CSR, both
orientations), UPGMA guide tree (built on the device from the resident distances), one consistency repetition (N > 50, as the reference), run through the C ABI.
`value` = pair-HMM cell updates per second (1 cell update = one (i,j) cell through forward+backward+posterior of ONE
model; the QuickProbs flavour runs 2 models per cell) over the whole step, inputs resident; `e2e` = same with the
host->device copy of the family and the device->host read-back of distances and CSR inside the timed region.
At N > 1 the pairs are sharded; after the tree every rank imports only the matrices QuickProbs' selectivity lets a
relaxation read (mlp_exchange_needed), the relaxed set stays sharded and every rank reads its own shard back.
`parity_digest` pins the result: CRC32 of the distance matrix and of the per-matrix device digests (summed over ranks)
against the committed single-GPU value (tests/golden/bench_digest.json).
`--impl reference` times the compiled reference (oracle/_ref/ref_qp) on the host cores on a bounded sample.
The line also carries `cpnp` (the c_p_np_aln flavour on the same family, measured at N = 1) and `families_per_sec`
(FASTA -> FASTA through the drop-in executables).
"""
import argparse, json, os, subprocess, sys, tarfile, tempfile, threading, time, zlib
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the config the metric is quoted on
    "A": dict(n=1000, length=300, name="synthetic 1,000 protein seqs x len 300: all-pairs posterior + consistency"),
    "small": dict(n=200, length=300, name="synthetic 200 protein seqs x len 300 (development size)"),
    # BASELINE.json configs[4]: the sparse set (16 M matrices) does not fit HBM -> streamed two-pass flow (mlp_stream_begin / mlp_restrict_pairs)
    "B": dict(n=4000, length=500, stream=True, scratch=24 << 30, cells=1 << 31,
              name="synthetic 4,000 protein seqs x len 500 stress: 8.0 M pairs, all-pairs posterior + consistency, streamed"),
    # the same streamed flow at a size where the ordinary flow also fits: the two digests are compared in the run
    "Bs": dict(n=400, length=120, stream=True, scratch=64 << 20, cells=0, clustered=(8, 50), compare=True,
               name="synthetic 400 protein seqs (8 sub-families) x len 120: streamed flow checked against the ordinary flow"),
}
FP32_ISSUE_PEAK = 148 * 128 * 1.965e9          # lane-instructions / s
# algorithmic FP32-slot equivalents per grid cell: 5-state forward+backward+posterior, partition function
# (FP64 ops counted as 2 slots), merge + MEA + threshold; the whole QuickProbs stage = 430, cpnp's local model 270
SLOTS_PER_CELL = {"hmm5": 358.0, "part": 62.0, "merge": 10.0, "stage_qp": 430.0, "local": 270.0, "stage_cpnp3": 700.0}
DIGEST_FILE = os.path.join(ROOT, "tests", "golden", "bench_digest.json")
# ncu --set full of tools/prof_run.py 192 300, main batch (17,254 pairs, 1.6237e9 cells, all C = 5), round-2 kernels
# (profiles/r2_posterior_kernels_ncu_full.txt): k_hmm_fwd_c 1.136 + 6.994 GB, k_hmm_bwd_c 7.557 + 6.985 GB of DRAM traffic
# -> 13.96 B per DP cell against 12 algorithmic (the slot layout pads 301 x 301 cells to 332 x 320 elements)
NCU_DRAM_BYTES_PER_CELL_HMM5 = (1.13579 + 6.99383 + 7.55665 + 6.98525) * 1e9 / 1.6237e9
NCU_TRAFFIC_SOURCE = "profiles/r2_posterior_kernels_ncu_full.txt"


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.stop_flag = gpu, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_family(wl, rank=0):
    from mlprobs_b200 import synth
    if wl.get("clustered"):
        return synth.family_clustered(wl["clustered"][0], wl["clustered"][1], wl["length"], seed=20220148 + 4)
    return synth.family_fast(wl["n"], wl["length"], seed=20220148 + 2)


def write_fasta(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">s%05d\n%s\n" % (i, s.decode()))


def ref_qp_sample(seqs, n_s, cores):
    """The UNMODIFIED QuickProbs CPU stage objects (oracle/_ref/ref_qp) on the first n_s sequences, all host cores."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_qp")
    tmp = os.path.join("/tmp", "mlp_ref_sample_%d.fa" % os.getpid())
    write_fasta(tmp, seqs[:n_s])
    out = subprocess.run([exe, "bench", tmp, "--threads", str(cores)], capture_output=True, text=True).stdout
    os.unlink(tmp)
    j = json.loads(out.strip().splitlines()[-1])
    return j["t_posterior_s"] + j["t_tree_s"] + j["t_relax_s"], j["cells"] * j["models"], j


def run_reference(args, wl):
    """Reference arm: the UNMODIFIED QuickProbs CPU code (oracle/_ref/ref_qp) on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_qp")
    cores = os.cpu_count() or 1
    seqs = make_family(wl)
    n_s = int(args.ref_sample)      # bounded sample of the same workload: the first n_s sequences (~10-20 s of CPU work per step)
    kind = "reference" if os.path.exists(exe) else "port"
    times, cells = [], 0
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            t, cells, _ = ref_qp_sample(seqs, n_s, cores)
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            ht, pt = O.hmm_tables(), O.part_tables(O.QP)
            sample = seqs[:n_s]
            t0 = time.time()
            dist, S, _ = O.posterior_stage(O.QP, 3, ht, pt, sample, threads=cores)
            t = time.time() - t0
            cells = 2 * sum((len(a) + 1) * (len(b) + 1) for i, a in enumerate(sample) for b in sample[i + 1:])
        if it >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    gcups = cells / (ms * 1e-3) / 1e9
    line = {"impl": "reference", "metric": "pair_hmm_cell_updates_per_second", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": wl["name"], "flavour": "quickprobs", "models_per_cell": 2},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": "first %d of %d sequences (%d pairs) of the same family, posterior stage + tree + consistency" % (n_s, wl["n"], n_s * (n_s - 1) // 2)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "alignments_per_sec": (n_s * (n_s - 1) // 2) / (ms * 1e-3)}
    print(json.dumps(line))


TRACE = {} if os.environ.get("MLP_BENCH_TRACE") else None    # developer: wall time of every phase of a step, printed to stderr


def _tr(name, t0):
    if TRACE is not None:
        TRACE[name] = TRACE.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


def one_step(eng, M, n, e2e, seqs=None, world=1, host_out=None, want_dist=False):
    """posterior stage -> [distances all-reduce] -> guide tree on the device -> [selective import] -> consistency [-> read-back of the own shard]."""
    t0 = time.perf_counter()
    if e2e:
        eng.set_sequences(seqs)                       # host -> device copy of the family inside the timed region (keeps the shard)
        t0 = _tr("e2e.set_sequences", t0)
    stats = []
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    stats.append(("posterior", eng.stats()))
    t0 = _tr("posterior", t0)
    if world > 1:
        eng.exchange_distances()
    d = eng.distances() if (e2e or want_dist) else None   # part of the result (the e2e arm reads it back every step); the tree no longer needs it on the host
    t0 = _tr("distances", t0)
    # guide tree on the device (tree_dev.cu): weights (saturated at 1e-6, ExtendedMSA.cpp:237-238) and selectivity distances stay resident
    # for the consistency stage; only a sharded run needs the selectivity distances on the host (for its import list)
    tree = eng.qp_guide_tree_device(1e-6, want_seldist=(world > 1))
    t0 = _tr("tree", t0)
    if world > 1:
        eng.exchange_needed(tree["seldist"], 200.0); stats.append(("exchange", eng.stats()))
        t0 = _tr("exchange_needed", t0)
    iters = 1 if n > 50 else 2
    for it in range(iters):
        cutoff = float(np.float32(0.01)) if it < iters - 1 else float(np.float32(1e-5))
        eng.relax(M.QP, None, None, 200.0, 3.0, cutoff)
        stats.append(("relax", eng.stats()))
        if world > 1 and it < iters - 1:
            eng.exchange(); stats.append(("exchange", eng.stats()))
    t0 = _tr("relax", t0)
    out = None
    if e2e:
        # device -> host read of this rank's part of the result (QuickProbs' own packed cell format, PackedSparseMatrix) into
        # caller-owned page-locked buffers.  Split call: the copy runs on its own stream and is waited for by the next step's
        # consistency stage (or by csr_packed_end after the last step), so it overlaps the next step's posterior stage.
        out = eng.csr_packed_begin(host_out)
        t0 = _tr("e2e.read_back_begin", t0)
    return stats, out, d


def cpnp_step(eng, M, mask):
    """c_p_np_aln flavour: posterior stage (models by `mask`), two unselective consistency repetitions (MSA.cpp:1172-1281)."""
    stats = []
    eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)
    stats.append(("posterior", eng.stats()))
    for _ in range(2):
        eng.relax(M.CPNP_P0, cutoff=0.01)
        stats.append(("relax", eng.stats()))
    return stats


def cpnp_object(M, dev, seqs, total_cells, n_sub):
    """Second object of the line: the c_p_np_aln flavour.  Posterior stage on the whole family for model class 0 (three models)
    and class 2 (local model only); the unselective consistency (every third sequence, two repetitions) is cubic in N, so it
    is timed on the first n_sub sequences and reported as MAC/s and per (pair, z) -- the reference does the same work."""
    out = {}
    npairs = len(seqs) * (len(seqs) - 1) // 2
    for label, mask, models, slots in (("class0_three_models", 7, 3, SLOTS_PER_CELL["stage_cpnp3"]), ("class2_local_only", 4, 1, SLOTS_PER_CELL["local"])):
        eng = M.Engine(dev)
        h, p = M.default_tables(M.CPNP_P0, 0.100675); eng.set_tables(h, p); eng.set_sequences(seqs)
        eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)            # warm-up (pool sizing)
        eng.posterior_all_pairs(M.CPNP_P0, mask, 0.01)
        st = eng.stats()
        ms = st["ms_total"]
        out[label] = {"posterior_ms": ms, "gcups": total_cells * models / (ms * 1e-3) / 1e9, "models_per_cell": models,
                      "kernel_ms": {k: v for k, v in st["ms_kernel"].items() if v},
                      "stage_frac_fp32_issue": total_cells * slots / (ms * 1e-3) / FP32_ISSUE_PEAK}
        eng.close()
    sub = seqs[:n_sub]
    eng = M.Engine(dev)
    h, p = M.default_tables(M.CPNP_P0, 0.100675); eng.set_tables(h, p); eng.set_sequences(sub)
    for rep in range(2):                                         # first pass warms the pools
        eng.posterior_all_pairs(M.CPNP_P0, 7, 0.01)
        cells_in = eng.total_cells()
        eng.relax(M.CPNP_P0, cutoff=0.01); r1 = eng.stats()
        eng.relax(M.CPNP_P0, cutoff=0.01); r2 = eng.stats()
    ns = len(sub); pz = ns * (ns - 1) // 2 * (ns - 2)
    out["relax_unselective"] = {"n_sub": ns, "ms_rep1": r1["ms_total"], "ms_rep2": r2["ms_total"], "pair_z_per_sec": pz / (r1["ms_total"] * 1e-3),
                                "cells_in": int(cells_in), "extrapolated_ms_per_rep_at_n": r1["ms_total"] * (npairs * (len(seqs) - 2)) / pz,
                                "note": "first %d sequences; cost is proportional to pairs x (N-2), the extrapolation to N = %d is labelled as such" % (ns, len(seqs))}
    eng.close()
    # the reference's own code for this flavour on the host cores: unmodified c_p_np_aln sources behind oracle/_ref/ref_cpnp, all models
    # (class 0), two consistency repetitions, on the first 64 sequences; the consistency part is cubic in N and is compared per (pair, z)
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cpnp")
    if os.path.exists(exe):
        n_ref = 64
        cores = os.cpu_count() or 1
        tmp = os.path.join("/tmp", "mlp_cpnp_ref_%d.fa" % os.getpid())
        write_fasta(tmp, seqs[:n_ref])
        r = subprocess.run([exe, "bench", tmp, "--pid", "0", "--reps", "2", "--threads", str(cores)], capture_output=True, text=True).stdout
        os.unlink(tmp)
        try:
            j = json.loads(r.strip().splitlines()[-1])
            pz_ref = 2 * j["pairs"] * (j["n"] - 2) / j["t_relax_s"]
            out["cpu_reference"] = {"kind": "reference", "cores": cores, "sample": "first %d sequences of the same family (%d pairs), model class 0, 2 consistency repetitions" % (n_ref, j["pairs"]),
                                    "posterior_gcups": j["cells"] * j["models"] / j["t_posterior_s"] / 1e9, "posterior_s": j["t_posterior_s"],
                                    "relax_pair_z_per_sec": pz_ref, "relax_s": j["t_relax_s"],
                                    "extrapolated_relax_s_per_rep_at_n": (npairs * (len(seqs) - 2)) / pz_ref,
                                    "note": "the extrapolation to N = %d assumes the cost stays proportional to pairs x (N-2); it is an extrapolation, not a measurement" % len(seqs)}
        except Exception as e:
            out["cpu_reference"] = {"error": str(e)[:200]}
    return out


def stream_step(eng, M, n, rank, world, want_digest=False):
    """BASELINE config #5's flow: streamed stage over all owned pairs (every batch finished, digested and dropped) -> distances
    all-reduce -> guide tree on the device -> ordinary stage over the pairs inside a <= 200-leaf subtree only -> selective import ->
    consistency on those pairs.  Returns (per-stage stats, combined per-matrix digest of this rank or None)."""
    st = {}
    t0 = time.perf_counter()
    eng.stream_begin(1)
    eng.posterior_all_pairs(M.QP, 3, 0.01); st["streamed_stage"] = eng.stats()
    streamed = eng.stream_end(want_digest)
    st["t_streamed_ms"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    if world > 1:
        eng.exchange_distances()
    tree = eng.qp_guide_tree_device(1e-6, want_seldist=True)
    sd = tree["seldist"].reshape(n, n)
    st["t_tree_ms"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    eng.restrict_pairs(sd, 200.0)
    eng.posterior_all_pairs(M.QP, 3, 0.01); st["ingroup_stage"] = eng.stats()
    st["t_ingroup_ms"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    if world > 1:
        eng.exchange_needed(sd, 200.0); st["exchange"] = eng.stats()
    st["t_exchange_ms"] = (time.perf_counter() - t0) * 1e3; t0 = time.perf_counter()
    eng.relax(M.QP, None, None, 200.0, 3.0, float(np.float32(1e-5))); st["relax"] = eng.stats()
    st["t_relax_ms"] = (time.perf_counter() - t0) * 1e3
    digest = None
    if want_digest:
        ingroup = sd <= 200.0
        np.fill_diagonal(ingroup, False)
        digest = np.where(ingroup, eng.set_digest().reshape(n, n), streamed.reshape(n, n))
        st["ingroup_pairs_all_ranks"] = int(ingroup.sum() // 2)
    eng.set_shard(rank, world)
    return st, digest


def run_streamed(args, wl, M, torch, dist, eng, seqs, rank, world, dev, sampler):
    """bench line of a streamed workload (B: 4,000 x 500; Bs: the same flow next to the ordinary one)."""
    n = len(seqs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_digest(dg):
        t = torch.from_numpy(np.ascontiguousarray(dg).view(np.int64).reshape(-1)).cuda()
        if dist is not None:
            dist.all_reduce(t)
        return zlib.crc32(t.cpu().numpy().tobytes()) & 0xffffffff

    lens = np.array([len(s) for s in seqs], np.int64) + 1
    total_cells = int((lens.sum() ** 2 - (lens ** 2).sum()) // 2)
    npairs = n * (n - 1) // 2
    eng.configure(wl["scratch"], wl["cells"])
    for _ in range(args.warmup):
        stream_step(eng, M, n, rank, world)
    barrier()
    t0 = time.perf_counter()
    last = None
    for k in range(args.steps):
        last, dg = stream_step(eng, M, n, rank, world, want_digest=(k == args.steps - 1))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps      # the last step includes the digest read-back (n*n*8 bytes per rank)
    crc_streamed = reduce_digest(dg)
    free_b, total_b = torch.cuda.mem_get_info(dev)
    compare = None
    if wl.get("compare"):
        eng.configure(0, 0)
        one_step(eng, M, n, False, world=world)
        compare = reduce_digest(eng.set_digest().reshape(n, n))
    t = torch.tensor([wall_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_ms = float(t[0])
    sampler.stop_flag = True; sampler.join(timeout=2)
    eng.close()
    if rank != 0:
        return
    models = 2
    km = {k: v for k, v in last["streamed_stage"]["ms_kernel"].items() if v}
    line = {"metric": "pair_hmm_cell_updates_per_second", "value": total_cells * models / (wall_ms * 1e-3) / 1e9, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": wl["name"], "flavour": "quickprobs (5-state pair-HMM f32 + partition function f64, 1 consistency rep)", "models_per_cell": models,
                       "n": n, "pairs": npairs, "cells": total_cells,
                       "flow": "streamed two-pass: every rank's pairs go through the posterior stage in batches whose matrices are finished (re-quantisation of a repetition without third sequences), digested and dropped; distances all-reduced; guide tree on the device; the pairs inside a <= 200-leaf subtree are recomputed, exchanged within the selectivity (NCCL) and relaxed.  value counts the all-pairs cells once (the recomputed pairs are extra work inside the same wall time)",
                       "l2_policy": "inputs larger than L2: each batch streams the dense DP layers of thousands of pairs"},
            "stage_ms_rank0": {k: round(last[k], 1) for k in last if k.startswith("t_")},
            "streamed_stage_kernel_ms_rank0": km,
            "streamed_matrices_cells_rank0": int(last["streamed_stage"]["nnz"]),
            "ingroup_pairs": last.get("ingroup_pairs_all_ranks"),
            "exchange_ms_rank0": last["exchange"]["ms_total"] if "exchange" in last else None,
            "hbm_used_gb_rank0": round((total_b - free_b) / 1e9, 1),
            "e2e": None,
            "e2e_note": "not measured for this workload: the streamed flow hands the matrices out batch by batch (digests here); a host consumer for them is the next step",
            "parity_digest": {"value": {"set_crc32": crc_streamed}, "ordinary_flow": ({"set_crc32": compare} if compare is not None else None),
                              "match": (crc_streamed == compare) if compare is not None else None,
                              "what": "CRC32 of the n x n per-matrix device digests (row pointers + cells of every matrix after the consistency repetition), summed over ranks; streamed flow vs the ordinary flow where that fits"},
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp32_issue", "stage_frac": total_cells * SLOTS_PER_CELL["stage_qp"] / max(world, 1) / (wall_ms * 1e-3) / FP32_ISSUE_PEAK,
                         "stage_frac_note": "430 algorithmic slots per cell x cells per GPU / wall time of the step / (148 SMs x 128 lanes x 1.965 GHz)"}}
    print(json.dumps(line))


def families_per_sec(seqs):
    """FASTA -> FASTA through the drop-in executables (process start, CUDA context, tail and file output included)."""
    out = {}
    exe = os.path.join(ROOT, "mlprobs_b200", "bin", "quickprobs_b200")
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "A.fa"); write_fasta(fa, seqs)
    t0 = time.time()
    r = subprocess.run([exe, fa, "-o", os.path.join(tmp, "A.out")], capture_output=True, text=True)
    dt = time.time() - t0
    ok = r.returncode == 0 and os.path.getsize(os.path.join(tmp, "A.out")) > 0 if os.path.exists(os.path.join(tmp, "A.out")) else False
    out["quickprobs_b200_config_A"] = {"families": 1, "seconds": dt, "families_per_sec": 1.0 / dt, "ok": bool(ok)}
    # persistent-process mode (csrc/serve.h, MLP_B200_SERVER=1): the second call finds the server and its CUDA context warm
    env = dict(os.environ, MLP_B200_SERVER="1", MLP_B200_SERVER_IDLE="20")
    subprocess.run([exe, fa, "-o", os.path.join(tmp, "A.out1")], capture_output=True, text=True, env=env)
    t0 = time.time()
    r = subprocess.run([exe, fa, "-o", os.path.join(tmp, "A.out2")], capture_output=True, text=True, env=env)
    dt = time.time() - t0
    same = os.path.exists(os.path.join(tmp, "A.out2")) and open(os.path.join(tmp, "A.out2"), "rb").read() == open(os.path.join(tmp, "A.out"), "rb").read()
    out["quickprobs_b200_config_A_persistent_process"] = {"families": 1, "seconds": dt, "families_per_sec": 1.0 / dt, "ok": r.returncode == 0,
                                                          "same_output_as_stand_alone": bool(same)}
    reg = os.path.join(ROOT, "tests", "golden", "regions", "inputs.tar.gz")
    if os.path.exists(reg):
        indir, outdir = os.path.join(tmp, "reg_in"), os.path.join(tmp, "reg_out")
        os.makedirs(indir); os.makedirs(outdir)
        with tarfile.open(reg) as tar:
            tar.extractall(indir, filter="data")
        nf = len(os.listdir(indir))
        t0 = time.time()
        r = subprocess.run([exe, indir, "-o", outdir], capture_output=True, text=True)
        dt = time.time() - t0
        out["quickprobs_b200_realignment_regions"] = {"families": nf, "seconds": dt, "families_per_sec": nf / dt, "ok": r.returncode == 0,
                                                      "note": "the 112 region files MLProbs' driver hands to quickprobs (BASELINE config #4), one process / one CUDA context"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="A")
    ap.add_argument("--ref-sample", type=int, default=144,
                    help="sequences of the workload the reference CPU arm aligns per step (144 -> 10,296 pairs, about 10 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cpnp object and the FASTA->FASTA runs (development)")
    ap.add_argument("--write-digest", action="store_true", help="record this run's parity digest as the committed expectation (N = 1 only)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import mlprobs_b200 as M
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    dev = local_rank
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    seqs = make_family(wl)
    n = len(seqs)
    eng = M.Engine(dev)
    h, p = M.default_tables(M.QP)
    eng.set_tables(h, p)
    eng.set_sequences(seqs)
    if world > 1:
        uid = [M.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)            # also selects this rank's shard of the cost-sorted pair list

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if wl.get("stream"):
        sampler = ClockSampler(dev); sampler.start()
        run_streamed(args, wl, M, torch, dist, eng, seqs, rank, world, dev, sampler)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    lens = np.array([len(s) for s in seqs], np.int64) + 1
    total_cells = int((lens.sum() ** 2 - (lens ** 2).sum()) // 2)
    npairs = n * (n - 1) // 2
    sampler = ClockSampler(dev); sampler.start()
    # ---- kernel-resident arm (inputs already in HBM)
    for _ in range(args.warmup):
        one_step(eng, M, n, False, world=world)
    barrier()
    ms_dev, launches, kms = [], 0, {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stats, _, _ = one_step(eng, M, n, False, world=world)
        ms_dev.append(sum(s["ms_total"] for _, s in stats))
        launches += sum(s["launches"] for _, s in stats)
        for name, s_ in stats:
            if name == "exchange":
                kms["exchange"] = kms.get("exchange", 0.0) + s_["ms_total"]
            else:
                for k, v in s_["ms_kernel"].items():
                    kms[k] = kms.get(k, 0.0) + v
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    dev_ms = float(np.mean(ms_dev))
    # ---- parity digest of the step's result (untimed): distances + every matrix of the relaxed set
    _, _, d_last = one_step(eng, M, n, False, world=world, want_dist=True)
    dg = torch.from_numpy(eng.set_digest().view(np.int64)).cuda()
    if dist is not None:
        dist.all_reduce(dg)                           # int64 wrap-around sum == sum mod 2^64 of the ranks' disjoint parts
    digest = {"distances_crc32": zlib.crc32(np.ascontiguousarray(d_last).tobytes()) & 0xffffffff,
              "set_crc32": zlib.crc32(dg.cpu().numpy().tobytes()) & 0xffffffff}
    # ---- end-to-end arm (host buffers in, host buffers out through the C ABI; every rank reads its own shard back)
    h2d = d2h = 0
    lay = eng.csr_layout()
    host_out = M.PinnedPackedBuffers(n, lay[1], int(lay[2] * 1.05))   # caller-owned page-locked result buffers, allocated once outside the timed region
    for _ in range(min(args.warmup, 2)):   # untimed: first use of the host->device / read-back path (allocations, page-locking)
        one_step(eng, M, n, True, seqs, world=world, host_out=host_out)
    eng.csr_packed_end()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stats, out, _ = one_step(eng, M, n, True, seqs, world=world, host_out=host_out)
        h2d = sum(s["h2d_bytes"] for _, s in stats) + sum(len(s) for s in seqs)
        d2h = n * n * 4 + eng.stats()["d2h_bytes"]   # distances + what the read-back enqueued (nz_off, nz_cnt, row sizes of the owned matrices, cells)
    eng.csr_packed_end()                       # the last step's read-back, inside the timed region
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if TRACE is not None and rank == 0:
        print("trace (ms, summed over all steps of both arms):", {k: round(v, 1) for k, v in TRACE.items()}, file=sys.stderr)
    sampler.stop_flag = True; sampler.join(timeout=2)
    if dist is not None:
        t = torch.tensor([wall_ms, e2e_ms, dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_ms, e2e_ms, dev_ms = [float(x) for x in t.tolist()]
        lt = torch.tensor([launches, h2d, d2h], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches, h2d, d2h = [int(x) for x in lt.tolist()]
        km = torch.tensor([kms.get(k, 0.0) for k in M.K_NAMES + ["exchange"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
        kms = {k: float(v) for k, v in zip(M.K_NAMES + ["exchange"], km.tolist())}
    eng.close()
    if rank == 0:
        models = 2
        value = total_cells * models / (wall_ms * 1e-3) / 1e9
        e2e = total_cells * models / (e2e_ms * 1e-3) / 1e9
        peaks = measured_peaks()
        per_kernel_ms = {k: v / args.steps for k, v in kms.items() if v}
        # dominant kernels: 5-state forward + backward sweeps (FP32-issue bound). Algorithmic work 358 lane-ops per cell,
        # algorithmic HBM bytes 12 per cell (forward layer written, read back, F+B written), see DESIGN.md.
        hmm_ms = per_kernel_ms.get("hmm_fwd", 0) + per_kernel_ms.get("hmm_bwd", 0)
        cells_rank = total_cells / max(world, 1)
        slot_rate = cells_rank * SLOTS_PER_CELL["hmm5"] / (hmm_ms * 1e-3) if hmm_ms else 0.0
        post_ms = sum(v for k, v in per_kernel_ms.items() if k not in ("relax", "exchange"))
        expected = None
        try:
            expected = json.load(open(DIGEST_FILE)).get(args.workload)
        except Exception:
            pass
        if args.write_digest and world == 1:
            allx = {}
            try:
                allx = json.load(open(DIGEST_FILE))
            except Exception:
                pass
            allx[args.workload] = digest
            json.dump(allx, open(DIGEST_FILE, "w"), indent=1)
            expected = digest
        line = {"metric": "pair_hmm_cell_updates_per_second", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": {"workload": wl["name"], "flavour": "quickprobs (5-state pair-HMM f32 + partition function f64, 1 consistency rep)",
                           "models_per_cell": models, "n": n, "pairs": npairs, "cells": total_cells,
                           "generator": "mlprobs_b200.synth.family_fast(p_sub=0.5): expected pairwise identity ~0.29",
                           "l2_policy": "inputs larger than L2: each batch streams >10 GB of dense DP layers (L2 = 126 MB)",
                           "multi_gpu": "pairs sharded round-robin over the cost-sorted list; distances all-reduced, then only the matrices within QuickProbs' selectivity are imported (mlp_exchange_needed); the relaxed set stays sharded" if world > 1 else "single GPU"},
                "device_ms_per_step": dev_ms, "kernel_ms_per_step": per_kernel_ms,
                "alignments_per_sec": npairs / (wall_ms * 1e-3),
                "gcups_posterior_stage": total_cells * models / max(world, 1) / (post_ms * 1e-3) / 1e9 if post_ms else None,
                "gcups_hmm5_fwd_bwd_per_gpu": cells_rank / (hmm_ms * 1e-3) / 1e9 if hmm_ms else None,
                "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                        "note": "bytes summed over ranks; every rank copies the family in and reads its own shard of the result back over its own PCIe link; the read-back of step k (mlp_get_csr_packed_begin) overlaps the posterior stage of step k+1 and is complete before that step's consistency stage starts, the last one is waited for inside the timed region"},
                "gpu_launches": int(launches),
                "parity_digest": {"value": digest, "expected": expected, "match": (digest == expected) if expected else None,
                                  "what": "CRC32 of the n x n distance matrix; CRC32 of the n x n per-matrix device digests (row pointers + cells of every matrix after the consistency repetition, summed over ranks)",
                                  "expected_source": "tests/golden/bench_digest.json (single-GPU run, bench.py --write-digest)"},
                "clocks": sampler.summary(),
                "roofline": {"bound": "fp32_issue",
                             "bound_note": "north_star and SURVEY 8d name FP32 instruction issue as the roofline of this path (log-space compare/select/polynomial work, no contraction for tensor cores); the HBM view of the same kernels is under roofline.hbm",
                             "kernel": "k_hmm_fwd_c+k_hmm_bwd_c", "achieved": slot_rate / 1e12, "peak": FP32_ISSUE_PEAK / 1e12,
                             "unit": "Tlane-op/s", "frac": slot_rate / FP32_ISSUE_PEAK,
                             "peak_source": "148 SMs x 128 FP32 lanes x 1.965 GHz (MEASURED_PEAKS.json sm_max_mhz); no measured FP32-issue peak exists in MEASURED_PEAKS.json",
                             "algorithmic_ops_per_cell": SLOTS_PER_CELL["hmm5"],
                             # the quantity north_star's 50 % target is stated on: the whole all-pairs posterior + consistency stage
                             "stage_frac": total_cells * SLOTS_PER_CELL["stage_qp"] / max(world, 1) / (wall_ms * 1e-3) / FP32_ISSUE_PEAK,
                             "stage_frac_note": "430 algorithmic slots per cell (SURVEY 8d: 358 HMM + 62 partition + 10 merge) x cells per GPU / wall time of the step (consistency, guide tree and exchange included) / peak",
                             "posterior_kernels_frac": cells_rank * SLOTS_PER_CELL["stage_qp"] / (post_ms * 1e-3) / FP32_ISSUE_PEAK if post_ms else None,
                             "traffic": cells_rank * NCU_DRAM_BYTES_PER_CELL_HMM5,
                             "traffic_unit": "DRAM bytes per step on this rank, both kernels, all of the step's launches (same scope as `achieved`)",
                             "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per cell measured at 192 x 300 (%s) x cells per step" % NCU_TRAFFIC_SOURCE,
                             "hbm": {"bound": "hbm", "achieved": cells_rank * 12.0 / (hmm_ms * 1e-3) / 1e9 if hmm_ms else None,
                                     "peak": peaks["hbm_gbs"] if peaks else 6650.0, "unit": "GB/s",
                                     "frac": (cells_rank * 12.0 / (hmm_ms * 1e-3) / 1e9) / (peaks["hbm_gbs"] if peaks else 6650.0) if hmm_ms else None,
                                     "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_cell": 12}}}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, seqs, args.ref_sample)
        if not args.no_extras and world == 1:
            line["cpnp"] = cpnp_object(M, dev, seqs, total_cells, 144)
            line["families_per_sec"] = families_per_sec(seqs)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(wl, seqs, n_s):
    cores = os.cpu_count() or 1
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_qp")
    sample = seqs[:n_s]
    if os.path.exists(exe):
        t, cells, _ = ref_qp_sample(seqs, n_s, cores)
        return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
                "sample": "first %d of %d sequences of the same family (%d pairs): reference PosteriorStage+tree+ConsistencyStage, %.2f s" % (n_s, len(seqs), n_s * (n_s - 1) // 2, t)}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    t0 = time.time()
    O.posterior_stage(O.QP, 3, ht, pt, sample, threads=cores)
    t = time.time() - t0
    cells = 2 * sum((len(a) + 1) * (len(b) + 1) for i, a in enumerate(sample) for b in sample[i + 1:])
    return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
            "sample": "first %d sequences, oracle port, posterior stage only, %.2f s" % (n_s, t)}


if __name__ == "__main__":
    main()
