#!/usr/bin/env python3
"""bench.py -- all-pairs posterior + consistency throughput (BASELINE.json metric) on N B200s of one node.

A step = one pass of the hot path over one synthetic protein family: posterior stage (QuickProbs flavour:
5-state pair-HMM + FP64 partition function per cell -> merged posterior -> MEA distance -> CSR, both
orientations), host UPGMA tree, one consistency repetition (N > 50, as the reference), run through the C ABI.
`value` = pair-HMM cell updates per second (1 cell update = one (i,j) cell through forward+backward+posterior of ONE
model; the QuickProbs flavour runs 2 models per cell) over the whole step, inputs resident; `e2e` = same with the
host->device copy of the family and the device->host read-back of distances and CSR inside the timed region.
`--impl reference` times the compiled reference (oracle/_ref/ref_qp) on the host cores on a bounded sample.
"""
import argparse, json, os, subprocess, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the config the metric is quoted on
    "A": dict(n=1000, length=300, name="synthetic 1,000 protein seqs x len 300: all-pairs posterior + consistency"),
    "small": dict(n=200, length=300, name="synthetic 200 protein seqs x len 300 (development size)"),
}
FP32_ISSUE_PEAK = 148 * 128 * 1.965e9          # lane-instructions / s
SLOTS_PER_CELL = {"hmm5": 358.0, "part": 62.0, "merge": 10.0}   # algorithmic FP32-slot equivalents per cell


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
    return synth.family_fast(wl["n"], wl["length"], seed=20220148 + 2)


def run_reference(args, wl):
    """Reference arm: the UNMODIFIED QuickProbs CPU code (oracle/_ref/ref_qp) on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_qp")
    cores = os.cpu_count() or 1
    seqs = make_family(wl)
    # bounded sample of the same workload: the first n_s sequences, sized for ~10-20 s of CPU work per step
    n_s = int(args.ref_sample)
    sample = seqs[:n_s]
    tmp = os.path.join("/tmp", "mlp_ref_sample_%d.fa" % os.getpid())
    with open(tmp, "w") as f:
        for i, s in enumerate(sample):
            f.write(">s%05d\n%s\n" % (i, s.decode()))
    kind = "reference"
    if not os.path.exists(exe):
        kind = "port"
    times, cells = [], 0
    for it in range(args.warmup + args.steps):
        if kind == "reference":
            out = subprocess.run([exe, "bench", tmp, "--threads", str(cores)], capture_output=True, text=True).stdout
            j = json.loads(out.strip().splitlines()[-1])
            t = j["t_posterior_s"] + j["t_tree_s"] + j["t_relax_s"]; cells = j["cells"] * j["models"]
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            ht, pt = O.hmm_tables(), O.part_tables(O.QP)
            t0 = time.time()
            dist, S, _ = O.posterior_stage(O.QP, 3, ht, pt, sample, threads=cores)
            t = time.time() - t0
            cells = 2 * sum((len(a) + 1) * (len(b) + 1) for i, a in enumerate(sample) for b in sample[i + 1:])
        if it >= args.warmup:
            times.append(t)
    os.unlink(tmp)
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


# ncu --set full, main batch of tools/prof_run.py 192 300 (17,073 of the 18,336 pairs; the first 1,263 go to the small
# density-measuring batch): k_hmm_fwd 1.394 + 6.868 GB, k_hmm_bwd 7.421 + 6.852 GB of DRAM traffic -> 14.57 B per DP cell against
# 12 algorithmic (the slot layout pads 301 x 301 cells to 332 x 320 elements, x1.17)
NCU_DRAM_BYTES_PER_CELL_HMM5 = (1.393716 + 6.868421 + 7.421081 + 6.852403) * 1e9 / (17073 * 301 * 301)


def one_step(eng, M, n, e2e, seqs=None, world=1, read_back=True, host_out=None):
    """posterior stage [+ exchange] + host tree + consistency [+ exchange]. Returns the per-stage stats."""
    if e2e:
        eng.set_sequences(seqs)                       # host -> device copy of the family inside the timed region
        if world > 1:
            eng.set_shard(int(os.environ.get("RANK", "0")), world)
    stats = []
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    stats.append(("posterior", eng.stats()))
    if world > 1:
        eng.exchange_begin()                          # cell broadcasts run over NVLink while the host builds the tree
    d = eng.distances()                               # the guide tree is host work between the stages, as in the reference
    w, sd, _, _ = M.qp_guide_tree(d)
    w = np.maximum(w, np.float32(1e-6))
    if world > 1:
        eng.exchange_end(); stats.append(("exchange", eng.stats()))
    iters = 1 if n > 50 else 2
    for it in range(iters):
        cutoff = float(np.float32(0.01)) if it < iters - 1 else float(np.float32(1e-5))
        eng.relax(M.QP, w, sd, 200.0, 3.0, cutoff)
        stats.append(("relax", eng.stats()))
        if world > 1:
            eng.exchange(); stats.append(("exchange", eng.stats()))
    out = None
    if e2e and read_back:
        out = eng.csr_packed(host_out)                # device -> host read of the step's result (QuickProbs' own packed cell
                                                      # format, PackedSparseMatrix) into caller-owned page-locked buffers
    return stats, out


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
        stats, _ = one_step(eng, M, n, False, world=world)
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
    # ---- end-to-end arm (host buffers in, host buffers out through the C ABI)
    h2d = d2h = 0
    host_out = None
    if rank == 0:   # caller-owned page-locked result buffers, allocated once outside the timed region
        lay = eng.csr_layout()
        host_out = M.PinnedPackedBuffers(n, lay[1], int(lay[2] * 1.05))
    for _ in range(min(args.warmup, 2)):   # untimed: first use of the host->device / read-back path (allocations, page-locking)
        one_step(eng, M, n, True, seqs, world=world, read_back=(rank == 0), host_out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stats, out = one_step(eng, M, n, True, seqs, world=world, read_back=(rank == 0), host_out=host_out)
        h2d = sum(s["h2d_bytes"] for _, s in stats) + sum(len(s) for s in seqs)
        d2h = n * n * 4 + (out.nbytes() if out is not None else 0)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.stop_flag = True; sampler.join(timeout=2)
    if dist is not None:
        t = torch.tensor([wall_ms, e2e_ms, dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_ms, e2e_ms, dev_ms = [float(x) for x in t.tolist()]
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
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
        line = {"metric": "pair_hmm_cell_updates_per_second", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": {"workload": wl["name"], "flavour": "quickprobs (5-state pair-HMM f32 + partition function f64, 1 consistency rep)",
                           "models_per_cell": models, "n": n, "pairs": npairs, "cells": total_cells,
                           "generator": "mlprobs_b200.synth.family_fast(p_sub=0.5): expected pairwise identity ~0.29",
                           "l2_policy": "inputs larger than L2: each batch streams >10 GB of dense DP layers (L2 = 126 MB)"},
                "device_ms_per_step": dev_ms, "kernel_ms_per_step": per_kernel_ms,
                "alignments_per_sec": npairs / (wall_ms * 1e-3),
                "gcups_posterior_stage": total_cells * models / max(world, 1) / (sum(v for k, v in per_kernel_ms.items() if k not in ("relax", "exchange")) * 1e-3) / 1e9,
                "gcups_hmm5_fwd_bwd_per_gpu": cells_rank / (hmm_ms * 1e-3) / 1e9 if hmm_ms else None,
                "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": {"bound": "fp32_issue",
                             "bound_note": "north_star and SURVEY 8d name FP32 instruction issue as the roofline of this path (log-space compare/select/polynomial work, no contraction for tensor cores); the HBM view of the same kernels is under roofline.hbm",
                             "kernel": "k_hmm_fwd+k_hmm_bwd", "achieved": slot_rate / 1e12, "peak": FP32_ISSUE_PEAK / 1e12,
                             "unit": "Tlane-op/s", "frac": slot_rate / FP32_ISSUE_PEAK,
                             "peak_source": "148 SMs x 128 FP32 lanes x 1.965 GHz (MEASURED_PEAKS.json sm_max_mhz); no measured FP32-issue peak exists in MEASURED_PEAKS.json",
                             "algorithmic_ops_per_cell": SLOTS_PER_CELL["hmm5"],
                             # DRAM bytes (dram__bytes_read + write) of the two launches per step, from the ncu --set full capture
                             # profiles/r1c_all_kernels_ncu_full.txt: 14.57 B per cell (k_hmm_fwd + k_hmm_bwd) against
                             # 12 algorithmic, scaled to this workload's cells per launch
                             "traffic": cells_rank * NCU_DRAM_BYTES_PER_CELL_HMM5,
                             "traffic_unit": "DRAM bytes per step on this rank, both kernels, all of the step's launches (same scope as `achieved`)",
                             "traffic_source": "ncu dram bytes per cell measured at 192 x 300 (profiles/r1c_all_kernels_ncu_full.txt) x cells per step",
                             "hbm": {"bound": "hbm", "achieved": cells_rank * 12.0 / (hmm_ms * 1e-3) / 1e9 if hmm_ms else None,
                                     "peak": peaks["hbm_gbs"] if peaks else 6650.0, "unit": "GB/s",
                                     "frac": (cells_rank * 12.0 / (hmm_ms * 1e-3) / 1e9) / (peaks["hbm_gbs"] if peaks else 6650.0) if hmm_ms else None,
                                     "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_cell": 12}}}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, seqs, args.ref_sample)
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(wl, seqs, n_s):
    cores = os.cpu_count() or 1
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_qp")
    sample = seqs[:n_s]
    if os.path.exists(exe):
        tmp = os.path.join("/tmp", "mlp_cpu_sample_%d.fa" % os.getpid())
        with open(tmp, "w") as f:
            for i, s in enumerate(sample):
                f.write(">s%05d\n%s\n" % (i, s.decode()))
        out = subprocess.run([exe, "bench", tmp, "--threads", str(cores)], capture_output=True, text=True).stdout
        os.unlink(tmp)
        j = json.loads(out.strip().splitlines()[-1])
        t = j["t_posterior_s"] + j["t_tree_s"] + j["t_relax_s"]
        return {"value": j["cells"] * j["models"] / t / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
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
