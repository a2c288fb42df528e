#!/usr/bin/env python3
"""Short per-kernel table from an ncu report (developer tool): time, registers, occupancy, issue rate, instructions, pipes, DRAM, stalls."""
import subprocess, csv, io, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
S = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
cols = [("ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("Minst", "smsp__inst_executed.sum"), ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2hit", "lts__t_sector_hit_rate.pct"), ("long_sb", S % "long_scoreboard"), ("short_sb", S % "short_scoreboard"), ("wait", S % "wait"), ("math", S % "math_pipe_throttle"),
        ("mio", S % "mio_throttle"), ("lg", S % "lg_throttle"), ("notsel", S % "not_selected"), ("branch", S % "branch_resolving"), ("noinst", S % "no_instruction"), ("barrier", S % "barrier"), ("bankconf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
print("%-34s" % "kernel" + "".join("%9s" % c[0] for c in cols))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").replace("(KArgs)", "").replace("(RelaxArgs)", "")
    vals = []
    for c, k in cols:
        v = r[ix[k]] if k in ix else ""
        try:
            f = float(v.replace(",", ""))
            u = rows[1][ix[k]]
            if c == "Minst": f /= 1e6
            if c in ("rdMB", "wrMB"):
                f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            if c == "ms": f = f * {"us": 1e-3, "ms": 1, "s": 1e3, "ns": 1e-6}.get(u, 1)
            vals.append("%9.2f" % f if abs(f) < 1e5 else "%9.3g" % f)
        except ValueError:
            vals.append("%9s" % v[:8])
    print("%-34s" % name[:33] + "".join(vals))
