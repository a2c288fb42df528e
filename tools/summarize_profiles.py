#!/usr/bin/env python3
"""Turn ncu reports / launch lists from gpurun_out/ into the small text summaries committed under profiles/."""
import csv, subprocess, sys, os, collections, io

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "sm__cycles_elapsed.avg.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarize_full(rep, dst, note):
    hdr, units, rows = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# %s\n# source: ncu --set full --clock-control none --import-source on (report %s)\n" % (note, os.path.basename(rep)))
        for r in rows:
            f.write("\n== %s\n" % r[ix["Kernel Name"]])
            for k in KEYS:
                if k in ix:
                    f.write("  %-88s %s %s\n" % (k, r[ix[k]], units[ix[k]]))


def summarize_launches(csvpath, dst, note):
    rows = [r for r in csv.reader(open(csvpath)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict(); order = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        d = per.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += ms
        order.append((name, ms))
    tot = sum(d[1] for d in per.values())
    with open(dst, "w") as f:
        f.write("# %s\n# source: ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare shares)\n" % note)
        f.write("%-16s %8s %12s %8s\n" % ("kernel", "launches", "total_ms", "share"))
        for k, d in per.items():
            f.write("%-16s %8d %12.3f %7.1f%%\n" % (k, d[0], d[1], 100 * d[1] / tot))
        f.write("\n# launch list (in order)\n")
        for name, ms in order:
            f.write("%-16s %10.4f ms\n" % (name, ms))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "full":
        summarize_full(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        summarize_launches(sys.argv[2], sys.argv[3], sys.argv[4])
