#!/usr/bin/env python3
"""Summarise `ncu -i rep --page source --csv --print-source sass,cuda` per CUDA source line: instructions executed and
stall samples (developer tool)."""
import csv, sys, re, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
per = collections.OrderedDict()
tot_inst = tot_samp = 0
hdr = None; fname = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No" and "Address" in r:
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)          # first "Source" = CUDA line
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        inst = int(float(r[ix["Instructions Executed"]] or 0)); samp = int(float(r[ix["# Samples"]] or 0))
        tinst = int(float(r[ix["Thread Instructions Executed"]] or 0))
    except ValueError:
        continue
    key = "%s:%s %s" % (fname, r[ix["Line No"]], re.sub(r"\s+", " ", r[ix["Source"]].strip())[:100])
    d = per.setdefault(key, [0, 0, 0, 0, 0, 0, 0])
    d[0] += inst; d[1] += samp; d[6] += tinst
    d[2] += int(float(r[ix["stall_long_sb"]] or 0)); d[3] += int(float(r[ix["stall_short_sb"]] or 0)) + int(float(r[ix["stall_mio"]] or 0))
    d[5] += int(float(r[ix["stall_wait"]] or 0)); d[4] += int(float(r[ix["stall_barrier"]] or 0)) if "stall_barrier" in ix else 0
    tot_inst += inst; tot_samp += samp
print("total inst %.3e samples %d" % (tot_inst, tot_samp))
for k, d in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6.2f%% inst (lanes %4.1f) %6.2f%% samp (long_sb %5d short/mio %5d wait %5d bar %5d) | %s"
          % (100.0 * d[0] / max(tot_inst, 1), d[6] / max(d[0], 1), 100.0 * d[1] / max(tot_samp, 1), d[2], d[3], d[5], d[4], k))
