#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` per source line: instructions executed and stall samples (developer tool)."""
import csv, sys, re, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# locate header row
for hi, r in enumerate(rows):
    if r and r[0] == "Address":
        break
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_samp = 0
per = collections.OrderedDict()
view = "sass"
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        inst = int(float(r[ix["Instructions Executed"]] or 0)); samp = int(float(r[ix["# Samples"]] or 0))
    except ValueError:
        continue
    key = r[ix["Source"]].strip()
    key = re.sub(r"\s+", " ", key)[:110]
    d = per.setdefault(key, [0, 0, 0, 0, 0.0, 0])
    d[0] += inst; d[1] += samp
    d[2] += int(float(r[ix["stall_long_sb"]] or 0)); d[3] += int(float(r[ix["stall_short_sb"]] or 0)) + int(float(r[ix["stall_mio"]] or 0))
    d[5] += int(float(r[ix["stall_wait"]] or 0))
    tot_inst += inst; tot_samp += samp
print("total inst %.3e samples %d" % (tot_inst, tot_samp))
for k, d in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6.2f%% inst %6.2f%% samp (long_sb %5d short/mio %5d wait %5d) | %s" % (100.0 * d[0] / max(tot_inst, 1), 100.0 * d[1] / max(tot_samp, 1), d[2], d[3], d[5], k))
