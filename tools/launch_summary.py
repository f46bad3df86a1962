#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/gpu_round.sh).
usage: launch_summary.py launches.csv "<command line>" > profiles/<tag>_launches_summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {k: hdr.index(k) for k in ("Kernel Name", "Block Size", "Grid Size", "Metric Value")}
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]) if "<" not in r[ix["Kernel Name"]] else r[ix["Kernel Name"]].split(">(")[0] + ">"
    key = (name, r[ix["Grid Size"]], r[ix["Block Size"]])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ix["Metric Value"]]) / 1e6
tot = sum(a[1] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(per-launch times under ncu are cold-cache and serialised; shares are what count; the list covers index load, ground truth, "
      "warm-up and the timed steps)\n")
print("launches   total ms    avg ms   share  grid block  kernel")
for (name, g, b), (n, ms) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{n:8d} {ms:10.3f} {ms / n:9.4f} {ms / tot * 100:6.1f}%  {g} {b}  {name}")
