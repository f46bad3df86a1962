#!/usr/bin/env python3
"""SASS-level view of an ncu report: executed warp instructions per opcode and the hottest basic
blocks.  usage: ncu_sass.py report.ncu-rep [nq]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
nq = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
ins = []
for r in rows:
    if "Address" in r and "Source" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            ins.append((d["Address"], d["Source"].strip(), int(d["Instructions Executed"]), int(d["# Samples"]),
                        float(d.get("Avg. Threads Executed", "0") or 0)))
        except ValueError:
            pass
tot = sum(i[2] for i in ins)
print(f"SASS instructions {len(ins)}  executed {tot}  per query {tot / nq:.0f}")
byop = collections.Counter()
for a, s, n, sm, th in ins:
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)?)", s)
    byop[m.group(1) if m else s[:10]] += n
for op, n in byop.most_common(28):
    print(f"  {op:18s} {n / nq:9.0f}/q {n / tot * 100:5.1f}%")
if len(sys.argv) > 3:
    print("--- listing (addr, exec/q, samples, threads, sass)")
    for a, s, n, sm, th in ins:
        print(f"{a[-5:]} {n / nq:8.1f} {sm:6d} {th:5.1f}  {s[:100]}")
