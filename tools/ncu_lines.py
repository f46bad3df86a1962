#!/usr/bin/env python3
"""Aggregate an ncu report's executed warp instructions and stall samples by CUDA source line.
usage: ncu_lines.py report.ncu-rep [top_n]   (needs kernels built with -lineinfo and --import-source on)"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, out = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            ie = hdr.index("Instructions Executed")
            ns = hdr.index("# Samples")
            try:
                out.append((int(r[ie]), int(r[ns]), cur, r[0], r[1].strip()[:100]))
            except ValueError:
                pass
    tot = sum(o[0] for o in out) or 1
    tots = sum(o[1] for o in out) or 1
    print(f"total warp instructions {tot}  samples {tots}")
    byf = {}
    for o in out:
        byf[o[2]] = byf.get(o[2], 0) + o[0]
    print({k: f"{v / tot * 100:.1f}%" for k, v in byf.items()})
    print(" inst%  smpl%  file:line  source")
    for o in sorted(out, reverse=True)[:top]:
        print(f"{o[0] / tot * 100:5.2f}  {o[1] / tots * 100:5.2f}  {o[2]}:{o[3]:>4}  {o[4]}")


if __name__ == "__main__":
    main()
