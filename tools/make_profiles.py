#!/usr/bin/env python3
"""Turn the raw outputs of tools/gpu_round.sh (gpurun_out/, scratch) into the tracked summaries under
profiles/.   usage: make_profiles.py <tag> <round-name>      e.g.  make_profiles.py r01a r01"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True).stdout


# ---- launch list -------------------------------------------------------------------------------
src = os.path.join(G, f"launches_{tag}.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
agg = {}
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = d["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0, d["Grid Size"], d["Block Size"]])
    a[0] += 1
    a[1] += float(d["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{rnd}_launches.csv"), "w") as f:
    f.write("".join(l for l in open(src) if l.startswith('"')))
with open(os.path.join(P, f"{rnd}_launches_summary.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --load-index /tmp/ix --steps 5 --warmup 3 --no-cpu-baseline\n")
    f.write("(per-launch times under ncu are cold-cache and serialised; shares are what count)\n\n")
    f.write(f"{'launches':>8} {'total ms':>10} {'avg ms':>9} {'share':>7}  grid block  kernel\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e6:9.4f} {a[1] / tot * 100:6.1f}%  {a[2]} {a[3]}  {k}\n")
print(open(os.path.join(P, f"{rnd}_launches_summary.txt")).read())

# ---- full capture of the search kernel -----------------------------------------------------------
rep = os.path.join(G, f"search_{tag}.ncu-rep")
raw = list(csv.reader(io.StringIO(run("ncu", "-i", rep, "--page", "raw", "--csv"))))
h, u, v = raw[0], raw[1], raw[2]
d = dict(zip(h, v))
dram = (float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[h.index("dram__bytes_read.sum")]] +
        float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[h.index("dram__bytes_write.sum")]])
json.dump({"kernel": d["Kernel Name"], "dram_bytes_per_launch": dram, "source": f"profiles/{rnd}_search_kernel_ncu.txt",
           "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u[h.index("gpu__time_duration.sum")]]},
          open(os.path.join(P, "search_kernel_traffic.json"), "w"), indent=1)
with open(os.path.join(P, f"{rnd}_search_kernel_ncu.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 1  python bench.py --load-index /tmp/ix --steps 5 --warmup 3 --no-cpu-baseline\n\n")
    f.write(run(sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep))
    f.write("\nstall samples (smsp__pcsamp_warps_issue_stalled_*):\n")
    st = []
    for k in h:
        if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued"):
            try:
                st.append((int(d[k]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    ts = sum(x[0] for x in st) or 1
    for n, k in sorted(st, reverse=True)[:10]:
        f.write(f"  {n / ts * 100:5.1f}%  {k}\n")
    for k in ("sm__icc_request_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "lts__t_bytes.sum", "l1tex__t_bytes.sum"):
        if k in d:
            f.write(f"  {k} = {d[k]} {u[h.index(k)]}\n")
    nq = 10000
    f.write("\nexecuted warp instructions by opcode (per query, 10,000 queries per launch):\n")
    f.write(run(sys.executable, os.path.join(ROOT, "tools", "ncu_sass.py"), rep, str(nq)))
    f.write("\nexecuted warp instructions by source line:\n")
    f.write(run(sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"))
print(open(os.path.join(P, f"{rnd}_search_kernel_ncu.txt")).read()[:3000])
for name in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name.replace(tag, rnd)))
