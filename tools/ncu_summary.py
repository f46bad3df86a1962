#!/usr/bin/env python3
"""Key numbers of an ncu report (one line per profiled kernel): duration, instruction count, issue
utilisation, occupancy, DRAM/L2 traffic, top stall reasons.  usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
            "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
            "smsp__inst_executed_op_local_st.sum", "derived__smsp__inst_executed_op_branch_pct"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "")[:120])
        for k in keys:
            if k in d and d[k] != "":
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
        stalls = []
        for k in hdr:
            if "average_warp_latency_issue_stalled" in k or ("warp_issue_stalled" in k and k.endswith("per_warp_active.pct")):
                try:
                    stalls.append((float(d[k]), k))
                except ValueError:
                    pass
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {v:8.2f}  {k}")


if __name__ == "__main__":
    main()
