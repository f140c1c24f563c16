#!/usr/bin/env python
"""Summarise an .ncu-rep (full set) and a launch-list csv into small tracked files under profiles/.
usage: python scripts/ncu_summary.py <tag> [gpurun_out/prof.ncu-rep] [gpurun_out/launches.csv]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof.ncu-rep"
lst = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/launches.csv"
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_drain", "smsp__pcsamp_warps_issue_stalled_imc_miss"]
try:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open("profiles/%s_ncu_full_summary.csv" % tag, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
        for k in KEEP:
            if k in idx:
                w.writerow([k, units[idx[k]]] + [r[idx[k]][:70] for r in rows[2:]])
    print(open("profiles/%s_ncu_full_summary.csv" % tag).read())
except Exception as exc:
    print("no full report:", exc)
try:
    rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t, n = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        try:
            t[r[ki][:90]] += float(r[vi].replace(",", ""))
            n[r[ki][:90]] += 1
        except ValueError:
            pass
    tot = sum(t.values())
    with open("profiles/%s_launch_shares.txt" % tag, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): shares only\n")
        for k, v in sorted(t.items(), key=lambda x: -x[1]):
            f.write("%-92s n=%4d %12.1f us %6.2f%%\n" % (k, n[k], v / 1e3, 100 * v / tot))
    print(open("profiles/%s_launch_shares.txt" % tag).read())
except Exception as exc:
    print("no launch list:", exc)
