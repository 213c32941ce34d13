#!/usr/bin/env python
"""Per-kernel summary of an ncu report (ncu --set full): one column per kernel (its LAST launch in the report),
one row per metric — the format of profiles/r*_ncu_*.csv that bench.py reads for roofline.traffic / issue_frac.

  python tools/ncu_summary.py report.ncu-rep "comment line" > profiles/r2_ncu_frame_lion4k.csv
"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

rep = sys.argv[1]
comment = sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
last = {}
order = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    if name not in last:
        order.append(name)
    last[name] = r
w = csv.writer(sys.stdout)
w.writerow(["# " + comment + "; one column per kernel (last launch of each in the report)"])
w.writerow(["metric", "unit"] + order)
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        w.writerow([m, units[i]] + [last[k][i] for k in order])
