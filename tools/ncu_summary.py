#!/usr/bin/env python
"""Print the handful of ncu metrics that matter for the pass kernel: tools/ncu_summary.py <rep>"""
import csv, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
for k in keys:
    if k in d: print(f"{k:78s} {d[k][1]:>18s} {d[k][0]}")
for k in hdr:
    m = re.match(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", k)
    if m:
        try:
            v = float(d[k][1].replace(",", ""))
        except ValueError:
            continue
        if v > 0.08: print(f"  stall {k.split('stalled_')[1]:60s} {v:8.3f}")
