#!/usr/bin/env python
"""Summarise an ncu report: one block of selected metrics per profiled kernel launch.

usage: ncu_summary.py <report.ncu-rep> [header line]     (reads the raw page through `ncu -i ... --page raw --csv`)
"""
import csv, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]

rep = sys.argv[1]
if len(sys.argv) > 2:
    print("# " + sys.argv[2] + "\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"{'Kernel Name':112s} {r[col['Kernel Name']][:110]}")
    for m in METRICS:
        if m in col:
            print(f"{m:95s} {units[col[m]]:16s} {r[col[m]]}")
    print()
