#!/usr/bin/env python3
"""Key metrics of an .ncu-rep (read here, no GPU needed):  ncu_summary.py file.ncu-rep [more-metric-substrings...]"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct", "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__inst_executed_pipe_xu", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg", "smsp__average_warp",
        "smsp__warp_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__inst_executed_pipe_alu",
        "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
hdr, units = rows[0], rows[1]
keep = KEEP + sys.argv[2:]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print(f"## {d['Kernel Name']}  grid {d['Grid Size']} block {d['Block Size']}")
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in keep) and v not in ("", "0", "n/a"):
            if "warp_issue_stalled" in h and "_per_warp_active.pct" not in h:
                continue
            print(f"  {h} [{u}] = {v}")
