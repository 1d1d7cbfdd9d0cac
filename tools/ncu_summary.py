#!/usr/bin/env python
"""Markdown table of the counters profiles/ quotes from an `ncu --set full` capture.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [min_ms]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
M = [("duration", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs/thread", "launch__registers_per_thread"),
     ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
     ("ALU (integer) pipe % of peak", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     ("FMA pipe % of peak", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
     ("LSU pipe % of peak", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("shared-memory wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
     ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     ("L2 throughput % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
     ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
     ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("warp instructions", "smsp__inst_executed.sum"),
     ("active threads / instruction", "smsp__thread_inst_executed_per_inst_executed.ratio")]
keep = [r for r in data if float(r[col["gpu__time_duration.sum"]]) >= min_ms]
names = []
for r in keep:
    n = re.sub(r"^void ", "", r[col["Kernel Name"]])
    names.append("`" + re.sub(r"\(.*", "", n) + "`")
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(keep))
for label, key in M:
    if key not in col:
        continue
    u = units[col[key]]
    vals = []
    for r in keep:
        v = r[col[key]]
        try:
            vals.append("%.4g" % float(v))
        except ValueError:
            vals.append(v)
    print("| %s (%s) | %s |" % (label, u, " | ".join(vals)))
