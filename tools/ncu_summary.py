#!/usr/bin/env python
"""Excerpt of one kernel launch of an ncu report as JSON (what profiles/*_ncu_summary.json hold).

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_summary.py raw.csv "<kernel description>" "<command>" "<capture>" > profiles/<tag>_ncu_summary.json
"""
import csv
import json
import re
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
                  r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|launch__(block_size|grid_size|registers_per_thread|"
                  r"occupancy_limit_(registers|shared_mem)|shared_mem_per_block_static)|sm__inst_executed_pipe_(alu|fma|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"smsp__warps_eligible\.avg\.per_cycle_active)$")

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, values = rows[head], rows[head + 1], rows[head + 2]
metrics = {n: {"unit": u, "value": v} for n, u, v in zip(names, units, values) if KEEP.match(n)}
print(json.dumps({"kernel": sys.argv[2], "command": sys.argv[3], "capture": sys.argv[4], "metrics": dict(sorted(metrics.items()))}, indent=1))
