#!/usr/bin/env python
"""Aggregates an ncu SASS source page by CUDA source line (ncu's own --print-source cuda CSV carries no metrics).

  ncu -i prof.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all prototype2_b200/libprtc.so && nvdisasm --print-line-info kernels.sm_100a.cubin > kernels.sass
  python tools/ncu_by_line.py sass.csv k_step_stream.sass k_step_stream [top]
"""
import csv
import re
import sys
from collections import defaultdict


def line_map(sass_path, func_pat):
    amap, cur, on = {}, None, False
    for ln in open(sass_path, errors="ignore"):
        if ln.startswith("//----") and ".text." in ln:
            on = func_pat in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return amap


def main():
    sass_csv, sass_txt, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(sass_txt, pat)
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    col = {h: i for i, h in enumerate(rows[hi])}
    agg = defaultdict(lambda: defaultdict(float))
    base = None
    for r in rows[hi + 1:]:
        if len(r) < len(col):
            continue
        try:
            addr = int(r[col["Address"]], 16) if not r[col["Address"]].isdigit() else int(r[col["Address"]])
        except ValueError:
            continue
        if base is None:
            base = addr
        key = amap.get(addr - base, (("?", 0), ""))[0] or ("?", 0)
        a = agg[key]
        for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_wait",
                  "stall_short_sb", "stall_branch_resolving", "stall_not_selected", "stall_no_inst", "stall_math"):
            try:
                a[k] += float(r[col[k]])
            except (ValueError, KeyError):
                pass
    ts = sum(a["# Samples"] for a in agg.values()) or 1
    ti = sum(a["Instructions Executed"] for a in agg.values()) or 1
    print("total samples %d, warp instructions %d, avg threads/inst %.2f" % (
        ts, ti, sum(a["Thread Instructions Executed"] for a in agg.values()) / ti))
    print("%-22s %7s %7s %6s %7s %7s %7s %7s" % ("file:line", "smp%", "inst%", "thr", "long_sb", "wait", "shortsb", "branch"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        ie = a["Instructions Executed"] or 1
        print("%-22s %6.2f%% %6.2f%% %6.1f %7d %7d %7d %7d" % (
            "%s:%d" % key, 100 * a["# Samples"] / ts, 100 * a["Instructions Executed"] / ti,
            a["Thread Instructions Executed"] / ie, a["stall_long_sb"], a["stall_wait"], a["stall_short_sb"],
            a["stall_branch_resolving"]))


if __name__ == "__main__":
    main()
