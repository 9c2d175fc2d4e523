#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of an `ncu --page source --csv` export to
CUDA source lines, using the line table of the same kernel from `nvdisasm -g -c`.

usage: ncu_by_line.py <source_page.csv> <nvdisasm_all.txt> <mangled kernel name> [file.cu]"""
import csv
import re
import sys
import collections


def main(csv_path, dis_path, kern, only=None):
    rows = list(csv.reader(open(csv_path)))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    per_addr = {}
    for r in rows[2:]:
        try:
            per_addr[int(r[ix["Address"]], 16)] = (float(r[ix["Instructions Executed"]]), float(r[ix["# Samples"]]),
                                                   float(r[ix["stall_no_inst"]]))
        except ValueError:
            pass
    base = min(per_addr)
    lines = open(dis_path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
    cur = None
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("\t.section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
        if m and cur:
            a = base + int(m.group(1), 16)
            if a in per_addr:
                e = agg[cur]
                e[0] += per_addr[a][0]
                e[1] += per_addr[a][1]
                e[2] += per_addr[a][2]
                e[3] += 1
    tot = sum(v[0] for v in agg.values())
    smp = sum(v[1] for v in agg.values())
    print(f"total warp instructions {tot:.0f}, samples {smp:.0f}, static instrs {sum(v[3] for v in agg.values())}")
    for (f, ln), v in sorted(agg.items()):
        if only and f != only:
            continue
        print(f"{f}:{ln:4d}  inst {100 * v[0] / tot:5.2f}%  samples {100 * v[1] / smp:5.2f}%  no_inst {v[2]:7.0f}  static {v[3]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
