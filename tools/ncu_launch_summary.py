#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel
totals of the LAST bench step (one step starts at bbox_kernel).  Times are cold-cache
and serialised under ncu; compare shares, not absolutes."""
import collections
import csv
import re
import sys


def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "bbox_kernel" in n]
    last = rows[starts[-1]:] if starts else rows
    agg = collections.OrderedDict()
    for r in last:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"<unnamed>::", "", n)[:72]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    o = open(out, "w") if out else sys.stdout
    print(f"# source: {path}\n# steps seen: {len(starts)}; launches in last step: {len(last)}; "
          f"sum of kernel time: {tot:.3f} ms", file=o)
    print(f"{'ms':>10} {'share':>7} {'launches':>9}  kernel", file=o)
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1]:10.3f} {100 * a[1] / tot:6.1f}% {a[0]:9d}  {n}", file=o)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
