#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (+ optional call order written
by tools/profile_step.py) into per-kernel and per-op shares.  Usage: summarize_launches.py CSV [call_order.txt]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    return {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}[row["Metric Unit"]]


tot = sum(us(r) for r in rows)
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares, not absolutes)\n")
agg = collections.OrderedDict()
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us(r)
print("by kernel:")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:9.1f} us  {k}")
if len(sys.argv) > 2:
    names = []
    for c in (l.strip() for l in open(sys.argv[2])):
        names += [c + "#rows", c + "#cols"] if c == "ffcb_rfft2" else ([c + "#cols", c + "#rows"] if c == "ffcb_irfft2" else [c])
    if len(names) == len(rows):
        agg = collections.OrderedDict()
        for n, r in zip(names, rows):
            a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us(r)
        print("\nby op (program order):")
        for k, (n, t) in agg.items():
            print(f"  {t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:9.1f} us  {k}")
