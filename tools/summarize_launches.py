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
    # map launches to program ops: an FFT op is one launch (fused plane kernel) or two (row + column pass)
    calls = [l.strip() for l in open(sys.argv[2])]
    agg = collections.OrderedDict()
    i, ok = 0, True
    for c in calls:
        if i >= len(rows):
            ok = False
            break
        n = 1
        if c in ("ffcb_rfft2", "ffcb_irfft2") and "plane" not in rows[i]["Kernel Name"]:
            n = 2
        t = sum(us(r) for r in rows[i:i + n])
        i += n
        a = agg.setdefault(c, [0, 0.0]); a[0] += 1; a[1] += t
    if ok and i == len(rows):
        print("\nby op (program order):")
        for k, (n, t) in agg.items():
            print(f"  {t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:9.1f} us  {k}")
