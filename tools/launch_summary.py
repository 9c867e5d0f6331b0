#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py launches.csv [last_n_launches]"""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), us))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:6d} launches {us:12.1f} us {100 * us / total:5.1f}%")
print(f"{'total':60s} {len(rows):6d} launches {total:12.1f} us")
