#!/usr/bin/env python3
"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share.
usage: launch_summary.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u, 1e-6)
    name = r[ik].split("(")[0]
    tot[name][0] += 1
    tot[name][1] += ms
all_ms = sum(v[1] for v in tot.values())
for k, (n, ms) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k:32s} launches={n:5d} total_ms={ms:10.3f} share={100 * ms / all_ms:5.1f}%")
print(f"{'all':32s} total_ms={all_ms:10.3f}")
