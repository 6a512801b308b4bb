#!/usr/bin/env python3
"""Aggregate ncu source-page warp-stall samples by SASS address region.
usage: ncu_regions.py src.csv [boundary_hex ...]   (boundaries are offsets from the first instruction)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, isrc, iall, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = rows[2:]
base = int(data[0][ia], 16)
bounds = [int(x, 16) for x in sys.argv[2:]] + [1 << 40]
reg = {}
for r in data:
    off = int(r[ia], 16) - base
    k = next(i for i, b in enumerate(bounds) if off < b)
    d = reg.setdefault(k, {"samples": 0, "inst": 0, "n": 0, "stalls": {}})
    d["samples"] += int(r[iall]); d["inst"] += int(r[iex]); d["n"] += 1
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: d["stalls"][h[i]] = d["stalls"].get(h[i], 0) + v
tot = sum(d["samples"] for d in reg.values())
lo = 0
for k in sorted(reg):
    d = reg[k]
    top = sorted(d["stalls"].items(), key=lambda x: -x[1])[:6]
    print(f"region {k} [0x{lo:x}, 0x{min(bounds[k], 1<<20):x}): sass={d['n']} samples={d['samples']} ({100*d['samples']/tot:.1f}%) inst_exec={d['inst']}  " +
          " ".join(f"{a[6:]}={100*b/max(1,d['samples']):.0f}%" for a, b in top))
    lo = bounds[k]
