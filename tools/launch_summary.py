#!/usr/bin/env python3
"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel and grid size, launches, total, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, gi, vi, ui = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
big = sys.argv[2] if len(sys.argv) > 2 else None     # grid-size string that identifies the device-resident batch, e.g. 42624
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0][:60]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    key = (name, r[gi])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s grid %-16s launches %4d  total %10.1f us  share %5.1f%%  avg %8.1f us" % (name, grid, n, t, 100 * t / tot, t / n))
