#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one line per launch, or
(with --agg) total time per kernel name."""
import csv
import re
import sys
import collections

path = sys.argv[1]
agg = "--agg" in sys.argv
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")
    us = float(r[vi].replace(",", "")) / 1e3
    if agg:
        k = (name, r[gi])
        tot[k] = (tot.get(k, (0, 0))[0] + us, tot.get(k, (0, 0))[1] + 1)
    else:
        print(f"{name:48s} {r[gi]:20s} {r[bi]:14s} {us:10.1f} us")
if agg:
    for (name, grid), (us, n) in tot.items():
        print(f"{name:48s} {grid:20s} n={n:4d} total {us:10.1f} us  avg {us / n:9.1f} us")
