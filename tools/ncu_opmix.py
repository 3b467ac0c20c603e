#!/usr/bin/env python3
"""Opcode mix of one kernel from an .ncu-rep captured with --import-source on: executed warp instructions and stall
samples per opcode.  usage: ncu_opmix.py file.ncu-rep kernel-regex [top]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
lines = out.splitlines()
# the report may hold several launches: take the first table only
tables, cur = [], None
for l in lines:
    if l.startswith('"Kernel Name"'):
        cur = []; tables.append((l, cur))
    elif cur is not None:
        cur.append(l)
name, body = tables[0]
rows = list(csv.reader(body))
h = rows[0]
ix, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= ie: continue
    toks = r[ix].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "STG", "LDS", "MUFU", "F2I", "I2F", "F2F")) and "." in op else "")
    n = int(r[ie] or 0)
    ops[op] += n; tot += n; samp[op] += int(r[isamp] or 0)
print(name[:160])
print(f"total warp instructions {tot}, static {len(rows)-1}")
ts = sum(samp.values()) or 1
for op, n in ops.most_common(top):
    print(f"  {op:14s} {n:12d} {100.0*n/tot:6.2f} %   samples {100.0*samp[op]/ts:6.2f} %")
