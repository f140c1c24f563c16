#!/usr/bin/env python
"""Instruction mix + top stall sites from `ncu --page source --print-source sass --csv` output."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ie, isamp, isrc = ix["Instructions Executed"], ix["# Samples"], ix["Source"]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > max(ie, isamp) and r[ie].isdigit():
        data.append(r)
tot = sum(int(r[ie]) for r in data)
ts = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
print("static instrs", len(data), "executed", tot, "samples", ts)
op, ops = collections.Counter(), collections.Counter()
for r in data:
    s = r[isrc].strip().split()
    o = s[1] if s and s[0].startswith("@") else (s[0] if s else "?")
    o = o.split(".")[0]
    op[o] += int(r[ie])
    ops[o] += int(r[isamp]) if r[isamp].isdigit() else 0
for o, c in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print("%-10s exec %6.2f%%  samples %6.2f%%" % (o, 100 * c / tot, 100 * ops[o] / ts))
top = sorted(data, key=lambda r: -(int(r[isamp]) if r[isamp].isdigit() else 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]
for r in top:
    print(r[isamp], r[ie], r[isrc][:100])
