#!/usr/bin/env python
"""Group the SASS lines of an ncu source-page CSV into runs with equal execution count and print
each run's share of executed instructions per destination pixel.
usage: tools/ncu_regions.py <source.csv> <dst pixels per launch> [min instr/px]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
px = float(sys.argv[2])
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
hdr = rows[1]
iS, iE, iSm, iA = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
data = []
for r in rows[2:]:
    if len(r) <= iE:
        continue
    try:
        e = int(r[iE])
    except ValueError:
        continue
    data.append((int(r[iA], 16) if r[iA].startswith('0x') else 0, r[iS], e, int(r[iSm] or 0)))
base = data[0][0]
prev, start, last, cnt, ops, smp = None, None, None, 0, collections.Counter(), 0


def flush():
    if prev is None or cnt * prev * 32 / px < thresh:
        return
    print("0x%04x-0x%04x n=%3d exec/line=%9d => %5.2f instr/px samples=%5d %s"
          % (start - base, last - base, cnt, prev, cnt * prev * 32 / px, smp, dict(ops.most_common(8))))


for a, sx, e, sm in data:
    if e != prev:
        flush()
        prev, start, cnt, ops, smp = e, a, 0, collections.Counter(), 0
    cnt += 1
    last = a
    smp += sm
    t = sx.split()
    ops[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]] += 1
flush()
