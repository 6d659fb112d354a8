#!/usr/bin/env python
"""Print the SASS lines of an ncu source-page CSV between two offsets with their execution counts and stall samples.
usage: tools/ncu_sass.py <source.csv> <from offset hex> <to offset hex>"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
hdr = rows[1]
iS, iE, iSm, iA = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
base = None
for r in rows[2:]:
    if len(r) <= iE or not r[iA].startswith('0x'):
        continue
    a = int(r[iA], 16)
    if base is None:
        base = a
    if lo <= a - base <= hi:
        print("0x%04x %9s %5s  %s" % (a - base, r[iE], r[iSm], r[iS]))
