#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and time per kernel.
usage: tools/ncu_launch_list.py <launches.csv> [title line]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
cnt, tot = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    cnt[r[iK]] += 1
    tot[r[iK]] += float(r[iV].replace(",", "")) * scale.get(r[iU], 1e-6)
if len(sys.argv) > 2:
    print(sys.argv[2])
print("(per-launch times under ncu are cold-cache and serialised: compare shares)")
print("%-72s %6s %12s %7s" % ("kernel", "count", "total ms", "share"))
total = sum(tot.values())
for k, t in tot.most_common():
    print("%-72s %6d %12.3f %6.1f%%" % (k[-72:], cnt[k], t, 100 * t / total))
