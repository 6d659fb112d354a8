#!/usr/bin/env python
"""Per-kernel SASS evidence of libiqo_cuda.so: opcode counts that prove what the kernels are made of
(IDP = dp4a/dp2a, IMMA = mma.sync integer tensor path, LDSM/STSM = ldmatrix/stmatrix, LDGSTS = cp.async,
UTMALDG = TMA tensor copy, SYNCS = mbarrier), plus registers / spills from the ptxas log.
usage: tools/sass_histogram.py [libiqo_cuda.so] [ptxas.log] > profiles/rN_sass_histogram.txt   (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "libiqo_b200", "lib", "libiqo_cuda.so")
log = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "libiqo_b200", "lib", "ptxas.log")
sass = subprocess.check_output(["cuobjdump", "-sass", so]).decode()
KEYS = ["IDP.4A", "IDP.2A", "IMMA", "LDSM", "STSM", "LDGSTS", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "PRMT", "IMAD", "SHFL", "BAR"]
regs = {}
cur = None
for line in open(log):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        regs.setdefault(cur, [0, 0])[0] = int(m.group(1))
    m = re.search(r"(\d+) bytes spill stores", line)
    if m and cur:
        regs.setdefault(cur, [0, 0])[1] = int(m.group(1))
funcs = collections.OrderedDict()
name = None
arch = set(re.findall(r"arch = (sm_\w+)", sass))
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        funcs[name]["total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                funcs[name][k] += 1


def short(n):
    try:
        d = subprocess.check_output(["c++filt", n]).decode().strip()
    except Exception:
        d = n
    d = re.sub(r"iqo_b200::\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*\)$", "", d)
    return d.replace("void ", "")


print("SASS of %s: cubin arch %s, %d kernels" % (os.path.relpath(so, ROOT), ",".join(sorted(arch)), len(funcs)))
print("%-58s %5s %5s %6s " % ("kernel", "regs", "spill", "instr") + " ".join("%7s" % k for k in KEYS))
tot = collections.Counter()
fam = collections.OrderedDict()
for n, c in funcs.items():
    r = regs.get(n, [0, 0])
    print("%-58s %5d %5d %6d " % (short(n)[:58], r[0], r[1], c["total"]) + " ".join("%7d" % c[k] for k in KEYS))
    tot.update(c)
    f = short(n).split("<")[0]
    fam.setdefault(f, collections.Counter()).update(c)
    fam[f]["kernels"] += 1
print()
print("per kernel family:")
for f, c in fam.items():
    print("  %-28s x%-3d " % (f, c["kernels"]) + " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k]))
print()
print("whole library: " + " ".join("%s=%d" % (k, tot[k]) for k in KEYS))
print("no UTC*MMA / LDTM (tcgen05) and no HMMA: the tensor-path kernel uses the integer mma.sync path (IMMA.16816 / IMMA.16832), see DESIGN.md 4.8")
