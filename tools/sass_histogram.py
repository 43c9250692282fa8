#!/usr/bin/env python
"""SASS opcode histogram of every kernel in libhedit_b200.so (cuobjdump -sass): the Blackwell-native evidence per kernel (UTCHMMA =
tcgen05.mma, UTMALDG / UTMASTG = TMA, LDTM / STTM = tcgen05.ld / st, MUFU, ...).  python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hedit_b200", "libhedit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU", "HMMA", "FFMA2", "FADD2", "FMNMX3", "F2FP", "STS", "LDS",
       "LDG", "STG", "STL", "LDL", "BAR"]
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op] += 1
        total[op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist.keys()), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(so)}: {len(hist)} kernels, {sum(total.values())} SASS instructions")
print("# library totals: " + ", ".join(f"{k} {total[k]}" for k in KEY if total[k]))
for (k, h), name in zip(hist.items(), demangle):
    n = sum(h.values())
    name = re.sub(r"\(hedit::\w+Params\)|\(.*\)$", "", name)[:150]
    print(f"{name}: {n} instr | " + ", ".join(f"{op} {h[op]}" for op in KEY if h[op]))
