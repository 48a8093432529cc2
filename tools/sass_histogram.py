#!/usr/bin/env python
"""
tools/sass_histogram.py -- opcode histogram per kernel of setk_b200/libsetk_b200.so
(`cuobjdump -sass`), the evidence of which hardware paths each kernel uses:
  HMMA = mma.sync tensor cores, UBLKCP = TMA bulk copies, SYNCS = mbarriers, LDGSTS = cp.async,
  USETMAXREG = per-role register budgets, LDTM / STTM = tensor-memory loads / stores (tcgen05.ld / st), FFMA2/FADD2/FMUL2 = packed fp32, DFMA = fp64.

    python tools/sass_histogram.py [lib] > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "setk_b200", "libsetk_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(setk::\w+Args.*\)|\(.*\)$", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["HMMA", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "LDTM", "STTM", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DMMA",
       "LDS", "STS", "BAR", "STL", "LDL"]
print(f"# opcode histogram of {os.path.relpath(lib, ROOT)} (cuobjdump -sass); static instruction counts")
print(f"# {'kernel':70s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEY))
tot = collections.Counter()
for k, h in hist.items():
    n = sum(h.values())
    if n < 40:
        continue
    print(f"{k[:72]:72s} {n:7d} " + " ".join(f"{h.get(x, 0):7d}" for x in KEY))
    tot.update(h)
print(f"{'ALL KERNELS':72s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(x, 0):7d}" for x in KEY))
