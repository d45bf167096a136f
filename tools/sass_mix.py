#!/usr/bin/env python3
"""ncu source page (SASS) -> executed-instruction mix of one kernel launch.

    ncu -i X.ncu-rep --page source --csv --kernel-name K --launch-count 1 > /tmp/k.csv ; python tools/sass_mix.py /tmp/k.csv [ncells]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iN:
        continue
    s = r[iS].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0] if s else "?"
    if not r[iN].isdigit():
        continue
    n = int(r[iN])
    mix[op] += n
    stall[op] += int(r[iSt] or 0)
    tot += n
ncell = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp instructions", tot, "" if not ncell else f"= {tot * 32 / ncell:.0f} thread instructions per cell")
ts = sum(stall.values()) or 1
for op, n in mix.most_common(28):
    print(f"{op:10s} {n:12d} {100.0 * n / tot:5.1f} %   samples {100.0 * stall[op] / ts:5.1f} %")
