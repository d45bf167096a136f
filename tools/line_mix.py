#!/usr/bin/env python3
"""ncu `--page source --print-source cuda,sass --csv` of one kernel -> executed instructions per CUDA source line.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name K --launch-count 1 > /tmp/k.csv
    python tools/line_mix.py /tmp/k.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, fpath, hdr = [], "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iN = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iN or not r[0].strip().isdigit():
        continue
    if r[iN].isdigit():
        out.append((int(r[iN]), fpath, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out) or 1
print("total", tot)
for n, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100.0 * n / tot:5.1f} %  {f}:{ln:<4d} {src}")
