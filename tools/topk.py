#!/usr/bin/env python3
"""Print ms/step and the per-kernel table of one or more bench JSON lines side by side."""
import json, sys
recs = []
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith("{"):
            recs.append((f, json.loads(l)))
names = []
for _, d in recs:
    for t in d["roofline"]["top_kernels"]:
        if t["kernel"] not in names:
            names.append(t["kernel"])
print("%-28s" % "ms/step", *["%14.4f" % d["ms_per_step"] for _, d in recs])
print("%-28s" % "e2e", *["%14.4f" % d["e2e"]["value"] for _, d in recs])
for n in names:
    row = []
    for _, d in recs:
        t = next((t for t in d["roofline"]["top_kernels"] if t["kernel"] == n), None)
        row.append("%5.2fx%7.1f" % (t["launches_per_step"], t["us_per_launch"]) if t else " " * 13)
    print("%-28s" % n[:28], *["%14s" % r for r in row])
