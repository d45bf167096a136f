#!/usr/bin/env python3
"""cuobjdump -sass of the built library -> where the Blackwell-specific / protocol instructions sit (profiles/r02_sass_excerpt.txt).

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = collections.OrderedDict([
    (r"UTMALDG", "TMA box loads (cp.async.bulk.tensor.2d) of the sigma = 1 Gaussian tiles"),
    (r"SYNCS", "mbarrier init / expect_tx / try_wait of those loads"),
    (r"LDGSTS", "cp.async row staging of the opt-in fused ocean sub-step"),
    (r"LDCU\.(128|64) UR\d+, c\[0x3\]", "constant-bank coefficients of qd_exp / qd_tanh (csrc/qd_math.cuh) instead of UMOV immediates"),
    (r"STG\.E\.128\.STRONG\.SYS", "flagged 16-byte lines of the band transport: value + arrival flag in ONE posted store"),
    (r"LDG\.E\.128\.STRONG\.SYS", "polling of flagged lines in the receiver's own memory"),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "qingdai_b200", "_lib", "libqd_b200.so")], capture_output=True, text=True, check=True).stdout
    hits = collections.OrderedDict((k, collections.OrderedDict()) for k in WANT)
    func = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            func = m.group(1)
            continue
        for k in WANT:
            if func and re.search(k, ln):
                ins = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().rstrip(";").strip()
                hits[k].setdefault(func, [0, ins])[0] += 1
    names = subprocess.run(["c++filt"], input="\n".join({f for h in hits.values() for f in h}), capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip({f for h in hits.values() for f in h}, names)) if names else {}
    print("cuobjdump -sass qingdai_b200/_lib/libqd_b200.so (sm_100a, end of round 2): where the Blackwell-specific / protocol instructions sit.")
    print("Per pattern: occurrences, kernel, first occurrence.  Made by tools/sass_excerpt.py.\n")
    for k, desc in WANT.items():
        print(f"== {k}: {desc}")
        if not hits[k]:
            print("   (none)")
        for fn, (n, ins) in hits[k].items():
            print(f"   {n:4d}  {dem.get(fn, fn)[:110]}\n         {ins}")
        print()


if __name__ == "__main__":
    main()
