"""Host-compiled check build of the kernel sources (TEST SCAFFOLDING, not a product path).

``qingdai_b200/csrc/*.cu[h]`` is compiled with g++ and ``-DQD_HOST_EMU`` so that every kernel runs
one "thread" at a time on the CPU.  CPU-only CI uses it to catch indexing / operand-order /
boundary-semantics mistakes before GPU minutes are spent.  ``qingdai_b200`` never imports or loads
it; the product raises when the CUDA library or a CUDA device is missing.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qingdai_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libqd_hostcheck.so")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "qd_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False):
    if force or _stale():
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DQD_HOST_EMU",
               "-x", "c++", os.path.join(CSRC, "qd_api.cu"), "-o", OUT]
        subprocess.run(cmd, check=True)
    return OUT


def library():
    from qingdai_b200._binding import Library
    return Library(build(), host_emulation=True)
