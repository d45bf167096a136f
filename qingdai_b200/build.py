"""In-tree build of libqd_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m qingdai_b200.build [--force]

-fmad=false keeps the reference's separate multiply/add rounding (per-step parity <= 1e-12);
-lineinfo lets ncu map SASS back to the kernel sources.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib", "libqd_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "--cudart", "static"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "qd_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: tuning variants (`--out _lib/x.so -DQD_FOO=1`, loaded with QD_B200_LIB=...); the product is OUT."""
    if out is None and not (force or _stale()):
        return OUT
    out = out or OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libqd_b200 cannot be built (and there is no CPU fallback)")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + list(defines) + [os.path.join(CSRC, "qd_api.cu"), "-o", out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out, defines=[a for a in sys.argv if a.startswith("-D")]))
