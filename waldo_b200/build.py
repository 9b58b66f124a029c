"""Build libwaldo_b200.so (sm_100a) in-tree with nvcc.  No torch dependency: the library is a plain C-ABI
shared object (include/waldo_b200.h) that the package loads with ctypes."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwaldo_b200.so")
SOURCES = ["waldo_abi.cu"]
HEADERS = ["wb_common.cuh", "wb_geom.cuh", "wb_prep.cuh", "wb_composite.cuh", "wb_composite_bwd.cuh", "wb_wif.cuh", "wb_pack.cuh", "wb_field.cuh", "wb_loss.cuh", "wb_conv.cuh",
           os.path.join("..", "..", "include", "waldo_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: waldo_b200 needs the CUDA toolkit to build its sm_100a kernels")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines` / `out` are for kernel experiments only (e.g. -DWB_NA_VARIANTS=1 into another file, loaded through
    the WALDO_B200_LIB environment variable); the package always builds and loads LIB."""
    if not force and not stale() and out == LIB:
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--use_fast_math=false",
           "-Xptxas", "-v" if verbose else "-O3",
           "-o", out] + [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [c for c in cmd if c != "--use_fast_math=false"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else LIB))
