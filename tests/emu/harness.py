"""Kernel-logic emulation for the CPU test suite (TEST INFRASTRUCTURE, never a product path).

Compiles waldo_b200/csrc/waldo_abi.cu as plain C++ (-DWB_HOST_EMU: one host thread per CTA, see wb_common.cuh)
into tests/emu/libwaldo_emu.so and points the ctypes binding at it, so that the index arithmetic and formulas of
every kernel can be checked against the oracle without a GPU.  The package itself refuses such a library
(waldo_has_device_code() == 0) and refuses host tensors; only this harness flips those two switches.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
SRC = os.path.join(ROOT, "waldo_b200", "csrc")
LIB = os.path.join(HERE, "libwaldo_emu.so")


def build(force=False):
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "waldo_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    cmd = ["g++", "-x", "c++", "-std=c++17", "-DWB_HOST_EMU", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
           "-o", LIB, os.path.join(SRC, "waldo_abi.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stderr)
    return LIB


@contextlib.contextmanager
def emulated():
    """Route waldo_b200's ctypes calls to the host emulation library for the duration of the block."""
    from waldo_b200 import _lib
    lib = _lib._declare(C.CDLL(build()))
    assert lib.waldo_has_device_code() == 0
    saved = (_lib._lib, _lib._allow_host_pointers)
    _lib._lib, _lib._allow_host_pointers = lib, True
    try:
        yield lib
    finally:
        _lib._lib, _lib._allow_host_pointers = saved
