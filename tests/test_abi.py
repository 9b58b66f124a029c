"""The C-ABI boundary without a GPU: the sm_100a library builds/loads, exports every symbol include/waldo_b200.h
declares, the ctypes mirrors match the header field by field, and the product never touches the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "waldo_b200.h")


def _header():
    txt = open(HEADER).read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def test_library_loads_and_exports_every_declared_symbol():
    from waldo_b200 import _lib
    lib = _lib.load()
    declared = set(re.findall(r"\b(waldo_[a-z_0-9]+)\s*\(", _header()))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.waldo_abi_version() == _lib.ABI_VERSION
    assert lib.waldo_has_device_code() == 1
    assert lib.waldo_last_error() is not None


def test_ctypes_structs_mirror_the_header():
    from waldo_b200 import _lib
    txt = _header()
    ctype_of = {"int": C.c_int, "float": C.c_float, "int64_t": C.c_longlong}
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(\w+);", txt, flags=re.S):
        body, name = m.group(1), m.group(2)
        st = _lib.STRUCT_OF[name]
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            is_ptr = "*" in decl
            decl = decl.replace("const ", "").replace("*", " ")
            toks = decl.replace(",", " ").split()
            typ, names = toks[0], toks[1:]
            for n in names:
                fields.append((n, typ, is_ptr))
        got = [f[0] for f in st._fields_]
        assert got == [f[0] for f in fields], (name, got, [f[0] for f in fields])
        for (fname, ftype), (_, typ, is_ptr) in zip(st._fields_, fields):
            if is_ptr:
                assert ftype is C.c_void_p, (name, fname)
            elif typ in ctype_of:
                assert ftype is ctype_of[typ], (name, fname)
            else:
                assert ftype is _lib.STRUCT_OF[typ], (name, fname)
    assert (_lib.MAX_LAYERS, _lib.MAX_CH, _lib.MAX_LYT, _lib.MAX_TPS_K) == tuple(
        int(re.search(rf"#define {k}\s+(\d+)", txt).group(1)) for k in ("WALDO_MAX_LAYERS", "WALDO_MAX_CH", "WALDO_MAX_LYT", "WALDO_MAX_TPS_K"))


def test_invalid_arguments_fail_with_a_message():
    from waldo_b200 import _lib
    lib = _lib.load()
    a = _lib.TpsFwd(1, 300, 16, None, None, None, None, None)
    rc = lib.waldo_tps_fwd(C.byref(a), None)
    assert rc == -1 and b"control points" in lib.waldo_last_error()
    assert lib.waldo_occ_fwd(1, 0, None, None, None) == -1


def test_product_never_imports_the_oracle_or_a_cpu_path():
    pkg = os.path.join(ROOT, "waldo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "waldo_oracle" not in src and "ref_loader" not in src, f
                if f.endswith(".py"):
                    assert "tests.emu" not in src and "libwaldo_emu" not in src, f


def test_cpu_tensors_are_refused():
    import torch
    import waldo_b200 as wb
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        wb.compute_occ(torch.zeros(1, 2, 3))
