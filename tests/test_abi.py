"""The C-ABI library loads (no GPU needed) and exports exactly what include/jt_vm.h declares."""
import ctypes
import os
import re

import pytest

from common import ROOT
from joint_tensorf_b200 import _lib

HEADER = os.path.join(ROOT, "include", "jt_vm.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(jt_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.lib()
    assert lib.jt_version() >= 1
    assert lib.jt_strerror(0) == b"ok"
    assert b"invalid argument" in lib.jt_strerror(-1)


def test_every_declared_symbol_is_exported_with_matching_arity():
    decl = declared()
    assert len(decl) >= 17
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in decl.items():
        assert hasattr(cdll, name), f"{name} declared in jt_vm.h but not exported"
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == nargs, (name, len(_lib.SIGNATURES[name]), nargs)
    for name in _lib.SIGNATURES:
        assert name in decl, f"{name} bound in _lib.py but missing from jt_vm.h"


def test_argument_validation_without_gpu():
    """Entry points reject null pointers before touching the device."""
    lib = _lib.lib()
    rc = lib.jt_exclusive_scan(0, 0, 0, 0)
    assert rc == -1
    rc = lib.jt_blur_cl(0, 0, 0, 4, 4, 4, 0, 65, 3, 0, 0)
    assert rc == -1
    with pytest.raises(_lib.JtError):
        _lib.check(rc, "jt_blur_cl")
