"""CPU-side checks of the C-ABI boundary: the library builds/loads here (no GPU) and exports every
symbol include/sparf_b200.h declares; no compute call is made."""
import os
import re

from sparf_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sparf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sparf_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    L = _lib.lib()
    declared = _header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(L, name), "library does not export %s" % name
    # the ctypes table and the header must name the same entry points
    assert sorted(_lib.exported_symbols()) == declared


def test_version_and_error_string():
    L = _lib.lib()
    assert L.sparf_version() == 100
    assert isinstance(L.sparf_last_error(), (bytes, type(None)))


def test_struct_layout_matches_header():
    # SparfMLP: 7 int32 + 2 float + pointer + (12+12+2+2) pointers; natural alignment on LP64
    import ctypes
    assert ctypes.sizeof(_lib.SparfMLP) == 40 + 8 * (1 + 28)
    assert ctypes.sizeof(_lib.SparfMLPGrad) == 8 * 28
