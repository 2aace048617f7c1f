"""The C-ABI library loads and exports every symbol include/theia_tad.h declares (no GPU needed)."""
import ctypes
import os
import re

from theia_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "theia_tad.h")).read()
    return sorted(set(re.findall(r"^(?:int|void|const char \*)\s*\*?(tad_\w+)\(", src, flags=re.M)))


def test_header_symbols_exported():
    L = _lib.load()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert set(names) == set(_lib.EXPORTS)


def test_abi_version_and_strerror():
    L = _lib.load()
    assert L.tad_abi_version() == 1
    assert L.tad_strerror(0) == b"ok"
    assert L.tad_strerror(-1) == b"invalid argument"
    assert L.tad_strerror(-99) == b"unknown error"


def test_struct_sizes_match_header():
    # sizes the cgo/ctypes bindings rely on (LP64)
    assert ctypes.sizeof(_lib.TadConfig) == 32
    assert ctypes.sizeof(_lib.TadColumns) == 24 + 10 * 8
    assert ctypes.sizeof(_lib.TadJobSpec) == 24 + 8 + 40
    assert ctypes.sizeof(_lib.TadRows) == 8 + 11 * 8
    assert ctypes.sizeof(_lib.TadStatus) == 16 + 256 + 8 * 8 + 2 * 8 + 9 * 8
