"""CPU: the C-ABI library loads, exports every symbol include/o2345.h declares, and the ctypes
table covers exactly that set (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "o2345.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(o2345_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    assert len(declared_symbols()) >= 25


def test_library_exports_every_declared_symbol():
    from o2345 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/o2345.h but not exported"
    assert lib.o2345_abi_version() == 1


def test_ctypes_table_matches_header():
    from o2345 import _lib
    assert sorted(_lib.EXPORTED) == declared_symbols()


def test_missing_library_fails_loudly(monkeypatch):
    from o2345 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libo2345_sm100.so")
    with pytest.raises(_lib.O2345Error):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    import torch
    from o2345 import _lib, ops
    with pytest.raises(_lib.O2345Error):
        ops._p(torch.zeros(4))
