"""The C-ABI library loads and exports every symbol include/simudo_b200.h
declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'simudo_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported_by_cuda_library():
    from simudo_b200.csrc import build
    so = build.build()                      # nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(so)
    syms = _declared_symbols()
    assert 'sb_assemble' in syms and 'sb_plan_create' in syms
    for s in syms:
        assert hasattr(lib, s), 'missing export ' + s
    from simudo_b200 import _lib
    assert sorted(_lib.EXPORTED) == syms


def test_tables_blob_layout_matches_c_struct():
    from simudo_b200.csrc import build
    from simudo_b200.fem import model as M
    from simudo_b200.fem.tables import pack_tables, SbTables
    lib = ctypes.CDLL(build.build())
    lib.sb_tables_size.restype = ctypes.c_int64
    assert lib.sb_tables_size() == ctypes.sizeof(SbTables)
    t = pack_tables(M.make_model([dict(kind=0, sign=-1)], []))
    assert t.rho.m == 11 and t.sup.m == 11 and t.ng == 25


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from simudo_b200 import _lib
    _lib._reset_for_tests()
    monkeypatch.setattr(_lib, 'SO_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.get()
    _lib._reset_for_tests()


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, 'simudo_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M), f
                assert 'libpdd_oracle' not in txt, f
