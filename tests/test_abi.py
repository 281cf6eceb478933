"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/bnpc_b200.h declares; the ctypes signatures cover exactly those symbols.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from bnpc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'bnpc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bnpc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_header_symbols():
    _lib.build()
    dll = ctypes.CDLL(_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(dll, n), f'{n} declared in include/bnpc_b200.h but not exported'


def test_ctypes_signatures_match_header():
    declared = set(_declared()) - {'bnpc_abi_version', 'bnpc_last_error'}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().abi_version() == 2


def test_sweep_args_layout_matches_c():
    # field order/size of the ctypes mirror of bnpc_sweep_args_t
    src = open(os.path.join(ROOT, 'include', 'bnpc_b200.h')).read()
    body = re.search(r'typedef struct \{((?:(?!typedef struct).)*?)\} bnpc_sweep_args_t;', src, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = re.findall(r'(?:const\s+)?[a-z0-9_]+\s*\*?\s*([A-Za-z0-9_]+)\s*;', body)
    assert fields == [f for f, _ in _lib.SweepArgs._fields_]
    assert ctypes.sizeof(_lib.SweepArgs) % 8 == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'SO_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.lib()
