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
    declared = set(_declared()) - set(_lib.OTHER_SYMBOLS)
    bound = set(_lib.SIGNATURES) | set(_lib.HOST_SCALAR_SIGNATURES)
    assert declared == bound, declared ^ bound
    assert _lib.lib().abi_version() == _lib.ABI_VERSION


def _struct_fields(name):
    src = open(os.path.join(ROOT, 'include', 'bnpc_b200.h')).read()
    body = re.search(r'typedef struct \{((?:(?!typedef struct).)*?)\} ' + name + ';', src, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    return re.findall(r'(?:const\s+)?[a-z0-9_]+\s*\*?\s*([A-Za-z0-9_]+)\s*(?:\[\d+\])?\s*;', body)


@pytest.mark.parametrize('c_name,mirror', [('bnpc_sweep_args_t', 'SweepArgs'), ('bnpc_chain_t', 'ChainWs'),
                                           ('bnpc_epoch_t', 'Epoch'), ('bnpc_rg_t', 'RgMove'),
                                           ('bnpc_chain_state_t', 'ChainState'), ('bnpc_moves_t', 'Moves'),
                                           ('bnpc_trace_t', 'Trace')])
def test_struct_layouts_match_c(c_name, mirror):
    # field order of the ctypes mirrors of the structs declared in the header
    cls = getattr(_lib, mirror)
    assert _struct_fields(c_name) == [f[0] for f in cls._fields_]
    assert ctypes.sizeof(cls) % 8 == 0


def test_struct_sizes_match_c(tmp_path):
    # compile a tiny C program against the header and compare sizeof() with the mirrors
    import subprocess
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include "bnpc_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(bnpc_sweep_args_t),sizeof(bnpc_chain_t),sizeof(bnpc_epoch_t),sizeof(bnpc_rg_t),'
                   'sizeof(bnpc_visit_t),sizeof(bnpc_cand_t),sizeof(bnpc_chain_state_t),sizeof(bnpc_moves_t),'
                   'sizeof(bnpc_trace_t));return 0;}\n')
    exe = tmp_path / 'sizes'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.SweepArgs), ctypes.sizeof(_lib.ChainWs), ctypes.sizeof(_lib.Epoch),
            ctypes.sizeof(_lib.RgMove), _lib.VISIT_BYTES, _lib.CAND_BYTES, ctypes.sizeof(_lib.ChainState),
            ctypes.sizeof(_lib.Moves), ctypes.sizeof(_lib.Trace)]
    assert got == want


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'SO_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.lib()
