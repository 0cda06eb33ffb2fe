"""The three descriptions of the C ABI agree: include/drb.h (the contract), drecpy_b200/_lib.py (the binding that
runs) and the ctypes stub printed in INTEGRATION.md (the binding a DRecPy maintainer would paste).  A struct that is
one field short on the Python side hands the library a truncated argument block -- this is the test for that."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import abi  # noqa: E402

from drecpy_b200 import _lib  # noqa: E402

STRUCTS = abi.parse_structs()


def _layout(cls):
    return [(n, getattr(cls, n).offset, getattr(cls, n).size) for n, _ in cls._fields_]


def test_header_declares_the_structs_the_binding_uses():
    names = {abi.py_name(c) for c, _ in STRUCTS}
    assert {'CdaeLayout', 'CdaeDesc', 'CdaeStepArgs', 'DmfLayout', 'DmfDesc', 'DmfStepArgs'} <= names
    for n in names:
        assert hasattr(_lib, n), f'{n} is declared in drb.h but drecpy_b200/_lib.py has no such Structure'


@pytest.mark.parametrize('c_name,fields', STRUCTS, ids=[c for c, _ in STRUCTS])
def test_lib_structs_match_the_header(c_name, fields):
    want = abi.ctypes_struct(fields)
    got = getattr(_lib, abi.py_name(c_name))
    assert [n for n, _ in got._fields_] == [n for n, _, _ in fields], c_name
    assert C.sizeof(got) == C.sizeof(want), c_name
    assert _layout(got) == _layout(want), c_name


@pytest.mark.parametrize('c_name,fields', STRUCTS, ids=[c for c, _ in STRUCTS])
def test_integration_md_structs_match_the_header(c_name, fields):
    _, _, _, code = abi.integration_block()
    assert code.strip(), 'INTEGRATION.md has no generated ABI block (python tools/abi.py --update-integration)'
    ns = {}
    exec(code, ns)
    got = ns[abi.py_name(c_name)]
    want = abi.ctypes_struct(fields)
    assert [n for n, _ in got._fields_] == [n for n, _, _ in fields], c_name
    assert C.sizeof(got) == C.sizeof(want) and _layout(got) == _layout(want), c_name
    assert code == abi.emit(), 'INTEGRATION.md is stale: python tools/abi.py --update-integration'


def test_sizes_and_offsets_against_the_c_compiler(tmp_path):
    """sizeof / offsetof as gcc sees include/drb.h == what ctypes computes for _lib.py's structures."""
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "drb.h"', 'int main(void) {']
    for c_name, fields in STRUCTS:
        lines.append(f'  printf("{c_name} %zu\\n", sizeof({c_name}));')
        for n, _, _ in fields:
            lines.append(f'  printf("{c_name}.{n} %zu\\n", offsetof({c_name}, {n}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'abi.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'abi'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for c_name, fields in STRUCTS:
        cls = getattr(_lib, abi.py_name(c_name))
        assert int(out[c_name]) == C.sizeof(cls), c_name
        for n, _, _ in fields:
            assert int(out[f'{c_name}.{n}']) == getattr(cls, n).offset, f'{c_name}.{n}'


def test_every_declared_function_is_bound_and_exported():
    declared = set(abi.parse_functions())
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f'{name} is declared in drb.h but libdrb.so does not export it'
