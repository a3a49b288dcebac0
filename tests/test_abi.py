"""The C-ABI library loads and exports every symbol include/esmk.h declares; argument
errors are reported through return codes + esmk_last_error (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'esmk.h')).read()
    return sorted(set(re.findall(r'ESMK_API[^;(]*?\b(esmk_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from esme import _lib
    syms = _header_symbols()
    assert len(syms) >= 16
    lib = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/esmk.h but not exported'
    assert sorted(_lib.SIGNATURES) == syms          # the ctypes binding covers exactly the header


def test_version_and_error_reporting():
    from esme import _lib
    assert _lib.lib.esmk_version() >= 100
    assert _lib.lib.esmk_gemm(None, None) != 0
    assert b'null' in _lib.lib.esmk_last_error()
    a = _lib.GemmArgs()
    a.M, a.N, a.K, a.lda = 4, 4, 7, 7                  # K not a multiple of 8
    assert _lib.lib.esmk_gemm(C.byref(a), None) != 0
    assert b'multiples of 8' in _lib.lib.esmk_last_error()
    with pytest.raises(_lib.EsmkError):
        _lib.check(_lib.lib.esmk_model_create(None, None, None), 'esmk_model_create')
    assert _lib.lib.esmk_workspace_bytes(None, 10, 1, 10) == 0


def test_no_cpu_fallback():
    import torch
    from esme import ops
    x = torch.zeros(4, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.layernorm(x, torch.ones(8, dtype=torch.bfloat16), None)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.linear(x, torch.zeros(8, 8, dtype=torch.bfloat16))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'esm-efficient_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                src = open(os.path.join(d, f)).read()
                assert 'oracle' not in src, f'{f} references the oracle'


def test_header_is_plain_c_and_links(tmp_path):
    """include/esmk.h compiles as C11 (no C++-isms in the ABI), the ctypes mirrors have the C struct sizes, and a C
    program that takes the address of every declared entry point links against the shared library."""
    import shutil
    import subprocess
    from esme import _lib
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    syms = _header_symbols()
    src = tmp_path / 'abi_check.c'
    src.write_text(
        '#include <stdio.h>\n#include "esmk.h"\n'
        'int main(void) {\n'
        '  void (*fns[])(void) = {' + ', '.join(f'(void (*)(void))&{s}' for s in syms) + '};\n'
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(fns) / sizeof(fns[0]), sizeof(esmk_gemm_args), sizeof(esmk_config),\n'
        '         sizeof(esmk_layer_weights), sizeof(esmk_weights), sizeof(esmk_qweight));\n'
        '  return esmk_version() >= 100 ? 0 : 1;\n}\n')
    exe = tmp_path / 'abi_check'
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(['gcc', '-std=c11', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'), str(src),
                        '-o', str(exe), '-L', libdir, '-lesmk', f'-Wl,-rpath,{libdir}'], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    n, gemm, cfg, layer, weights, qw = map(int, out.stdout.split())
    assert n == len(syms)
    assert (gemm, cfg, layer, weights, qw) == (C.sizeof(_lib.GemmArgs), C.sizeof(_lib.Config), C.sizeof(_lib.LayerWeights),
                                               C.sizeof(_lib.Weights), C.sizeof(_lib.QWeight))


def test_integration_md_config_struct_matches_header():
    """The ctypes `_Cfg` snippet of INTEGRATION.md B3 must list exactly the fields of `esmk_config` (a maintainer
    copying a short struct would hand the library garbage in the trailing fields)."""
    import re
    header = open(os.path.join(ROOT, 'include', 'esmk.h')).read()
    body = re.search(r'typedef struct \{([^}]*)\} esmk_config;', header, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if decl:
            ctype, names = decl.split(None, 1)
            fields += [(n.strip(), ctype) for n in names.split(',')]
    md = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    snippet = md[md.index('class _Cfg(ctypes.Structure)'):]
    snippet = snippet[:snippet.index('\n# fill')]
    names = re.findall(r"'(\w+)'", snippet)
    assert names == [n for n, _ in fields], (names, fields)
    assert [n for n, t in fields if t == 'float'] == ['residue_scaling'] and "('residue_scaling', ctypes.c_float)" in snippet
    from esme import _lib
    assert [n for n, _ in _lib.Config._fields_] == names
