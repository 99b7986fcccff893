"""The C-ABI shared library loads on a CPU-only box and exports every symbol
include/chroma_b200.h declares; struct layouts agree with the header."""
import ctypes as C
import os
import re
import subprocess

import pytest

from chroma_lite_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'chroma_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.cb_abi_version() == 4


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(CbMaterial),sizeof(CbSurface),sizeof(CbGeometryDesc),sizeof(CbGeometryInfo),'
                   'sizeof(CbPhotonBank),sizeof(CbPropagateStats));return 0;}\n' % HEADER)
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-o', str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mine = [C.sizeof(t) for t in (_lib.CbMaterial, _lib.CbSurface, _lib.CbGeometryDesc, _lib.CbGeometryInfo,
                                  _lib.CbPhotonBank, _lib.CbPropagateStats)]
    assert sizes == mine


def test_no_gpu_means_loud_failure_not_fallback():
    lib = _lib.load()
    if lib.cb_device_count() > 0:
        pytest.skip('a GPU is visible here')
    with pytest.raises(_lib.ChromaB200Error):
        _lib.init(0)
    # calls before cb_init fail with an error code and a message
    p = C.c_void_p()
    assert lib.cb_malloc(16, C.byref(p)) != 0
    assert b'cb_init' in lib.cb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'chroma_lite_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'liborc' not in text, f
