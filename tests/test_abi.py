"""The C-ABI library loads on a CPU-only box and exports every symbol include/peppan_b200.h
declares; compute entry points are NOT called here (no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'peppan_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from peppan_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, 'declared in include/peppan_b200.h but not exported: %s' % missing


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device pb_init must fail loudly with PB_ERR_NODEVICE."""
    import pytest
    from peppan_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.pb_init(0, 0, 1, None, ctypes.byref(h))
    if rc == 0:
        lib.pb_destroy(h)
        pytest.skip('a GPU is present')
    assert rc == -6
    assert b'no CPU fallback' in lib.pb_last_error(None)
    with pytest.raises(_lib.PbError):
        _lib.Context(0)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'peppan_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, fn)).read()
                for needle in ('import pb_oracle', 'from pb_oracle', 'libpb_oracle', 'oracle/pb_oracle.py', '#include "../../oracle', 'orc_'):
                    assert needle not in txt, '%s references the oracle (%s)' % (fn, needle)


def test_plain_c_client_links_and_runs(tmp_path):
    """tests/c/abi_smoke.c: a C99 program against include/peppan_b200.h and the shared library"""
    import subprocess
    exe = os.path.join(tmp_path, 'abi_smoke')
    libdir = os.path.join(ROOT, 'peppan_b200')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'c', 'abi_smoke.c'),
                           '-L', libdir, '-lpeppan_b200', '-Wl,-rpath,' + libdir, '-o', exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'C ABI OK' in out.stdout, out.stdout + out.stderr
