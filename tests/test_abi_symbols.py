"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/b200nufft.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'b200nufft.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200nufft_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    from pynufft_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    raw = ctypes.CDLL(_lib.LIBPATH)
    for s in syms:
        assert hasattr(raw, s), 'missing symbol ' + s
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.b200nufft_version() >= 100


def test_argument_errors_without_gpu():
    from pynufft_b200 import _lib
    lib = _lib.load()
    rc = lib.b200nufft_plan_create(None, 0, 2, None, None, None, 0, None, 1, None, None, None, None, None)
    assert rc == -1
    assert b'out is NULL' in lib.b200nufft_last_error()


def test_no_cpu_processor():
    import pytest
    import torch
    import pynufft_b200
    with pytest.raises(RuntimeError):
        pynufft_b200.NUFFT()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pynufft_b200.NUFFT('cuda:0')


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'pynufft_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('no CPU fallback', ''), f
