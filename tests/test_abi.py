"""The product library must export every symbol include/peps_b200.h declares (no compute without a GPU) and must
refuse to run without a CUDA device instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

from peps_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "peps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(peps_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES.keys())


def test_product_library_exports_every_symbol():
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.peps_backend_name.restype = ctypes.c_char_p
    assert lib.peps_backend_name() == b"cuda-sm_100a"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.PepsConfig(2, 2, 2, 2, 1, 0, 2, 2, 0.0)
    h = ctypes.c_void_p()
    rc = lib.peps_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc != 0
    assert b"no CUDA device" in lib.peps_last_error(None)
