"""The C-ABI shared library loads without a GPU and exports every symbol the header
declares; compute entry points fail loudly (no CPU fallback).  CPU only."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "exauq_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(exq_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    from exauq_toolbox_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/exauq_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_no_cpu_fallback_without_gpu():
    import numpy as np

    from exauq_toolbox_b200 import _lib

    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(_lib.ExqError):
        _lib.corr("Matern52", np.zeros(2), np.zeros((3, 2)), np.zeros((3, 2)))
    lib = _lib.load()
    out = ctypes.c_void_p()
    x = np.zeros((4, 2))
    rc = lib.exq_gp_create(4, 2, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                           x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 0, ctypes.byref(out))
    assert rc == _lib.EXQ_ERR_CUDA


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "exauq-toolbox_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src and "mogp_emulator" not in src.replace(
                "mogp-emulator", "").replace("``mogp_emulator``", "") or fn == "refapi.py", fn
