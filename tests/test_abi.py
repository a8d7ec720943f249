"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/celda_cuda.h declares (no compute calls here: there is no GPU in the build container)."""
import ctypes
import os

import celda_b200
from celda_b200 import _lib


def test_library_exports_every_declared_symbol():
    so = _lib.build()
    assert os.path.exists(so)
    L = ctypes.CDLL(so)
    decl = _lib.declared_functions()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/celda_cuda.h but not exported"
    for must in ("celda_colSumByGroupSparse", "celda_rowSumByGroupSparse", "celda_colSumByGroupChangeSparse",
                 "celda_rowSumByGroupChangeSparse", "celda_chain_create", "celda_chain_iterate", "celda_chain_loglik"):
        assert must in decl


def test_signatures_have_no_foreign_types():
    """Plain pointers and sizes only (no torch / numpy / SEXP types in the ABI)."""
    for name, (ret, args) in _lib.declared_functions().items():
        for t in [ret] + args:
            assert t in _lib._TYPES, (name, t)


def test_no_cpu_fallback(monkeypatch):
    """The product fails loudly when the CUDA library is missing."""
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libcelda_cuda.so")
    monkeypatch.setattr(_lib, "_lib", None)
    try:
        celda_b200.launch_count()
    except celda_b200.CeldaCudaError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected CeldaCudaError")


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "celda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("no cpu fallback", ""), f
