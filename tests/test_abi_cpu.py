"""CPU: the C-ABI library loads and exports every symbol include/gansynth_b200.h declares, with the
argument counts the ctypes binding uses (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "gansynth_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(gs_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from gansynth_b200 import _lib
    ge.build()
    lib = _lib.load()
    decl = _header_functions()
    assert len(decl) >= 37
    for name, nargs in decl.items():
        assert hasattr(lib, name), "missing export %s" % name
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == nargs, "%s: header has %d args, binding %d" % (
                name, nargs, len(_lib.SIGNATURES[name]))
    for name in _lib.SIGNATURES:
        assert name in decl, "binding %s is not declared in the header" % name
    assert lib.gs_version() >= 100
    assert lib.gs_last_error() is not None


def test_product_refuses_to_run_without_cuda_tensors():
    """No CPU fallback: a CPU tensor must raise, not compute."""
    import torch
    from gansynth_b200 import _lib
    from gansynth_b200.kernels import CudaBackend
    with pytest.raises(_lib.GansynthLibraryError):
        CudaBackend().lrelu(torch.zeros(8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gansynth_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
