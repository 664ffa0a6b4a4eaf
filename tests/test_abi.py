"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/gpcsd_b200.h declares,
and the ctypes table in gpcsd_b200/_lib.py binds exactly that set.  No compute calls here."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpcsd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpcsd_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as ge
    lib_path = ge.ensure_built()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in include/gpcsd_b200.h is not exported" % s
    lib.gpcsd_abi_version.restype = ctypes.c_int
    assert lib.gpcsd_abi_version() == 1


def test_ctypes_table_matches_header():
    from gpcsd_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    lib.gpcsd_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.gpcsd_last_error(), bytes)


def test_product_does_not_import_oracle():
    """The product package must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "gpcsd_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpcsd_b200._lib import GpcsdLibraryError
    from gpcsd_b200.gpcsd1d import GPCSD1D
    np.random.seed(0)
    x = np.linspace(0, 2300, 24)[:, None]
    t = np.linspace(0, 50, 20)[:, None]
    m = GPCSD1D(np.random.randn(24, 20, 2), x, t)
    with pytest.raises(GpcsdLibraryError):
        m.loglik()
    with pytest.raises(GpcsdLibraryError):
        m.spatial_cov.compKphi_1d(100.0)
