"""CPU: the C-ABI library loads and exports every symbol include/mdh_b200.h declares."""
import ctypes
import pathlib
import re

import pytest

from mdhelper_b200 import _lib

ROOT = pathlib.Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "mdh_b200.h").read_text()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(mdh_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    assert "mdh_rdf_accumulate" in syms and "mdh_sq_accumulate" in syms
    assert len(syms) >= 18


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in the header but not exported"


def test_binding_table_covers_the_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_torch_types_in_signatures():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    assert "torch" not in code and "at::" not in code and "std::" not in code


def test_abi_version_and_error_string():
    L = _lib.lib()
    assert L.mdh_abi_version() == 1
    assert isinstance(L.mdh_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu():
    """NULL / bad arguments are reported through the return code, not a crash."""
    L = _lib.lib()
    assert L.mdh_ctx_create(0, None, None) == _lib.MDH_EINVAL
    assert b"out is NULL" in L.mdh_last_error()
    n = ctypes.c_int64()
    assert L.mdh_launch_count(None, ctypes.byref(n)) == _lib.MDH_EINVAL
    assert L.mdh_ctx_destroy(None) == _lib.MDH_OK


def test_product_path_has_no_cpu_fallback():
    """Without a CUDA device the analysis classes must fail loudly."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mdhelper_b200 import synthetic
    from mdhelper_b200.analysis.structure import RadialDistributionFunction
    u = synthetic.lj_fluid(64, 1, seed=1)
    with pytest.raises(RuntimeError):
        RadialDistributionFunction(u.atoms, n_bins=8, range=(0.0, 2.0)).run()
    # and nothing under mdhelper_b200/ may import the oracle
    pkg = ROOT / "mdhelper_b200"
    for f in pkg.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), flags=re.M), f
