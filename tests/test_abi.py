"""The C-ABI library loads and exports every symbol include/physs_b200.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "physs_b200.h")


HEADER_BIG = os.path.join(ROOT, "include", "physs_b200_big.h")


def _declared_symbols(header=HEADER):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(physs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    assert "physs_kf_filter_f64" in syms and "physs_rts_smooth_f64" in syms


def test_library_exports_every_declared_symbol():
    from physs_gp_b200 import _lib
    from physs_gp_b200.build import build_library
    build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), "libphyss_b200.so does not export %s" % s


def test_big_block_library_exports_every_declared_symbol():
    """libphyss_b200_big.so (cuBLAS / cuSOLVER-backed large-block path) against include/physs_b200_big.h."""
    from physs_gp_b200 import _lib
    from physs_gp_b200.build import build_big
    build_big()
    lib = ctypes.CDLL(_lib.BIG_LIB_PATH)
    syms = _declared_symbols(HEADER_BIG)
    assert "physs_kf_filter_big_f64" in syms and "physs_rts_smooth_big_f64" in syms
    for s in syms:
        assert hasattr(lib, s), "libphyss_b200_big.so does not export %s" % s
    assert sorted(_lib.BIG_SIGNATURES) == syms
    # bad sizes are rejected before any CUDA / library call
    big = _lib.load_big()
    st = big.physs_kf_filter_big_f64(None, 0, 4, 2, None, None, None, None, None, None, None, None, 0, 1e-5, None, 0,
                                     None, None, None)
    assert st == 1 and b"bad sizes" in big.physs_big_last_error()


def test_binding_table_covers_header():
    from physs_gp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.physs_abi_version() == _lib.ABI_VERSION


def test_bad_arguments_return_status_not_crash():
    from physs_gp_b200 import _lib
    lib = _lib.load()
    # T < 1 is rejected before any CUDA call
    st = lib.physs_kf_filter_f64(None, 1, 0, 0, 0, 2, 1, 0, 0, None, 0, None, 0, None, 0, None, 0, None, 0,
                                 None, 0, None, 0, None, 0, None, None, 0, 0, 1e-5, None, None, None, None)
    assert st == 1
    assert b"bad sizes" in lib.physs_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from physs_gp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libphyss_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "physs_gp_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "oracle/" not in txt or fn.endswith(".py") is False or "oracle/" not in txt.replace("# oracle/", ""), fn


def test_no_cuda_device_raises_not_falls_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from physs_gp_b200 import data, filters, kernels, sdes
    d = data.TemporalData(np.arange(5.0), np.zeros([5, 1, 1]))
    prior = sdes.LTI_SDE(sdes.Independent([kernels.Matern32(1.0)]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        filters.filter_loop(d, prior, R=np.eye(1)[None])
