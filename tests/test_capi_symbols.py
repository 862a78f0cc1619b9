"""CPU: the C-ABI library loads and exports every symbol include/stanmath_cuda.h
declares, the ctypes table mirrors the header, and -- with no GPU -- compute
calls fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stanmath_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from math_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"


def test_ctypes_table_matches_header():
    from math_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.lib()  # binds every signature; raises on a missing symbol


def test_no_cpu_fallback():
    import math_b200 as mb
    try:
        n = mb.runtime.device_count()
    except mb.BackendError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present: nothing to check here")
    with pytest.raises(mb.BackendError):
        mb.MatrixCuda.from_host(np.zeros((3, 2)))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "math_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f), errors="replace").read()
            assert "pyoracle" not in text and "glm_oracle" not in text


def test_synthetic_host_statistics():
    import math_b200 as mb
    x = mb.synthetic_host(12345, 0, 200000, 3)
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert np.array_equal(x[1000:1010], mb.synthetic_host(12345, 1000, 10, 3))
    y = mb.synthetic_host(7, 0, 10000, 1, kind=1, lo=0, hi=4)
    assert y.min() == 0 and y.max() == 4 and y.dtype == np.int32
