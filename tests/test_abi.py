"""CPU-only: the C-ABI library loads and exports every symbol include/cvxopt_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "cvxopt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cvxb_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from cvxopt_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `make` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in the header but not exported" % s
    # and the ctypes signature table covers exactly the header
    assert sorted(_lib.exported_symbols()) == syms


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly, never fall back."""
    import numpy as np
    import cvxopt_b200
    if cvxopt_b200.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError):
        cvxopt_b200.kkt_chol(np.zeros((4, 2), order="F"), {"l": 4, "q": [], "s": []})


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cvxopt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "kkt_oracle" not in src and "oracle/_ref" not in src and "import cvxopt\n" not in src, f


def test_dims_validation():
    from cvxopt_b200.kkt import make_dims
    with pytest.raises(TypeError):
        make_dims({"l": -1, "q": [], "s": []})
    with pytest.raises(TypeError):
        make_dims({"l": 1, "q": [0], "s": []})
    d, keep, cdim, cp = make_dims({"l": 2, "q": [3], "s": [2, 3]})
    assert (cdim, cp) == (2 + 3 + 4 + 9, 2 + 3 + 3 + 6)
