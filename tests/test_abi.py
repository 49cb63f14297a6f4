"""CPU-side checks of the C-ABI library: it builds, loads, exports every
symbol include/cmf_b200.h declares, and fails loudly without a GPU (no compute
calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cmf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmf_[a-z_A-Z0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from cmfpy_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(built_lib, name), "libcmf_b200.so lacks %s" % name
    # and the Python binding covers the header exactly
    assert sorted(_lib.exported_symbols()) == declared


def test_abi_version_and_error_channel(built_lib):
    assert built_lib.cmf_abi_version() == 1
    rc = built_lib.cmf_mu_create(None, None)
    assert rc != 0
    assert b"null" in built_lib.cmf_last_error()


def test_no_cpu_fallback(built_lib):
    """Without a visible sm_100 GPU the solver must refuse to construct."""
    from cmfpy_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    from cmfpy_b200 import CMF
    X = np.random.default_rng(0).random((8, 64))
    with pytest.raises((RuntimeError, ValueError)):
        CMF(2, 4, verbose=False).fit(X)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; nothing under cmfpy_b200/ may use it."""
    pkg = os.path.join(ROOT, "cmfpy_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), "%s references the oracle" % f


def test_host_side_argument_errors():
    from cmfpy_b200 import CMF, ModelDimensions
    from cmfpy_b200.model import NOT_FITTED_ERROR
    with pytest.raises(ValueError):
        ModelDimensions(maxlag=3, n_components=2)
    with pytest.raises(ValueError):
        ModelDimensions(np.zeros((2, 5)), n_components=2)
    with pytest.raises(ValueError):
        ModelDimensions(np.zeros((2, 5)), maxlag=2)
    d = dict(ModelDimensions(np.zeros((2, 5)), maxlag=2, n_components=3))
    assert d == dict(n_features=2, n_timepoints=5, maxlag=2, n_components=3)
    m = CMF(2, 3, verbose=False)
    with pytest.raises(ValueError):
        m.motifs
    with pytest.raises(ValueError):
        m.factors
    with pytest.raises(ValueError):
        m.fit(-np.ones((4, 20)))
    assert isinstance(NOT_FITTED_ERROR, ValueError)
