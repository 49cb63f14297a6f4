"""Import the UNMODIFIED reference (``/root/reference``) behind two shims.

TEST INFRASTRUCTURE ONLY, and usable only in the build container: the GPU box
has no ``/root/reference``.  Nothing in ``tests -m gpu``, ``smoke()`` or
``bench.py`` calls this at run time; it exists for ``oracle/make_golden.py``
and for the container-only cross-check in ``tests/test_oracle.py``.

Why shims are needed: ``cmfpy/common.py:9`` uses
``np.float`` (removed in NumPy >= 1.24), ``cmfpy/model.py:7`` imports h5py and
``cmfpy/visualize.py:5-6`` imports matplotlib (both absent here).  None of them
touches the MU arithmetic.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CMFPY_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cmfpy"))


def import_reference():
    """Returns the reference ``cmfpy`` package (imported once)."""
    if "cmfpy" in sys.modules and getattr(sys.modules["cmfpy"], "_b200_shimmed", False):
        return sys.modules["cmfpy"]
    if not available():
        raise ImportError("reference checkout not present at %s" % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float                      # shim 1: removed NumPy alias
    for name in ("h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
        if name not in sys.modules:           # shim 2: absent plotting / IO deps
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    gs = sys.modules["matplotlib.gridspec"]
    if not hasattr(gs, "GridSpec"):
        gs.GridSpec = object
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "gridspec"):
        if not hasattr(mpl, sub):
            setattr(mpl, sub, sys.modules["matplotlib." + sub])
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import cmfpy
    finally:
        sys.path.remove(REFERENCE_ROOT)
    cmfpy._b200_shimmed = True
    return cmfpy
