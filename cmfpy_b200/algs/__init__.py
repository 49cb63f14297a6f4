"""Solver registry: same plug-in point and the same four names as reference cmfpy/algs/__init__.py:10-15.
The multiplicative-update solver is the hot path; the gradient solvers reuse its contraction
kernels and HALS keeps its residual on the device (section 8f)."""
from .gradient_descent import BlockDescent, GradDescent
from .hals import HALSUpdate
from .mult import MultUpdate

ALGORITHMS = {
    "gd": GradDescent,
    "bcd": BlockDescent,
    "mult": MultUpdate,
    "hals": HALSUpdate,
}
