"""Solver registry: same plug-in point as reference cmfpy/algs/__init__.py:10-15.
The multiplicative-update solver is the hot path; the two
gradient solvers reuse its contraction kernels (section 8f).  HALS is not provided."""
from .gradient_descent import BlockDescent, GradDescent
from .mult import MultUpdate

ALGORITHMS = {
    "mult": MultUpdate,
    "gd": GradDescent,
    "bcd": BlockDescent,
}
