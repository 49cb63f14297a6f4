"""Solver registry: same plug-in point as reference cmfpy/algs/__init__.py:10-15.
Only the multiplicative-update solver is in scope."""
from .mult import MultUpdate

ALGORITHMS = {
    "mult": MultUpdate,
}
