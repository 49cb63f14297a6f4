"""cmfpy_b200 - B200-native multiplicative-update engine behind the cmfpy API.

    from cmfpy_b200 import CMF
    model = CMF(n_components=3, maxlag=20, verbose=False, tol=0)
    model.fit(X)            # X: N x T, non-negative
    model.motifs, model.factors, model.loss_hist
"""
from .model import CMF, ModelDimensions
from .algs import ALGORITHMS
from ._lib import release_cached_memory

__version__ = "0.1.0"
__all__ = ["CMF", "ModelDimensions", "ALGORITHMS", "release_cached_memory"]
