"""Projected gradient descent and block coordinate descent on the GPU: drop-ins
for reference cmfpy/algs/gradient_descent.py (`GradDescent`, `BlockDescent`),
same constructor, methods and step-size adaptation.

The gradients are the multiplicative-update terms (gW = den_W - num_W,
gH = den_H - num_H), so the solver reuses the MU contraction kernels; the
Lipschitz constant of the W step (`lipschitz_W`, gradient_descent.py:54-69) is
found by a power iteration that stays on the device.  The reference needs a
SciPy that still accepts `eigh(eigvals=...)`; this one has no such dependency.
"""
import ctypes as C
import time
import warnings

import numpy as np

from .. import _lib
from .base import DeviceOptimizer


class GradDescent(DeviceOptimizer):
    """Gradient descent update rules (reference gradient_descent.py:15-123)."""

    block_descent = False
    batchable = False          # converged() adapts the step size: CMF.fit must call it every iteration

    def __init__(self, data, dims, tol=1e-5, patience=3, step_decrement=5., **kwargs):
        kwargs["denominators"] = "direct"           # the gradients contract the residual itself
        super().__init__(data, dims, patience=patience, tol=tol, **kwargs)
        self.step_size = 1e-4                       # gradient_descent.py:29
        self.step_decrement = step_decrement
        self._warned_unsettled = False
        _lib.check(self._lib.cmf_gd_cache(self._h))  # cache_gW, cache_gH (:37-38)

    # -- gradients (gradient_descent.py:40-52), read back on demand ------------
    @property
    def gW(self):
        L, N, K = self.maxlag, self.n_features, self.n_components
        num, den = np.empty((L, N, K)), np.empty((L, N, K))
        _lib.check(self._lib.cmf_mu_get_w_terms(self._h, num.ctypes.data, den.ctypes.data, _lib.CMF_F64))
        return den - num

    @property
    def gH(self):
        K, T = self.n_components, self.n_timepoints
        num, den = np.empty((K, T)), np.empty((K, T))
        _lib.check(self._lib.cmf_mu_h_terms(self._h, num.ctypes.data, den.ctypes.data, _lib.CMF_F64))
        return den - num

    def cache_gW(self):
        return self.gW

    def cache_gH(self):
        return self.gH

    def lipschitz_W(self):
        """Largest eigenvalue of the block-Toeplitz autocorrelation matrix of H (:54-69)."""
        lam = C.c_double(0)
        _lib.check(self._lib.cmf_gd_lipschitz_w(self._h, C.byref(lam)))
        return float(lam.value)

    def lipschitz_state(self):
        """(settled, iterations) of the last power iteration: the reference's `eigh` is exact, a Rayleigh quotient that
        has not settled is a lower bound of lambda_max (a W step that is too long)."""
        ok, it = C.c_int(0), C.c_int(0)
        _lib.check(self._lib.cmf_gd_lipschitz_state(self._h, C.byref(ok), C.byref(it)))
        return bool(ok.value), int(it.value)

    def lipschitz_H(self):
        raise NotImplementedError()                 # as in the reference (:71-79)

    def update(self):
        """One update (:81-92; BlockDescent: :132-147); returns the loss."""
        loss = C.c_double(0)
        _lib.check(self._lib.cmf_gd_step(self._h, int(self.block_descent), float(self.step_size), C.byref(loss)))
        if not self._warned_unsettled:
            settled, iters = self.lipschitz_state()
            if not settled:
                self._warned_unsettled = True
                warnings.warn("lipschitz_W: the power iteration had not settled after %d iterations; the W step "
                              "may be longer than the reference's 1 / lambda_max" % iters, RuntimeWarning)
        return float(loss.value)

    def update_many(self, n_steps, return_times=False):
        losses, secs = [], []
        for _ in range(n_steps):
            t0 = time.perf_counter()
            losses.append(self.update())            # synchronous: the loss comes back to the host
            secs.append(time.perf_counter() - t0)
        return (losses, secs) if return_times else losses

    def converged(self, loss_hist):
        """Convergence test that also shrinks the H step when the loss went up (:94-113)."""
        d_loss = np.diff(loss_hist[-self.patience:])
        if d_loss[-1] > 0:
            self.step_size /= self.step_decrement
            return False
        return bool(np.all(np.abs(d_loss) < self.tol))

    @property
    def unnormalized_loss(self):
        """0.5 ||resids||^2 (:115-118)."""
        return 0.5 * (self.loss * self.normX) ** 2


class BlockDescent(GradDescent):
    """Block coordinate descent (reference gradient_descent.py:126-147): same steps, Gauss-Seidel order."""

    block_descent = True
