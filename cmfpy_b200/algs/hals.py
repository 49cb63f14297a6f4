"""HALS on the GPU: drop-in for reference cmfpy/algs/hals.py (`HALSUpdate`) and
the accelerated inner-iteration scheme it inherits (cmfpy/algs/accelerated.py).

One coordinate block at a time, the residual `est - X` kept current on the device
after every block (csrc/hals_kernels.cuh).  The sweeps are sequential in
(component, lag) by definition, so this solver is HBM-bound - every block update
is a pass over the N x T residual; it serves problems up to BASELINE config B
comfortably and is not meant for config C.
"""
import ctypes as C
import time

from .. import _lib
from .base import DeviceOptimizer


class HALSUpdate(DeviceOptimizer):
    """Hierarchical alternating least squares (reference hals.py:11-70)."""

    batchable = False

    def __init__(self, data, dims, patience=3, tol=1e-5, max_iter=1,
                 weightW=1, weightH=1, stop_thresh=0, **kwargs):
        if kwargs.get("precision", "auto") == "auto":
            kwargs["precision"] = "fp32"        # the sweeps are fp32 passes over the residual; nothing for tensor cores
        super().__init__(data, dims, patience=patience, tol=tol, **kwargs)
        self.max_iter = max_iter
        self.stop_thresh = stop_thresh
        self.weightW = weightW
        self.weightH = weightH
        if max_iter * min(1, weightH, weightW) < 1:                   # accelerated.py:47-48
            raise ValueError("Requires at least 1 iteration for both W and H.")
        if self.n_timepoints <= self.maxlag:
            raise ValueError("HALS needs more time points than lags")

    def _accelerated_update(self, sweep, weight):
        """accelerated.py:50-69: one sweep, then more while the budget lasts and the factor still moves."""
        more = self.max_iter * weight > 1
        d = C.c_double(0)
        _lib.check(sweep(self._h, C.byref(d) if more else None))
        init_diff = diff = d.value
        itr = 1
        while itr < self.max_iter * weight and diff > self.stop_thresh * init_diff:
            itr += 1
            _lib.check(sweep(self._h, C.byref(d)))
            diff = d.value

    def update(self):
        """accelerated.py:71-84: W sweeps, H sweeps, residual from scratch, loss."""
        lib = self._lib
        _lib.check(lib.cmf_hals_begin(self._h))
        self._accelerated_update(lib.cmf_hals_sweep_w, self.weightW)
        self._accelerated_update(lib.cmf_hals_sweep_h, self.weightH)
        loss = C.c_double(0)
        _lib.check(lib.cmf_hals_end(self._h, C.byref(loss)))
        return float(loss.value)

    def update_many(self, n_steps, return_times=False):
        losses, secs = [], []
        for _ in range(n_steps):
            t0 = time.perf_counter()
            losses.append(self.update())
            secs.append(time.perf_counter() - t0)
        return (losses, secs) if return_times else losses
