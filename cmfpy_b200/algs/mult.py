"""Multiplicative-update solver on the GPU: drop-in for reference
cmfpy/algs/mult.py (`MultUpdate`), same constructor and methods."""
import ctypes as C

import numpy as np

from .. import _lib
from .base import DeviceOptimizer


class MultUpdate(DeviceOptimizer):
    """Multiplicative update rule (reference cmfpy/algs/mult.py:7-48).

    Extra keyword options (all optional, passed through `CMF(**alg_opts)`):
      precision : "auto" (default: "tf32x3" where the tensor-core kernels cover the shape, else "fp32"),
                  "tf32x3" (tcgen05 tensor cores, every product as three TF32 MMAs on hi/lo operand pairs,
                  two-level accumulation: meets the fp32 parity bar), "fp32" (exact FFMA contractions) or
                  "tf32" (plain TF32 tensor cores: ~3x faster, loss trajectories within ~1e-3)
      denominators : "direct" (contract est, as the reference does), "gram" (exact identity
                  through the lag Gram operators of W and H; tensor-core modes only) or "auto"
      loss_precision : how the loss is formed when both denominators are on the Gram route (cmf_mu_set_loss_mode):
                  "auto" (default; tf32x3: from the W terms of the updated factors - no reconstruction at all - while
                  loss^2 >= 0.1, else a one-pass reconstruction on large problems, else the full one; error < 1e-5
                  relative), "full" (the residual formed explicitly every iteration) or "wterms" (always the
                  identity, any precision); W and H are never affected
      device    : CUDA device ordinal
      devices   : list of CUDA ordinals: time-sharded solve over several GPUs (algs/multi_gpu.py)
      seed      : seed of the random initialisation when initW/initH are absent
    """

    def __new__(cls, data=None, dims=None, *args, devices=None, **kwargs):
        # devices=[g0, g1, ...]: the time-sharded solve, one GPU per shard, driven from this process
        if cls is MultUpdate and devices is not None and len(devices) > 1:
            from .multi_gpu import MultiGpuMultUpdate
            return MultiGpuMultUpdate(data, dims, *args, devices=devices, **kwargs)
        return super().__new__(cls)

    def __init__(self, data, dims, patience=3, tol=1e-5, devices=None, **kwargs):
        if devices is not None:
            kwargs.setdefault("device", int(list(devices)[0]))
        super().__init__(data, dims, patience=patience, tol=tol, **kwargs)

    def update(self):
        """One MU iteration (reference mult.py:15-25); returns the loss."""
        return self.update_many(1)[0]

    def update_many(self, n_steps, return_times=False):
        """`n_steps` iterations with a single host synchronisation.  Returns the
        list of losses (what n_steps calls of update() would have returned) and,
        optionally, the per-iteration device times in seconds."""
        losses = np.empty(n_steps, dtype=np.float64)
        ms = np.empty(n_steps, dtype=np.float32) if return_times else None
        _lib.check(self._lib.cmf_mu_step(
            self._h, n_steps, losses.ctypes.data_as(C.POINTER(C.c_double)),
            ms.ctypes.data_as(C.POINTER(C.c_float)) if return_times else None))
        if return_times:
            return [float(x) for x in losses], [float(x) * 1e-3 for x in ms]
        return [float(x) for x in losses]

    # kernel-level access used by the parity tests -------------------------
    def _compute_mult_W(self):
        """reference mult.py:27-40 -> (num, denom), each L x N x K."""
        L, N, K = self.maxlag, self.n_features, self.n_components
        _lib.check(self._lib.cmf_mu_recon(self._h))
        _lib.check(self._lib.cmf_mu_w_terms(self._h))
        num, den = np.empty((L, N, K)), np.empty((L, N, K))
        _lib.check(self._lib.cmf_mu_get_w_terms(self._h, num.ctypes.data, den.ctypes.data, _lib.CMF_F64))
        return num, den

    def _compute_mult_H(self):
        """reference mult.py:42-48 -> (num, denom), each K x T."""
        K, T = self.n_components, self.n_timepoints
        num, den = np.empty((K, T)), np.empty((K, T))
        _lib.check(self._lib.cmf_mu_h_terms(self._h, num.ctypes.data, den.ctypes.data, _lib.CMF_F64))
        return num, den

    def kernel_ms(self):
        out = (C.c_float * 4)()
        _lib.check(self._lib.cmf_mu_kernel_ms(self._h, out))
        return dict(recon=out[0], w_terms=out[1], h_terms=out[2], elementwise=out[3])

    def set_profiling(self, on=True):
        _lib.check(self._lib.cmf_mu_set_profiling(self._h, int(on)))

    def launch_table(self):
        """{kernel label: (launches, total ms)} since set_profiling(2)."""
        return _lib.launch_table(self._lib, self._h)
