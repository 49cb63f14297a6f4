"""Host-side solver base: the duck type `CMF.fit` drives.

Mirrors the interface of reference cmfpy/algs/base.py:12-97 (constructor
signature, `update`, `converged`, `loss`, `W`, `H`, `X`, `normX`, `est`,
`resids`, dimension attributes) on top of the C ABI.  All state lives on the
GPU; `W`, `H`, `est`, `resids` materialise fresh float64 NumPy arrays on access.
"""
import ctypes as C
from numbers import Integral

import numpy as np

from .. import _lib


def _as_host_matrix(a, name):
    if not isinstance(a, np.ndarray):
        a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    return a


def resolve_precision(lib, precision, N, K, L):
    """'auto' = the fastest mode that meets the reference's fp32 parity bar (loss trajectory within 1e-4 of the
    float64 reference): the error-compensated tensor-core mode 'tf32x3' where its kernels cover the shape, the
    exact-fp32 FFMA kernels otherwise.  Plain 'tf32' is faster and less accurate; it is never chosen implicitly."""
    if precision != "auto":
        return precision
    return "tf32x3" if lib.cmf_precision_supported(_lib.CMF_PREC_TF32X3, N, K, L) else "fp32"


class DeviceOptimizer:
    """Common plumbing for solvers that live behind libcmf_b200."""

    def __init__(self, data, model_dimensions, initW=None, initH=None,
                 tol=1e-5, patience=3, precision="auto", device=0, seed=None,
                 denominators="auto", normalize=None, loss_precision="auto"):
        # reference base.py:20-21
        if patience < 1 or not isinstance(patience, Integral):
            raise ValueError("Patience must be a positive integer.")
        if precision != "auto" and precision not in _lib.PRECISIONS:
            raise ValueError("precision must be 'auto' or one of %s" % sorted(_lib.PRECISIONS))
        if denominators not in _lib.DENOMINATORS:
            raise ValueError("denominators must be one of %s" % sorted(_lib.DENOMINATORS))
        self.denominators = denominators
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.patience = patience
        self.tol = tol
        self.device = device

        # reference base.py:33-34: copy the model dimensions onto the solver
        for k, v in model_dimensions:
            setattr(self, k, v)
        N, T, K, L = self.n_features, self.n_timepoints, self.n_components, self.maxlag
        self.precision = precision = resolve_precision(self._lib, precision, N, K, L)

        # a NumPy array (the reference's contract) or a row-major float32 matrix already on the device
        # (datasets.DeviceMatrix, a torch / cupy array: anything with __cuda_array_interface__)
        from ..datasets import describe_device_matrix, is_device_matrix
        on_device = is_device_matrix(data)
        if on_device:
            x_ptr, x_ld, x_shape = describe_device_matrix(data)
            x_dt, x_mem = _lib.CMF_F32, _lib.CMF_DEVICE
            self._X, self._X_device = None, data                 # read back only if `.X` / `.resids` are asked for
        else:
            X = _as_host_matrix(data, "data")
            x_ptr, x_ld, x_shape = X.ctypes.data, T, X.shape
            x_dt, x_mem = _lib.np_dtype_code(X), _lib.CMF_HOST
            self._X = X
        if tuple(x_shape) != (N, T):
            raise ValueError("data has shape %s, dimensions say %s" % (tuple(x_shape), (N, T)))

        p = _lib.Params(n_features=N, n_components=K, maxlag=L, t_local=T, t_global=T,
                        t_offset=0, device=device, precision=_lib.PRECISIONS[precision],
                        stream=None, denominators=_lib.DENOMINATORS[denominators])
        _lib.check(self._lib.cmf_mu_create(C.byref(self._h), C.byref(p)))
        if loss_precision not in _lib.LOSS_MODES:
            raise ValueError("loss_precision must be one of %s" % sorted(_lib.LOSS_MODES))
        _lib.check(self._lib.cmf_mu_set_loss_mode(self._h, _lib.LOSS_MODES[loss_precision]))
        _lib.check(self._lib.cmf_mu_set_data(self._h, x_ptr, x_dt, x_mem, x_ld, T))
        if normalize is not None:
            # the reference normalises in its dataset classes, on the host, before the solver sees the data
            # (songbird.py:18-19, maze.py:71-72, vox_celeb.py:100-102); here the rows are scaled where they live
            from ..common import row_scales
            s1, s2, sa = (np.empty(N) for _ in range(3))
            _lib.check(self._lib.cmf_mu_row_stats(self._h, s1.ctypes.data, s2.ctypes.data, sa.ctypes.data))
            self.row_scale = row_scales(normalize, s1, s2, sa, T)
            _lib.check(self._lib.cmf_mu_scale_rows(self._h, self.row_scale.ctypes.data))
            if not on_device:
                self._X = X * self.row_scale[:, None]
        ss, neg = C.c_double(0), C.c_int(0)
        _lib.check(self._lib.cmf_mu_data_stats(self._h, C.byref(ss), C.byref(neg)))
        self.normX = float(np.sqrt(ss.value))               # base.py:25
        self.has_negative = bool(neg.value)

        # reference base.py:37-41
        if initW is None or initH is None:
            initW, initH = self.initialize(seed)
        W0, H0 = _as_host_matrix(initW, "initW"), _as_host_matrix(initH, "initH")
        if W0.shape != (L, N, K) or H0.shape != (K, T):
            raise ValueError("initW/initH must have shapes %s and %s" % ((L, N, K), (K, T)))
        if W0.dtype != H0.dtype:
            H0 = H0.astype(W0.dtype)
        _lib.check(self._lib.cmf_mu_set_factors(self._h, W0.ctypes.data, H0.ctypes.data,
                                                _lib.np_dtype_code(W0), _lib.CMF_HOST, T))
        self.cache_resids()

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cmf_mu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference interface -----------------------------------------------
    def update(self):
        raise NotImplementedError("Base class must override update(...)")

    def initialize(self, seed=None):
        return self.rand_init(seed)

    def rand_init(self, seed=None):
        """reference base.py:78-88: U[0,1) factors rescaled by sqrt(alpha),
        alpha = <X, est> / ||est||^2; the reconstruction and both reductions
        run on the device (the reference draws from the global unseeded numpy
        RNG; here the generator is seedable)."""
        rng = np.random.default_rng(seed)
        N, T, K, L = self.n_features, self.n_timepoints, self.n_components, self.maxlag
        W = rng.random((L, N, K), dtype=np.float32)
        H = rng.random((K, T), dtype=np.float32)
        _lib.check(self._lib.cmf_mu_set_factors(self._h, W.ctypes.data, H.ctypes.data,
                                                _lib.CMF_F32, _lib.CMF_HOST, T))
        xe, ee = C.c_double(0), C.c_double(0)
        _lib.check(self._lib.cmf_mu_init_stats(self._h, C.byref(xe), C.byref(ee)))
        with np.errstate(invalid="ignore"):          # (negative data: CMF.fit raises right after construction)
            s = np.float32(np.sqrt(xe.value / ee.value))
        return s * W, s * H

    def cache_resids(self):
        """reference base.py:57-62: refresh the residual norm (and est, where a solver step reads it: with both MU
        denominators on the Gram route est is only formed when `.est` / `.resids` ask for it)."""
        _lib.check(self._lib.cmf_mu_recon_loss(self._h))

    def converged(self, loss_hist):
        """reference base.py:64-76."""
        d_loss = np.diff(loss_hist[-self.patience:])
        return bool(np.all(np.abs(d_loss) < self.tol))

    @property
    def loss(self):
        """reference base.py:90-97: ||resids||_F / ||X||_F."""
        out = C.c_double(0)
        _lib.check(self._lib.cmf_mu_loss(self._h, C.byref(out)))
        return float(out.value)

    @property
    def X(self):
        if self._X is None:                                  # device data: one read-back, on demand
            d = self._X_device
            X = d.to_host() if hasattr(d, "to_host") else np.asarray(d.cpu(), dtype=np.float64)
            scale = getattr(self, "row_scale", None)
            self._X = X if scale is None else X * scale[:, None]
        return self._X

    @property
    def W(self):
        L, N, K = self.maxlag, self.n_features, self.n_components
        out = np.empty((L, N, K), dtype=np.float64)
        _lib.check(self._lib.cmf_mu_get_W(self._h, out.ctypes.data, _lib.CMF_F64, _lib.CMF_HOST))
        return out

    @property
    def H(self):
        K, T = self.n_components, self.n_timepoints
        out = np.empty((K, T), dtype=np.float64)
        _lib.check(self._lib.cmf_mu_get_H(self._h, out.ctypes.data, _lib.CMF_F64, _lib.CMF_HOST, T))
        return out

    @property
    def est(self):
        N, T = self.n_features, self.n_timepoints
        out = np.empty((N, T), dtype=np.float64)
        _lib.check(self._lib.cmf_mu_get_est(self._h, out.ctypes.data, _lib.CMF_F64, _lib.CMF_HOST, T))
        return out

    @property
    def resids(self):
        return self.est - self.X

    # -- extras --------------------------------------------------------------
    @property
    def path_name(self):
        return self._lib.cmf_mu_path_name(self._h).decode()

    @property
    def launch_count(self):
        n = C.c_longlong(0)
        _lib.check(self._lib.cmf_mu_launch_count(self._h, C.byref(n)))
        return n.value
