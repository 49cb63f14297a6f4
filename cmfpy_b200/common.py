"""GPU-backed versions of the numeric routines of reference cmfpy/common.py
that the MU path uses.  Inputs/outputs are NumPy arrays in the reference
layouts; the work runs on the GPU through the C ABI."""
import numpy as np

from . import _lib

EPSILON = float(np.finfo(np.float64).eps)       # reference common.py:9


def _prep(a):
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def cmf_predict(W, H, precision="fp32", device=0):
    """est = sum_l W[l] @ shift(H, l); reference common.py:50-58."""
    W, H = _prep(W), _prep(H)
    if H.dtype != W.dtype:
        H = H.astype(W.dtype)
    L, N, K = W.shape
    K2, T = H.shape
    if K2 != K:
        raise ValueError("W and H disagree on the number of components")
    out = np.empty((N, T), dtype=W.dtype)
    _lib.check(_lib.load().cmf_predict(W.ctypes.data, H.ctypes.data, out.ctypes.data,
                                       _lib.np_dtype_code(W), N, T, K, L, device,
                                       _lib.PRECISIONS[precision]))
    return out


def tensor_transconv(W, X, precision="fp32", device=0):
    """out[:, t] = sum_l W[l].T @ X[:, t+l]; reference common.py:61-86."""
    W, X = _prep(W), _prep(X)
    if X.dtype != W.dtype:
        X = X.astype(W.dtype)
    L, N, K = W.shape
    N2, T = X.shape
    if N2 != N:
        raise ValueError("W and X disagree on the number of features")
    out = np.empty((K, T), dtype=W.dtype)
    _lib.check(_lib.load().cmf_tensor_transconv(W.ctypes.data, X.ctypes.data, out.ctypes.data,
                                                _lib.np_dtype_code(W), N, T, K, L, device,
                                                _lib.PRECISIONS[precision]))
    return out


def score(W, H, X, precision="fp32", device=0):
    """R^2 = 1 - ||cmf_predict(W, H) - X||^2 / ||X||^2 (reference model.py:202-221), reduced on the
    device by the fused reconstruction + residual kernel: no N x T array comes back."""
    import ctypes as C
    W, H, X = _prep(W), _prep(H), _prep(X)
    H, X = H.astype(W.dtype, copy=False), X.astype(W.dtype, copy=False)
    L, N, K = W.shape
    if H.shape[0] != K or X.shape != (N, H.shape[1]):
        raise ValueError("W, H and X disagree on their dimensions")
    r2 = C.c_double(0.0)
    _lib.check(_lib.load().cmf_score(W.ctypes.data, H.ctypes.data, X.ctypes.data, _lib.np_dtype_code(W),
                                     N, H.shape[1], K, L, device, _lib.PRECISIONS[precision], C.byref(r2)))
    return r2.value


NORMALIZATIONS = ("l2", "l1", "std")


def row_scales(mode, s1, s2, sabs, n_timepoints):
    """Per-feature scale factors of the reference's dataset normalisations from the per-feature sums over ALL
    time points (s1 = sum x, s2 = sum x^2, sabs = sum |x|):
      "l2"  : 1 / (1e-6 + ||row||_2)            datasets/songbird.py:18-19
      "l1"  : 1 / (1e-8 + ||row||_1)            datasets/maze.py:71-72
      "std" : 1 / std(row), population variance, 1 where the variance is 0
              (StandardScaler(with_mean=False), datasets/vox_celeb.py:100-102)"""
    s1, s2, sabs = (np.asarray(a, dtype=np.float64) for a in (s1, s2, sabs))
    if mode == "l2":
        return 1.0 / (1e-6 + np.sqrt(s2))
    if mode == "l1":
        return 1.0 / (1e-8 + sabs)
    if mode == "std":
        var = np.maximum(s2 / n_timepoints - (s1 / n_timepoints) ** 2, 0.0)
        std = np.sqrt(var)
        std[std < 10 * np.finfo(np.float64).eps] = 1.0
        return 1.0 / std
    raise ValueError("normalize must be one of %s" % (NORMALIZATIONS,))


def shift_cols(X, lag):
    """reference common.py:89-98 (a view; no device work)."""
    T = X.shape[1]
    return X[:, :T - lag] if lag > 0 else X[:, -lag:]


def s_dot(A, B, shift, precision="fp32", device=0):
    """A @ shift(B, shift); reference common.py:13-29.  A single-lag
    reconstruction: W has one non-zero lag slice."""
    A, B = _prep(A), _prep(B)
    T = B.shape[1]
    if shift >= 0:
        W = np.zeros((shift + 1,) + A.shape, dtype=A.dtype)
        W[shift] = A
        return cmf_predict(W, B, precision, device)
    # left shift: reverse time, shift right, reverse back
    W = np.zeros((-shift + 1,) + A.shape, dtype=A.dtype)
    W[-shift] = A
    return cmf_predict(W, B[:, ::-1], precision, device)[:, ::-1]


def s_T_dot(A, B, shift, precision="fp32", device=0):
    """A[:, s:] @ B[:, :T-s].T (mirror for s<0); reference common.py:32-47.
    With A = X, B = H, s = l this is exactly the lag-l W numerator
    (reference mult.py:37), so it is read off the W-terms kernel."""
    from .algs.mult import MultUpdate
    from .model import ModelDimensions
    A, B = _prep(A), _prep(B)
    if shift < 0:
        return s_T_dot(A[:, ::-1], B[:, ::-1], -shift, precision, device)
    N, K = A.shape[0], B.shape[0]
    dims = ModelDimensions(A, maxlag=shift + 1, n_components=K)
    alg = MultUpdate(A, dims, initW=np.zeros((shift + 1, N, K), dtype=A.dtype),
                     initH=B.astype(A.dtype), precision=precision, device=device)
    num, _ = alg._compute_mult_W()
    alg.close()
    return num[shift].astype(A.dtype)
