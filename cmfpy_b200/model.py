"""CMF model object with the reference's API (reference cmfpy/model.py:24-245),
driving the GPU multiplicative-update solver.

Kept from the reference: constructor arguments, `fit`, `predict`, `score`,
`motifs` (W: L x N x K), `factors` (H: K x T), `n_features`, `n_timesteps`,
`loss_hist`, `time_hist`, `argsort_units`, the error types; working versions of
`sort_components`, `compute_loadings`, `renormalize`; `load_cmfjl_model` (needs
h5py, as in the reference) and an .npz checkpoint with the same dataset names
(`save_model` / `load_model`) for warm restarts through `initW` / `initH`.
Out of scope: plotting.
"""
import time

import numpy as np

from .algs import ALGORITHMS
from .common import cmf_predict

NOT_FITTED_ERROR = ValueError(
    "This CMF instance is not fitted yet. Call 'fit' with appropriate"
    "arguments before using this method."
)


class ModelDimensions:
    """Holds dimensions of a CMF model (reference model.py:24-69)."""

    def __init__(self, data=None, n_features=None, n_timepoints=None,
                 maxlag=None, n_components=None):
        if data is None:
            if (n_features is None) or (n_timepoints is None):
                raise ValueError("Must either specify 'data' or ('n_features' "
                                 "and 'n_timepoints').")
        else:
            n_features, n_timepoints = data.shape
        if maxlag is None:
            raise ValueError("Must specify 'n_lags'.")
        if n_components is None:
            raise ValueError("Must specify 'n_components'.")
        self.n_features = n_features
        self.n_timepoints = n_timepoints
        self.maxlag = maxlag
        self.n_components = n_components

    def __iter__(self):
        yield from self.__dict__.items()


class CMF(object):
    """Convolutive matrix factorization: X ~ sum_l W[l] @ shift(H, l)."""

    def __init__(self, n_components, maxlag, n_iter_max=100,
                 l1_W=0.0, l1_H=0.0, verbose=True, alg_name='mult',
                 **alg_opts):
        """Same parameters as reference model.py:80-120.  `l1_W`/`l1_H` are
        stored and, as in the reference, not used by any solver.  Solver options
        (`tol`, `patience`, `initW`, `initH`, and the device-side `precision`,
        `device`, `seed`) travel through **alg_opts."""
        self.n_components = n_components
        self.maxlag = maxlag
        self.n_iter_max = n_iter_max
        self.l1_W = l1_W
        self.l1_H = l1_H
        self.alg_name = alg_name
        self.alg_opts = alg_opts
        self.verbose = verbose

    def fit(self, data):
        """Fits the model (reference model.py:122-176): negativity check, one
        solver object, `loss_hist = [loss0, loss1, ...]`, cumulative
        `time_hist`, early stop through `converged`.

        `time_hist` holds device time measured with CUDA events (a host clock
        around an asynchronous launch would measure nothing).  When `tol == 0`
        the convergence test can never fire (strict `<`, base.py:73), so the
        iterations are submitted in one batch with a single synchronisation."""
        from .datasets import is_device_matrix
        on_device = is_device_matrix(data)        # e.g. datasets.Synthetic(...).device_data(): stays on the GPU
        if not on_device:
            data = np.asarray(data)
            if (data < 0).any():
                raise ValueError('Negative values in data to fit')

        dims = ModelDimensions(data, maxlag=self.maxlag, n_components=self.n_components)
        if self.alg_name not in ALGORITHMS:
            raise KeyError("alg_name %r is not provided by cmfpy_b200 (available: %s)"
                           % (self.alg_name, sorted(ALGORITHMS)))
        algorithm = ALGORITHMS[self.alg_name](data, dims, **self.alg_opts)
        if on_device and algorithm.has_negative:      # the same check (model.py:138-139), reduced on the device
            algorithm.close()
            raise ValueError('Negative values in data to fit')

        self.loss_hist = [algorithm.loss]
        self.time_hist = [0.0]

        # (patience == 1 makes np.diff(loss_hist[-1:]) empty and np.all([]) True: the reference then stops after the
        # first iteration whatever tol is, so only patience >= 2 may skip the test)
        if (algorithm.tol == 0 and algorithm.patience >= 2 and not self.verbose and self.n_iter_max > 0 and
                getattr(algorithm, "batchable", True)):
            losses, secs = algorithm.update_many(self.n_iter_max, return_times=True)
            for loss, dur in zip(losses, secs):
                self.time_hist.append(self.time_hist[-1] + dur)
                self.loss_hist.append(loss)
        else:
            iterations = range(self.n_iter_max)
            if self.verbose:
                try:
                    from tqdm import trange
                    iterations = trange(self.n_iter_max)
                except ImportError:
                    pass
            for itr in iterations:
                losses, secs = algorithm.update_many(1, return_times=True)
                self.time_hist.append(self.time_hist[-1] + secs[0])
                self.loss_hist.append(losses[0])
                if algorithm.converged(self.loss_hist):
                    break

        self._W = algorithm.W
        self._H = algorithm.H
        self._precision = algorithm.precision
        self._device = algorithm.device
        algorithm.close()

    def predict(self):
        """Low-rank reconstruction, N x T (reference model.py:191-200)."""
        return cmf_predict(self.motifs, self.factors,
                           precision="fp32", device=getattr(self, "_device", 0))

    def score(self, data):
        """R^2 = 1 - ||predict - data||^2 / ||data||^2 (reference model.py:202-221)."""
        from .common import score
        return score(self.motifs, self.factors, np.asarray(data), precision="fp32",
                     device=getattr(self, "_device", 0))

    @property
    def motifs(self):
        """W (lags x features x components)."""
        try:
            return self._W
        except AttributeError:
            raise NOT_FITTED_ERROR

    @property
    def factors(self):
        """H (components x timebins)."""
        try:
            return self._H
        except AttributeError:
            raise NOT_FITTED_ERROR

    @property
    def n_features(self):
        return self.motifs.shape[1]

    @property
    def n_timesteps(self):
        return self.factors.shape[1]

    def sort_components(self, data):
        """Orders the components by explanatory power (reference model.py:178-189; there `data` is an undefined
        name - it is an argument here).  The per-component reconstructions and residual norms run on the device."""
        ind = np.argsort(compute_loadings(data, self.motifs, self.factors, device=getattr(self, "_device", 0)))
        self._W = self._W[:, :, ind]
        self._H = self._H[ind, :]
        return ind

    def argsort_units(self):
        """Units ordered by dominant component then peak lag (reference model.py:247-259)."""
        W = self.motifs
        top = np.argmax(W.sum(axis=0), axis=-1)
        peak = np.argmax(W[:, np.arange(self.n_features), top], axis=0)
        return np.lexsort((peak, top))


def compute_loadings(data, W, H, device=0, precision="fp32"):
    """||W_k (*) H_k - data||_F / (||data||_F + EPSILON) for every component k (reference model.py:278-293, which
    calls names that do not exist there).  data is uploaded once; each component costs one fused
    reconstruction + residual reduction on the device and nothing N x T comes back."""
    import ctypes as C
    from . import _lib
    from .common import EPSILON
    lib = _lib.load()
    data = np.ascontiguousarray(data, dtype=np.float64)
    W = np.ascontiguousarray(W, dtype=np.float64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    L, N, K = W.shape
    T = H.shape[1]
    if data.shape != (N, T) or H.shape[0] != K:
        raise ValueError("data, W and H disagree on their dimensions")
    h = C.c_void_p()
    p = _lib.Params(n_features=N, n_components=1, maxlag=L, t_local=T, t_global=T, t_offset=0, device=device,
                    precision=_lib.PRECISIONS[precision], stream=None, denominators=_lib.CMF_DEN_DIRECT)
    _lib.check(lib.cmf_mu_create(C.byref(h), C.byref(p)))
    try:
        _lib.check(lib.cmf_mu_set_data(h, data.ctypes.data, _lib.CMF_F64, _lib.CMF_HOST, T, T))
        ss, neg = C.c_double(0), C.c_int(0)
        _lib.check(lib.cmf_mu_data_stats(h, C.byref(ss), C.byref(neg)))
        data_mag = float(np.sqrt(ss.value))
        loadings = []
        for k in range(K):
            Wk = np.ascontiguousarray(W[:, :, k:k + 1])
            Hk = np.ascontiguousarray(H[k:k + 1, :])
            _lib.check(lib.cmf_mu_set_factors(h, Wk.ctypes.data, Hk.ctypes.data, _lib.CMF_F64, _lib.CMF_HOST, T))
            _lib.check(lib.cmf_mu_recon_loss(h))
            r = C.c_double(0)
            _lib.check(lib.cmf_mu_resid_sumsq(h, C.byref(r)))
            loadings.append(float(np.sqrt(r.value)) / (data_mag + EPSILON))
    finally:
        lib.cmf_mu_destroy(h)
    return loadings


def renormalize(W, H):
    """Rows of H to unit energy, the scale moved into W (reference model.py:296-311); returns new arrays."""
    from .common import EPSILON
    row_norms = np.linalg.norm(H, axis=1) + EPSILON
    return W * row_norms[None, None, :], H / row_norms[:, None]


# ---- model files --------------------------------------------------------------------------
_MODEL_KEYS = ("data", "W", "H", "time_hist", "loss_hist")


def load_cmfjl_model(path):
    """Loads a model saved by cmf.jl (HDF5 / JLD) into a CMF object; returns (data, model) like reference
    model.py:346-363: Julia stores column-major, so `data` and `H` are transposed and the axes of `W` swapped."""
    try:
        import h5py
    except ImportError as e:                       # the reference imports h5py at module level (model.py:7)
        raise ImportError("load_cmfjl_model needs h5py, which is not installed") from e
    with h5py.File(path, "r") as f:
        data = np.array(f["data"]).T
        W = np.swapaxes(np.array(f["W"]), 0, 2)
        L, _, K = W.shape
        model = CMF(K, L)
        model._W = W
        model._H = np.array(f["H"]).T
        model.time_hist = np.array(f["time_hist"])
        model.loss_hist = np.array(f["loss_hist"])
    return data, model


def save_cmfjl_model(path, model, data=None):
    """Writes a fitted model in the HDF5 layout of cmf.jl, the inverse of `load_cmfjl_model` (reference
    model.py:346-363): datasets `data`, `W`, `H`, `time_hist`, `loss_hist`, stored the way Julia's column-major
    arrays appear to a row-major reader (`data`, `H` transposed, the lag and component axes of `W` swapped).
    Needs h5py, like the reference's loader."""
    try:
        import h5py
    except ImportError as e:
        raise ImportError("save_cmfjl_model needs h5py, which is not installed") from e
    with h5py.File(path, "w") as f:
        if data is not None:
            f["data"] = np.ascontiguousarray(np.asarray(data, dtype=np.float64).T)
        f["W"] = np.ascontiguousarray(np.swapaxes(np.asarray(model.motifs, dtype=np.float64), 0, 2))
        f["H"] = np.ascontiguousarray(np.asarray(model.factors, dtype=np.float64).T)
        f["time_hist"] = np.asarray(getattr(model, "time_hist", []), dtype=np.float64)
        f["loss_hist"] = np.asarray(getattr(model, "loss_hist", []), dtype=np.float64)


def _npz_path(path):
    """np.savez appends '.npz' to a name that lacks it; save and load must agree on the file name."""
    path = str(path)
    return path if path.endswith(".npz") else path + ".npz"


def save_model(path, model, data=None):
    """Checkpoint of a fitted model under the dataset names cmf.jl uses (`data`, `W`, `H`, `time_hist`,
    `loss_hist`), in this package's own row-major layouts, as a compressed .npz.  `data` is optional."""
    out = {"W": model.motifs, "H": model.factors,
           "time_hist": np.asarray(getattr(model, "time_hist", [])),
           "loss_hist": np.asarray(getattr(model, "loss_hist", []))}
    if data is not None:
        out["data"] = np.asarray(data)
    np.savez_compressed(_npz_path(path), **out)


def load_model(path, **cmf_kwargs):
    """Inverse of `save_model`: returns (data or None, model).  Resume a fit with
    `CMF(K, L, initW=model.motifs, initH=model.factors, ...).fit(data)`."""
    with np.load(_npz_path(path)) as f:
        W, H = f["W"], f["H"]
        L, _, K = W.shape
        model = CMF(K, L, **cmf_kwargs)
        model._W, model._H = W, H
        model.time_hist, model.loss_hist = list(f["time_hist"]), list(f["loss_hist"])
        data = f["data"] if "data" in f.files else None
    return data, model
