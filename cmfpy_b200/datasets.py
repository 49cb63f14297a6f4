"""The step before the solver, on the device.

`Synthetic` mirrors reference cmfpy/datasets/synthetic.py:7-46 (same
constructor arguments, `name`, `W`, `H`, `noise`, `data`, `generate()`);
`spectrogram` is the `generate` step of reference
cmfpy/datasets/vox_celeb.py:58-104.  Both leave their result in device memory
as a `DeviceMatrix`, which `CMF.fit` and the solvers accept in place of a
NumPy array, so a data set of several GiB never crosses PCIe; the NumPy views
the reference exposes are read back on first access.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


class DeviceMatrix:
    """A row-major float32 matrix in device memory owned by libcmf_b200 (`cmf_dmat_t`).  Exposes `shape`,
    `__cuda_array_interface__` (so torch / cupy can wrap it without a copy) and `to_host()`."""

    def __init__(self, handle):
        self._lib = _lib.load()
        self._h = handle
        ptr, rows, cols, ld, dev = C.c_void_p(), C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_int()
        _lib.check(self._lib.cmf_dmat_info(self._h, C.byref(ptr), C.byref(rows), C.byref(cols), C.byref(ld),
                                           C.byref(dev)))
        self.ptr, self.shape, self.ld, self.device = ptr.value, (rows.value, cols.value), ld.value, dev.value
        self.dtype = np.dtype(np.float32)
        self.ndim = 2

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f4", "data": (int(self.ptr), False), "version": 3,
                "strides": (self.ld * 4, 4)}

    def to_host(self, dtype=np.float64):
        out = np.empty(self.shape, dtype=dtype)
        _lib.check(self._lib.cmf_dmat_get(self._h, out.ctypes.data, _lib.np_dtype_code(out), self.shape[1]))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cmf_dmat_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def is_device_matrix(a):
    return hasattr(a, "__cuda_array_interface__") and not isinstance(a, np.ndarray)


def describe_device_matrix(a):
    """(pointer, leading dimension in elements, (rows, cols)) of a 2-D row-major float32 device array."""
    d = a.__cuda_array_interface__
    if d["typestr"] != "<f4" or len(d["shape"]) != 2:
        raise ValueError("device data must be a 2-D float32 array")
    rows, cols = d["shape"]
    strides = d.get("strides") or (cols * 4, 4)
    if strides[1] != 4 or strides[0] % 4:
        raise ValueError("device data must be row-major")
    return int(d["data"][0]), strides[0] // 4, (int(rows), int(cols))


class Synthetic:
    """Synthetic data (reference datasets/synthetic.py:7-39), generated on the GPU.

    Sparse non-negative H, one Gaussian-bump motif per feature on a random
    component, uniform noise, `data = cmf_predict(W, H) + noise`.  The
    reference draws from NumPy's generators - partly the global, unseeded one
    (:28, :43) - so its stream cannot be reproduced; here every value is a
    function of (seed, global element index), the same for any device and any
    time sharding (`t_offset`, `t_local` select the columns this object holds).
    `W`, `H`, `noise`, `data` are float64 NumPy arrays as in the reference, read
    back on first access; `device_data()` / `device_generate()` hand the
    N x T matrix to a solver without leaving the device."""

    def __init__(self, n_components=3, n_features=100, n_lags=100, n_timebins=10000,
                 H_sparsity=0.9, noise_scale=1.0, seed=None, device=0, precision="auto",
                 t_offset=0, t_local=None):
        self.name = "synthetic"
        self._lib = _lib.load()
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed) & (2**64 - 1)
        t_local = n_timebins - t_offset if t_local is None else t_local
        from .algs.base import resolve_precision
        precision = resolve_precision(self._lib, precision, n_features, n_components, n_lags)
        self._dims = (n_components, n_features, n_lags, t_local)
        p = _lib.SynthParams(n_components=n_components, n_features=n_features, n_lags=n_lags,
                             n_timebins=n_timebins, t_offset=t_offset, t_local=t_local,
                             H_sparsity=H_sparsity, noise_scale=noise_scale, seed=self.seed,
                             device=device, precision=_lib.PRECISIONS[precision])
        self._h = C.c_void_p()
        _lib.check(self._lib.cmf_synth_create(C.byref(self._h), C.byref(p)))
        self._cache = {}

    def _get(self, what, shape):
        if what not in self._cache:
            out = np.empty(shape, dtype=np.float64)
            _lib.check(self._lib.cmf_synth_get(self._h, what, out.ctypes.data, _lib.CMF_F64, shape[-1]))
            self._cache[what] = out
        return self._cache[what]

    @property
    def W(self):
        K, N, L, _ = self._dims
        return self._get(_lib.SYNTH_W, (L, N, K))

    @property
    def H(self):
        K, _, _, T = self._dims
        return self._get(_lib.SYNTH_H, (K, T))

    @property
    def noise(self):
        _, N, _, T = self._dims
        return self._get(_lib.SYNTH_NOISE, (N, T))

    @property
    def data(self):
        _, N, _, T = self._dims
        return self._get(_lib.SYNTH_DATA, (N, T))

    def generate(self):
        """reference synthetic.py:38-39: `data + noise` (the noise enters a second time)."""
        _, N, _, T = self._dims
        return self._get(_lib.SYNTH_GENERATE, (N, T))

    def _matrix(self, what):
        m = C.c_void_p()
        _lib.check(self._lib.cmf_synth_matrix(self._h, what, C.byref(m)))
        return DeviceMatrix(m)

    def device_data(self):
        """`data` as a device matrix (shares the generator's buffer)."""
        return self._matrix(_lib.SYNTH_DATA)

    def device_generate(self):
        """`generate()` as a device matrix."""
        return self._matrix(_lib.SYNTH_GENERATE)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cmf_synth_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tukey_window(n, alpha=0.25):
    """scipy.signal.get_window(('tukey', 0.25), n) (periodic, as `spectrogram` requests it): the default window
    of scipy.signal.spectrogram, which the reference does not override (vox_celeb.py:92-98)."""
    if n <= 1 or alpha <= 0:
        return np.ones(max(n, 0))
    m = n + 1                                   # periodic = symmetric window of n + 1 points without the last
    if alpha >= 1:
        k = np.arange(m)
        return (0.5 * (1 - np.cos(2 * np.pi * k / (m - 1))))[:n]
    k = np.arange(m, dtype=np.float64)
    width = int(np.floor(alpha * (m - 1) / 2.0))
    w = np.ones(m)
    n1, n3 = k[:width + 1], k[m - width - 1:]
    w[:width + 1] = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (m - 1))))
    w[m - width - 1:] = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (m - 1))))
    return w[:n]


def spectrogram(audio, sampling_rate, seg_length=20e-3, overlap=0.3, normalize=True, window=None, device=0,
                to_host=False):
    """The spectrogram of reference VoxCeleb.generate (vox_celeb.py:58-104) on the device:
    `scipy.signal.spectrogram(audio, fs, nperseg=round(seg_length*fs), noverlap=round(nperseg*overlap))` with
    scipy's defaults, then (normalize) every frequency bin divided by its standard deviation over time
    (`StandardScaler(with_mean=False)`).  Returns a `DeviceMatrix` (frequency bins x segments) ready for
    `CMF.fit`, or a float64 NumPy array with `to_host=True`."""
    lib = _lib.load()
    nperseg = round(int(seg_length * sampling_rate))          # vox_celeb.py:89-90
    noverlap = round(int(nperseg * overlap))
    if window is None:
        window = tukey_window(nperseg)
    window = np.ascontiguousarray(window, dtype=np.float64)
    if window.shape != (nperseg,):
        raise ValueError("window must have nperseg = %d entries" % nperseg)
    if is_device_matrix(audio):
        d = audio.__cuda_array_interface__
        if d["typestr"] != "<f4" or (d.get("strides") not in (None, (4,))):
            raise ValueError("device audio must be a contiguous float32 vector")
        ptr, dt, mem, n = int(d["data"][0]), _lib.CMF_F32, _lib.CMF_DEVICE, int(np.prod(d["shape"]))
    else:
        a = np.ascontiguousarray(audio)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = a.ravel()
        ptr, dt, mem, n = a.ctypes.data, _lib.np_dtype_code(a), _lib.CMF_HOST, a.size
    m = C.c_void_p()
    _lib.check(lib.cmf_spectrogram(ptr, dt, mem, n, float(sampling_rate), nperseg, noverlap,
                                   window.ctypes.data, 1 if normalize else 0, device, C.byref(m)))
    S = DeviceMatrix(m)
    return S.to_host() if to_host else S
