"""Multiplicative updates over several GPUs of one box from ONE process:
`CMF(K, L, devices=[0, 1, ...]).fit(X)`.

The reference has a single solver plug-in point, `ALGORITHMS[alg_name](data, dims, **alg_opts)`
(cmfpy/model.py:80-82, :146); `devices` travels through `alg_opts` like every other solver option.
The work is split as SURVEY.md section 8(e) describes: GPU g owns a contiguous range of time columns
of X, est and H; W and the W-step terms are replicated.  Every shard is one solver handle of the C ABI
on its own GPU; the handles are wired to one another through peer memory
(`cmf_mu_peer_attach_local`) and one host thread per device drives `cmf_mu_step_sharded`, whose
collectives - the reduce-scatter / W update / all-gather kernel, the halo pushes, the loss ring -
are the library's own kernels over NVLink.  No torch, no NCCL.  (One process per GPU under torchrun
is `cmfpy_b200.dist.ShardedMultUpdate`.)
"""
import ctypes as C
import threading
from numbers import Integral

import numpy as np

from .. import _lib
from .base import _as_host_matrix, resolve_precision


def shard_ranges(T, G, tail_columns=0, align=256):
    """Contiguous column ranges [t0, t1) of G shards.  `tail_columns` = 0: sizes differing by at most one.
    Otherwise the LAST shard ends up about that many columns shorter than the others (which share what it gives up,
    in multiples of `align`): with the Gram-route denominators only the shard that sees the end of the data computes
    the end-of-data corrections - a few small kernels per iteration while every other rank waits at the next
    exchange; `tail_handicap` estimates their cost in columns."""
    base, rem = divmod(T, G)
    sizes = [base + (1 if g < rem else 0) for g in range(G)]
    if G > 1 and tail_columns > 0:
        per = int(round(tail_columns / G / align)) * align          # what each of the other shards takes over
        if per > 0 and per * (G - 1) <= sizes[-1] // 4:
            for g in range(G - 1):
                sizes[g] += per
            sizes[-1] -= per * (G - 1)
    out, t0 = [], 0
    for n in sizes:
        out.append((t0, t0 + n))
        t0 += n
    return out


def tail_handicap(N, K, L, T, G, gram, num_sms=148):
    """Columns of MU work that the end-of-data corrections of the Gram route cost the last of G time shards per
    iteration (see shard_ranges).  They are dominated by two reconstructions of ONE 256-column time tile, which
    occupy ceil(N / 128) SMs where an iteration's two full contractions use all of them - 256 * num_sms /
    ceil(N / 128) column-equivalents - plus about half as much again for the small FFMA contractions after them
    (measured at config C: 0.26 ms of the last rank per iteration = 7 800 columns; the estimate gives 7 104)."""
    if G < 2 or not gram:
        return 0
    return int(1.5 * 256 * num_sms / -(-N // 128))


class MultiGpuMultUpdate:
    """Duck type of reference MultUpdate (cmfpy/algs/mult.py:7-48) on several GPUs."""

    batchable = True

    def __init__(self, data, dims, initW=None, initH=None, tol=1e-5, patience=3, precision="auto",
                 devices=(0,), seed=None, denominators="auto", **unused):
        if unused:
            raise TypeError("unexpected solver options for a multi-GPU solve: %s" % sorted(unused))
        if patience < 1 or not isinstance(patience, Integral):            # reference base.py:20-21
            raise ValueError("Patience must be a positive integer.")
        devices = [int(d) for d in devices]
        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("devices must be distinct CUDA ordinals")
        if denominators not in _lib.DENOMINATORS:
            raise ValueError("denominators must be one of %s" % sorted(_lib.DENOMINATORS))
        self._lib = lib = _lib.load()
        self.devices, self.G = devices, len(devices)
        self.device = devices[0]
        self.tol, self.patience = tol, patience
        for k, v in dims:                                                 # reference base.py:33-34
            setattr(self, k, v)
        N, T, K, L = self.n_features, self.n_timepoints, self.n_components, self.maxlag
        self.precision = resolve_precision(lib, precision, N, K, L)
        from ..datasets import is_device_matrix
        if is_device_matrix(data):          # the shards are loaded from one host copy (each GPU takes its columns)
            data = data.to_host(np.float32) if hasattr(data, "to_host") else np.asarray(data.cpu())
        X = _as_host_matrix(data, "data")
        if X.shape != (N, T):
            raise ValueError("data has shape %s, dimensions say %s" % (X.shape, (N, T)))
        self._X = X
        gram = denominators == "gram" or (denominators == "auto" and self.precision != "fp32" and
                                          2.0 * N * K * L * (T / self.G) >= 2e11 and N >= 4 * K)
        self.ranges = shard_ranges(T, self.G, tail_handicap(N, K, L, T, self.G, gram))
        if self.G > 1 and min(t1 - t0 for t0, t1 in self.ranges) < max(L - 1, 1):
            raise ValueError("each of the %d time shards needs at least L-1 = %d columns (T = %d)" % (self.G, L - 1, T))
        self._h = [C.c_void_p() for _ in devices]
        self._attached = False
        try:
            self._build(X, initW, initH, seed, denominators)
        except Exception:
            self.close()
            raise

    # -- construction ---------------------------------------------------------------------
    def _build(self, X, initW, initH, seed, denominators):
        lib = self._lib
        N, T, K, L = self.n_features, self.n_timepoints, self.n_components, self.maxlag
        ss_total, neg = 0.0, False
        for g, (t0, t1) in enumerate(self.ranges):
            p = _lib.Params(n_features=N, n_components=K, maxlag=L, t_local=t1 - t0, t_global=T, t_offset=t0,
                            device=self.devices[g], precision=_lib.PRECISIONS[self.precision], stream=None,
                            denominators=_lib.DENOMINATORS[denominators])
            _lib.check(lib.cmf_mu_create(C.byref(self._h[g]), C.byref(p)))
            ncols = min(t1 - t0 + L - 1, T - t0)                          # own columns + the static right halo of X
            _lib.check(lib.cmf_mu_set_data(self._h[g], X.ctypes.data + t0 * X.itemsize, _lib.np_dtype_code(X),
                                           _lib.CMF_HOST, T, ncols))
            ss, ng = C.c_double(0), C.c_int(0)
            _lib.check(lib.cmf_mu_data_stats(self._h[g], C.byref(ss), C.byref(ng)))
            ss_total += ss.value
            neg = neg or bool(ng.value)
        self.normX = float(np.sqrt(ss_total))                            # reference base.py:25
        self.has_negative = neg
        for h in self._h:
            _lib.check(lib.cmf_mu_set_norm_x(h, self.normX))
        arr = (C.c_void_p * self.G)(*[h.value for h in self._h])
        for g in range(self.G):
            _lib.check(lib.cmf_mu_peer_attach_local(self._h[g], g, self.G, arr))
        self._attached = True
        if initW is None or initH is None:                               # reference base.py:37-41
            self._rand_init(seed)
        else:
            self._set_factors(initW, initH)
        self._parallel(lambda g: _lib.check(lib.cmf_mu_recon_loss(self._h[g])))
        self._loss = None

    def _set_factors(self, W0, H0):
        lib = self._lib
        N, T, K, L = self.n_features, self.n_timepoints, self.n_components, self.maxlag
        W0, H0 = _as_host_matrix(W0, "initW"), _as_host_matrix(H0, "initH")
        if W0.shape != (L, N, K) or H0.shape != (K, T):
            raise ValueError("initW/initH must have shapes %s and %s" % ((L, N, K), (K, T)))
        if W0.dtype != H0.dtype:
            H0 = H0.astype(W0.dtype)
        for g, (t0, t1) in enumerate(self.ranges):
            _lib.check(lib.cmf_mu_set_factors(self._h[g], W0.ctypes.data, H0.ctypes.data + t0 * H0.itemsize,
                                              _lib.np_dtype_code(W0), _lib.CMF_HOST, T))
        self._parallel(lambda g: _lib.check(lib.cmf_mu_halo_exchange_peer(self._h[g])))

    def _rand_init(self, seed):
        """reference base.py:78-88 on the shards: alpha = <X, est> / ||est||^2 summed over the GPUs."""
        lib = self._lib
        rng = np.random.default_rng(seed)
        N, T, K, L = self.n_features, self.n_timepoints, self.n_components, self.maxlag
        self._set_factors(rng.random((L, N, K), dtype=np.float32), rng.random((K, T), dtype=np.float32))
        xe, ee = [0.0] * self.G, [0.0] * self.G

        def stats(g):
            a, b = C.c_double(0), C.c_double(0)
            _lib.check(lib.cmf_mu_init_stats(self._h[g], C.byref(a), C.byref(b)))
            xe[g], ee[g] = a.value, b.value
        self._parallel(stats)
        s = float(np.sqrt(sum(xe) / sum(ee)))
        for h in self._h:
            _lib.check(lib.cmf_mu_scale_factors(h, s, s))

    def _parallel(self, fn):
        """fn(g) for every shard, one host thread per device (ctypes releases the GIL; the exchange kernels of
        different GPUs wait for one another, so the calls must be in flight together)."""
        if self.G == 1:
            fn(0)
            return
        errs = [None] * self.G

        def run(g):
            try:
                fn(g)
            except BaseException as e:      # noqa: BLE001 - re-raised on the caller's thread
                errs[g] = e
        ts = [threading.Thread(target=run, args=(g,)) for g in range(self.G)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in errs:
            if e is not None:
                raise e

    # -- reference interface --------------------------------------------------------------
    def update_many(self, n_steps, return_times=False):
        lib = self._lib
        if n_steps <= 0:
            return ([], []) if return_times else []
        losses = [np.empty(n_steps, dtype=np.float64) for _ in range(self.G)]
        import time
        t0 = time.perf_counter()
        if self.G == 1:
            _lib.check(lib.cmf_mu_step(self._h[0], n_steps, losses[0].ctypes.data_as(C.POINTER(C.c_double)), None))
        else:
            self._parallel(lambda g: _lib.check(lib.cmf_mu_step_sharded(
                self._h[g], n_steps, losses[g].ctypes.data_as(C.POINTER(C.c_double)))))
        dt = (time.perf_counter() - t0) / n_steps       # the calls return after a device synchronisation
        out = [float(x) for x in losses[0]]
        self._loss = out[-1]
        if return_times:
            return out, [dt] * n_steps
        return out

    def update(self):
        return self.update_many(1)[0]

    def converged(self, loss_hist):                                      # reference base.py:64-76
        d_loss = np.diff(loss_hist[-self.patience:])
        return bool(np.all(np.abs(d_loss) < self.tol))

    @property
    def loss(self):                                                      # reference base.py:90-97
        if self._loss is None:
            ss = 0.0
            for h in self._h:
                v = C.c_double(0)
                _lib.check(self._lib.cmf_mu_resid_sumsq(h, C.byref(v)))
                ss += v.value
            self._loss = float(np.sqrt(ss) / self.normX)
        return self._loss

    @property
    def X(self):
        return self._X

    @property
    def W(self):
        L, N, K = self.maxlag, self.n_features, self.n_components
        out = np.empty((L, N, K), dtype=np.float64)
        _lib.check(self._lib.cmf_mu_get_W(self._h[0], out.ctypes.data, _lib.CMF_F64, _lib.CMF_HOST))
        return out

    @property
    def H(self):
        K, T = self.n_components, self.n_timepoints
        out = np.empty((K, T), dtype=np.float64)
        for g, (t0, t1) in enumerate(self.ranges):
            _lib.check(self._lib.cmf_mu_get_H(self._h[g], out.ctypes.data + t0 * 8, _lib.CMF_F64, _lib.CMF_HOST, T))
        return out

    @property
    def est(self):
        N, T = self.n_features, self.n_timepoints
        out = np.empty((N, T), dtype=np.float64)
        for g, (t0, t1) in enumerate(self.ranges):
            _lib.check(self._lib.cmf_mu_get_est(self._h[g], out.ctypes.data + t0 * 8, _lib.CMF_F64, _lib.CMF_HOST, T))
        return out

    @property
    def resids(self):
        return self.est - self._X

    @property
    def path_name(self):
        return self._lib.cmf_mu_path_name(self._h[0]).decode()

    @property
    def launch_count(self):
        total = 0
        for h in self._h:
            n = C.c_longlong(0)
            _lib.check(self._lib.cmf_mu_launch_count(h, C.byref(n)))
            total += n.value
        return total

    def close(self):
        hs = [h for h in getattr(self, "_h", []) if h is not None and h.value]
        if getattr(self, "_attached", False):
            for h in hs:
                self._lib.cmf_mu_peer_detach(h)
            self._attached = False
        for h in hs:
            self._lib.cmf_mu_destroy(h)
        self._h = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
