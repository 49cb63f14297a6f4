"""CPU oracle for the cmfpy multiplicative-update (MU) hot path.

TEST INFRASTRUCTURE ONLY.  This module is the checker, never the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``cmfpy_b200/`` imports
it, and the product path raises when its CUDA library is missing.

It is a NumPy restatement of the reference algorithm (``/root/reference``,
degleris1/cmfpy), written from the maths of each call site rather than from its
source text.  Each function cites the reference lines it restates.

Parity pinning (see ``oracle/make_golden.py`` and ``tests/test_oracle.py``):
  * the reference's own known-answer vectors for ``s_dot`` / ``s_T_dot``
    (reference ``tests/test_numeric.py:15-55``) are embedded in the tests;
  * golden ``.npz`` fixtures under ``tests/golden/`` were generated in the build
    container by importing the *unmodified* reference behind two import shims
    (``oracle/ref_shim.py``) and hold X, W0, H0, single-step intermediates and
    101-point loss trajectories; the oracle is checked against all of them.

The arithmetic in the reference lives in NumPy/OpenBLAS DGEMM (un-pinned;
NumPy 2.3.5 here), so agreement is by tolerance (float64: ~1e-12), never bitwise.
"""
from numbers import Integral

import numpy as np

# reference cmfpy/common.py:9  (np.finfo(float).eps; float64 machine epsilon)
EPSILON = float(np.finfo(np.float64).eps)


# --------------------------------------------------------------------------
# shift primitives
# --------------------------------------------------------------------------
def shift_cols(X, lag):
    """Columns of ``X`` that survive a right-shift by ``lag`` (a view).

    reference cmfpy/common.py:89-98.  lag>0 keeps the first T-lag columns,
    lag<=0 keeps the last T+lag columns.
    """
    T = X.shape[1]
    return X[:, :T - lag] if lag > 0 else X[:, -lag:]


def s_dot(A, B, shift):
    """``A @ shift(B, shift)``: columns of B moved right (shift>0) or left
    (shift<0), vacated columns zero.  Output is always ``A.rows x B.cols``.

    reference cmfpy/common.py:13-29; known answers tests/test_numeric.py:15-33.
    """
    T = B.shape[1]
    out = np.zeros((A.shape[0], T), dtype=np.result_type(A, B))
    if shift >= T or -shift >= T:
        return out
    if shift > 0:
        out[:, shift:] = A @ B[:, :T - shift]
    elif shift < 0:
        out[:, :T + shift] = A @ B[:, -shift:]
    else:
        out[:] = A @ B
    return out


def s_T_dot(A, B, shift):
    """``A[:, s:] @ B[:, :T-s].T`` for s>0 (mirror for s<0).

    reference cmfpy/common.py:32-47; known answers tests/test_numeric.py:36-55.
    With A=X, B=H, shift=l this is the W numerator of lag l (mult.py:37).
    """
    T = A.shape[1]
    if shift > 0:
        return A[:, shift:] @ B[:, :T - shift].T
    if shift < 0:
        return A[:, :T + shift] @ B[:, -shift:].T
    return A @ B.T


def shift_and_stack(H, L):
    """(L*K) x T stack, block ``l`` = H shifted right by ``l`` (zero filled).

    reference cmfpy/common.py:101-111.  Row index is ``l*K + k``.
    """
    K, T = H.shape
    S = np.zeros((L * K, T), dtype=H.dtype)
    for l in range(min(L, T)):
        S[l * K:(l + 1) * K, l:] = H[:, :T - l]
    return S


# --------------------------------------------------------------------------
# the three contractions
# --------------------------------------------------------------------------
def cmf_predict(W, H):
    """Reconstruction ``est[:, t] = sum_l W[l] @ H[:, t-l]`` (t-l<0 dropped).

    reference cmfpy/common.py:50-58.  Lag-by-lag slice accumulation: the same
    per-lag rank-K GEMMs as the reference, without its zero-pad copies.
    """
    L, N, K = W.shape
    T = H.shape[1]
    est = np.zeros((N, T), dtype=np.result_type(W, H))
    for l in range(min(L, T)):
        est[:, l:] += W[l] @ H[:, :T - l]
    return est


def cmf_predict_stacked(W, H):
    """Same result as :func:`cmf_predict` via one stacked GEMM
    ``W.transpose(1,0,2).reshape(N, L*K) @ shift_and_stack(H, L)``
    (identity measured against the reference in SURVEY.md section 4)."""
    L, N, K = W.shape
    return W.transpose(1, 0, 2).reshape(N, L * K) @ shift_and_stack(H, L)


def tensor_transconv(W, X):
    """``out[:, t] = sum_l W[l].T @ X[:, t+l]`` (t+l>=T dropped); K x T.

    reference cmfpy/common.py:61-86.
    """
    L, N, K = W.shape
    T = X.shape[1]
    out = np.zeros((K, T), dtype=np.result_type(W, X))
    for l in range(min(L, T)):
        out[:, :T - l] += W[l].T @ X[:, l:]
    return out


def w_terms(X, est, H, L):
    """Numerator / denominator of the W step.

    ``num[l] = X[:, l:] @ H[:, :T-l].T``, ``den[l] = est[:, l:] @ H[:, :T-l].T``
    reference cmfpy/algs/mult.py:27-40.
    """
    N, T = X.shape
    K = H.shape[0]
    num = np.zeros((L, N, K), dtype=np.result_type(X, H))
    den = np.zeros_like(num)
    for l in range(min(L, T)):
        Hl = H[:, :T - l].T
        num[l] = X[:, l:] @ Hl
        den[l] = est[:, l:] @ Hl
    return num, den


def h_terms(X, est, W):
    """Numerator / denominator of the H step: transposed convolutions of X and
    est with W.  reference cmfpy/algs/mult.py:42-48."""
    return tensor_transconv(W, X), tensor_transconv(W, est)


# --------------------------------------------------------------------------
# the solver
# --------------------------------------------------------------------------
def rand_init(X, L, K, rng):
    """``W~U[0,1)^{LxNxK}``, ``H~U[0,1)^{KxT}`` rescaled by sqrt(alpha) with
    ``alpha = <X, est> / ||est||^2``.

    reference cmfpy/algs/base.py:78-88.  The reference draws from the global
    unseeded ``numpy.random``; here the generator is explicit so that runs are
    reproducible (parity tests always pass initW/initH explicitly).
    """
    N, T = X.shape
    W = rng.random((L, N, K))
    H = rng.random((K, T))
    est = cmf_predict(W, H)
    alpha = float((X * est).sum() / np.linalg.norm(est) ** 2)
    return np.sqrt(alpha) * W, np.sqrt(alpha) * H


class MultUpdateOracle:
    """Duck-type of reference ``MultUpdate`` (cmfpy/algs/mult.py:7-48 on top of
    cmfpy/algs/base.py:12-97).

    ``dtype=np.float64`` restates the reference exactly; ``np.float32`` runs the
    same steps in single precision (used to calibrate what an fp32 device path
    can be expected to reach).  ``reuse_est=True`` skips the reference's third
    reconstruction per iteration by keeping the one computed for the loss; the
    results are identical because the inputs of the skipped call are identical
    (SURVEY.md section 3.2).
    """

    def __init__(self, data, maxlag, n_components, initW=None, initH=None,
                 tol=1e-5, patience=3, dtype=np.float64, rng=None,
                 reuse_est=False):
        if patience < 1 or not isinstance(patience, Integral):   # base.py:20-21
            raise ValueError("Patience must be a positive integer.")
        self.dtype = np.dtype(dtype)
        self.X = np.asarray(data, dtype=self.dtype)
        self.normX = float(np.linalg.norm(self.X))               # base.py:25
        self.tol, self.patience = tol, patience
        self.n_features, self.n_timepoints = self.X.shape
        self.maxlag, self.n_components = maxlag, n_components
        if initW is None or initH is None:                       # base.py:37-41
            rng = rng if rng is not None else np.random.default_rng()
            W, H = rand_init(self.X, maxlag, n_components, rng)
        else:
            W, H = initW, initH
        self.W = np.asarray(W, dtype=self.dtype)
        self.H = np.asarray(H, dtype=self.dtype)
        self.reuse_est = reuse_est
        self.eps = self.dtype.type(EPSILON)
        self.cache_resids()

    def cache_resids(self):                                      # base.py:57-62
        self.est = cmf_predict(self.W, self.H)
        self.resids = self.est - self.X

    @property
    def loss(self):                                              # base.py:90-97
        return float(np.linalg.norm(self.resids) / self.normX)

    def converged(self, loss_hist):                              # base.py:64-76
        d = np.diff(loss_hist[-self.patience:])
        return bool(np.all(np.abs(d) < self.tol))

    def update(self):                                            # mult.py:15-25
        L = self.maxlag
        est = self.est if self.reuse_est else cmf_predict(self.W, self.H)
        num, den = w_terms(self.X, est, self.H, L)               # mult.py:27-40
        self.W = self.W * num / (den + self.eps)                 # mult.py:18
        est = cmf_predict(self.W, self.H)                        # mult.py:43
        num, den = h_terms(self.X, est, self.W)                  # mult.py:45-46
        self.H = self.H * num / (den + self.eps)                 # mult.py:22
        self.cache_resids()                                      # mult.py:24
        return self.loss                                         # mult.py:25


def fit(data, maxlag, n_components, n_iter_max=100, **kw):
    """The bookkeeping of reference ``CMF.fit`` (cmfpy/model.py:122-176):
    negativity check, ``loss_hist=[loss0]`` then one entry per update, early
    stop through ``converged``.  Returns (W, H, loss_hist)."""
    data = np.asarray(data)
    if (data < 0).any():                                         # model.py:138
        raise ValueError("Negative values in data to fit")
    alg = MultUpdateOracle(data, maxlag, n_components, **kw)
    loss_hist = [alg.loss]                                       # model.py:149
    for _ in range(n_iter_max):                                  # model.py:157
        loss_hist.append(alg.update())
        if alg.converged(loss_hist):                             # model.py:171
            break
    return alg.W, alg.H, loss_hist


# --------------------------------------------------------------------------
# Gradient solvers (reference cmfpy/algs/gradient_descent.py)
# --------------------------------------------------------------------------
def lipschitz_W(H, L):
    """Largest eigenvalue of the block-Toeplitz matrix of the lag autocorrelations of H
    (gradient_descent.py:54-69).  The reference calls ``eigh(hW, eigvals=(n-1, n-1))``, a keyword SciPy >= 1.14
    no longer accepts; ``subset_by_index`` is its replacement and returns the same eigenvalue."""
    from scipy.linalg import eigh
    K = H.shape[0]
    first_row = [s_T_dot(H, H, l) for l in range(L)]             # :56
    hW = np.empty((K * L, K * L))
    for i in range(L):                                           # :58-67
        for j in range(L):
            blk = first_row[j - i] if i <= j else first_row[i - j].T
            hW[i * K:(i + 1) * K, j * K:(j + 1) * K] = blk
    return float(eigh(hW, subset_by_index=[K * L - 1, K * L - 1], eigvals_only=True)[0])   # :69


def projected_step(x, dx, ss):
    """x <- max(x - ss dx, 0) (gradient_descent.py:148-159); returns a new array."""
    return np.maximum(x - ss * dx, 0.0)


class GradDescentOracle(MultUpdateOracle):
    """Duck-type of reference ``GradDescent`` (gradient_descent.py:15-123); ``block=True`` gives ``BlockDescent``
    (:126-147)."""

    def __init__(self, data, maxlag, n_components, step_decrement=5., block=False, **kw):
        super().__init__(data, maxlag, n_components, **kw)
        self.W, self.H = self.W.copy(), self.H.copy()
        self.step_size = 1e-4                                    # :29
        self.step_decrement = step_decrement
        self.block = block
        self.cache_gW()                                          # :37-38
        self.cache_gH()

    def cache_gW(self):                                          # :40-45
        self.gW = np.stack([s_T_dot(self.resids, self.H, l) for l in range(self.maxlag)])
        return self.gW

    def cache_gH(self):                                          # :47-52
        self.gH = tensor_transconv(self.W, self.resids)
        return self.gH

    def lipschitz_W(self):
        return lipschitz_W(self.H, self.maxlag)

    def update(self):
        if not self.block:                                       # :81-92
            lam = self.lipschitz_W()
            self.W = projected_step(self.W, self.gW, 1.0 / lam)
            self.H = projected_step(self.H, self.gH, self.step_size)
            self.cache_resids()
            self.cache_gW()
            self.cache_gH()
        else:                                                    # :132-147
            self.W = projected_step(self.W, self.gW, 1.0 / self.lipschitz_W())
            self.cache_resids()
            self.cache_gH()
            self.H = projected_step(self.H, self.gH, self.step_size)
            self.cache_resids()
            self.cache_gW()
        return self.loss

    def converged(self, loss_hist):                              # :94-113
        d_loss = np.diff(loss_hist[-self.patience:])
        if d_loss[-1] > 0:
            self.step_size /= self.step_decrement
            return False
        return bool(np.all(np.abs(d_loss) < self.tol))


# --------------------------------------------------------------------------
# HALS (reference cmfpy/algs/hals.py on top of cmfpy/algs/accelerated.py)
# --------------------------------------------------------------------------
FACTOR_MIN = 0.0                                                 # common.py:10


class HALSOracle(MultUpdateOracle):
    """Duck-type of reference ``HALSUpdate``: hierarchical alternating least squares, one coordinate block at a
    time with the residual ``est - X`` kept current after every block.

    The reference removes a block's own contribution from the residual, solves for the block and adds the new
    contribution back (hals.py:89-96, 129-157, 165-181).  With r the residual BEFORE the removal that is

        W[l,:,k] <- max((W[l,:,k] ||h||^2 - r h) / (||h||^2 + eps), 0),  h = H[k] shifted right by l   (:89-105)
        H[k,t]   <- max((H[k,t] ||Wk||^2 - <Wk, r[:, t:t+L]>) / (||Wk||^2 + eps), 0)                   (:129-157)

    followed by r += (new - old) * block; both forms are the same arithmetic up to rounding.  Order of the sweeps:
    components outermost, lags inside (:80-84, :115-123); within the H sweep the entries t = l (mod L), t < T - L are
    independent (disjoint windows, "batches") and the entry t = T - L + l, whose motif is cut off by the end of the
    data, follows on its own (:122-123, :165-181).
    """

    def __init__(self, data, maxlag, n_components, max_iter=1, weightW=1, weightH=1, stop_thresh=0, **kw):
        super().__init__(data, maxlag, n_components, **kw)
        self.W, self.H = self.W.copy(), self.H.copy()
        self.max_iter, self.stop_thresh, self.weightW, self.weightH = max_iter, stop_thresh, weightW, weightH
        if max_iter * min(1, weightH, weightW) < 1:              # accelerated.py:47-48
            raise ValueError("Requires at least 1 iteration for both W and H.")

    # -- one sweep over W (hals.py:40-48, 78-105) ---------------------------
    def sweep_W(self):
        L, N, K = self.W.shape
        T = self.n_timepoints
        r = self.resids
        for k in range(K):
            for l in range(L):
                h = np.zeros(T)
                h[l:] = self.H[k, :T - l]
                hn2 = float(h @ h)
                old = self.W[l, :, k].copy()
                new = np.maximum((old * hn2 - r @ h) / (hn2 + EPSILON), FACTOR_MIN)
                r += np.outer(new - old, h)
                self.W[l, :, k] = new

    # -- one sweep over H (hals.py:54-70, 113-181) ----------------------------
    def sweep_H(self):
        L, N, K = self.W.shape
        T = self.n_timepoints
        r = self.resids
        w2 = (self.W ** 2).sum(axis=1).T                        # K x L: squared norms along the features (:58)
        for k in range(K):
            Wk = self.W[:, :, k].T                               # N x L
            for l in range(L):
                for t in range(l, T - L, L):                     # the batch: disjoint windows (:32-33, 126-157)
                    win = r[:, t:t + L]
                    old = self.H[k, t]
                    new = max((old * w2[k].sum() - float((Wk * win).sum())) / (w2[k].sum() + EPSILON), FACTOR_MIN)
                    win += (new - old) * Wk
                    self.H[k, t] = new
                t = T - L + l                                    # the entry whose motif is cut by the end (:123, 165-181)
                m = T - t
                win = r[:, t:t + L]
                n2 = w2[k, :m].sum()
                old = self.H[k, t]
                new = max((old * n2 - float((Wk[:, :m] * win).sum())) / (n2 + EPSILON), FACTOR_MIN)
                win += (new - old) * Wk[:, :m]
                self.H[k, t] = new

    def _accelerated(self, var_name, weight, sweep):             # accelerated.py:50-69
        prev = getattr(self, var_name).copy()
        sweep()
        init_diff = diff = float(np.linalg.norm(getattr(self, var_name) - prev))
        itr = 1
        while itr < self.max_iter * weight and diff > self.stop_thresh * init_diff:
            itr += 1
            prev = getattr(self, var_name).copy()
            sweep()
            diff = float(np.linalg.norm(getattr(self, var_name) - prev))

    def update(self):                                            # accelerated.py:71-84
        self._accelerated("W", self.weightW, self.sweep_W)
        self._accelerated("H", self.weightH, self.sweep_H)
        self.cache_resids()
        return self.loss


# --------------------------------------------------------------------------
# T-sharded restatement (what the multi-GPU path must reproduce)
# --------------------------------------------------------------------------
def sharded_update(X, W, H, n_shards):
    """One MU iteration computed shard-by-shard over the time axis, using only
    an (L-1)-column halo of H on each side, a right halo of X, and a sum of the
    per-shard W terms.  Mathematically identical to ``MultUpdateOracle.update``
    (float rounding differs by summation order only).  Used by the tests to pin
    the halo / truncation rules of SURVEY.md section 8(e).
    Returns (W_new, H_new, loss)."""
    L, N, K = W.shape
    T = X.shape[1]
    assert T % n_shards == 0 and T // n_shards >= L - 1
    Tg, h = T // n_shards, L - 1
    eps = X.dtype.type(EPSILON)

    def hwin(Hfull, g):      # H[:, t0-h : t1+h) with zeros outside [0, T)
        t0, t1 = g * Tg, (g + 1) * Tg
        out = np.zeros((K, Tg + 2 * h), dtype=Hfull.dtype)
        lo, hi = max(t0 - h, 0), min(t1 + h, T)
        out[:, lo - (t0 - h):hi - (t0 - h)] = Hfull[:, lo:hi]
        return out

    def xwin(g):             # X[:, t0 : t1+h) with zeros beyond T
        t0, t1 = g * Tg, (g + 1) * Tg
        out = np.zeros((N, Tg + h), dtype=X.dtype)
        hi = min(t1 + h, T)
        out[:, :hi - t0] = X[:, t0:hi]
        return out

    def est_ext(Wc, Hw, g):  # est[:, t0 : t1+h), forced to zero beyond T
        full = cmf_predict(Wc, Hw)[:, h:]          # drop the left-halo columns
        t0 = g * Tg
        nvalid = min(T - t0, Tg + h)
        full[:, nvalid:] = 0
        return full

    num = np.zeros_like(W)
    den = np.zeros_like(W)
    for g in range(n_shards):                      # W terms: local partial sums
        Hw = hwin(H, g)
        Xw = xwin(g)[:, :Tg]
        Ew = est_ext(W, Hw, g)[:, :Tg]
        for l in range(L):
            Hl = Hw[:, h - l:h - l + Tg].T
            num[l] += Xw @ Hl
            den[l] += Ew @ Hl
    Wn = W * num / (den + eps)                     # after the all-reduce

    Hn = np.empty_like(H)
    for g in range(n_shards):                      # H terms: local, right halo
        Hw = hwin(H, g)
        Xw = xwin(g)
        Ew = est_ext(Wn, Hw, g)
        nH = np.zeros((K, Tg), dtype=X.dtype)
        dH = np.zeros((K, Tg), dtype=X.dtype)
        for l in range(L):
            nH += Wn[l].T @ Xw[:, l:l + Tg]
            dH += Wn[l].T @ Ew[:, l:l + Tg]
        Hn[:, g * Tg:(g + 1) * Tg] = H[:, g * Tg:(g + 1) * Tg] * nH / (dH + eps)

    sq = 0.0
    for g in range(n_shards):                      # loss: own columns only
        Ew = est_ext(Wn, hwin(Hn, g), g)[:, :Tg]
        sq += float(((Ew - xwin(g)[:, :Tg]) ** 2).sum())
    return Wn, Hn, float(np.sqrt(sq) / np.linalg.norm(X))
