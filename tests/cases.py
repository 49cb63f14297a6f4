"""Seeded synthetic cases shared by the golden generator (oracle/make_golden.py),
the parity tests and bench.py.  Inputs are rounded to float32 before anybody
sees them, so the float64 reference and the fp32 device path start from
bit-identical numbers.

Shapes follow BASELINE.json `configs` (A..E) plus small odd shapes covering the
edge cases of SURVEY.md section 7 (K not a multiple of 8, L > T/4, L = 1, K = 1,
N not a multiple of any tile size).
"""
import numpy as np

# name: (N, T, K, L, kind, n_iter, store_full)
CASES = {
    # BASELINE config 1 (the reference's own CPU-runnable case)
    "A":        (100, 250, 3, 20, "planted", 100, True),
    # the same shape with X from the reference's own Synthetic generator
    # (inputs live in the .npz: the generator uses the global numpy RNG)
    "A_synth":  (100, 250, 3, 20, "ref_synthetic", 100, True),
    # small odd shapes
    "odd_k5":   (37, 301, 5, 7, "planted", 30, True),
    "l_gt_t4":  (19, 64, 2, 24, "uniform", 30, True),
    "lag1":     (33, 200, 4, 1, "uniform", 30, True),
    "k1":       (16, 128, 1, 9, "planted", 30, True),
    "n1":       (1, 96, 2, 5, "uniform", 20, True),
    "zeros":    (24, 160, 3, 6, "zeros_rows", 20, True),
    # shapes the tensor-core path accepts (25 <= K <= 32), with edge cases:
    # N not a multiple of 32, L not a multiple of 4/16, L = 1, L > 64
    "tc_k32":   (96, 700, 32, 12, "planted", 30, True),
    "tc_k27":   (130, 1000, 27, 5, "planted", 30, True),
    "tc_l1":    (64, 512, 32, 1, "uniform", 20, True),
    "tc_l70":   (40, 600, 30, 70, "planted", 20, True),
    "mid":      (128, 2048, 8, 16, "planted", 100, False),
    # BASELINE config 2 at full size
    "B":        (256, 65536, 8, 32, "planted", 100, False),
    # configs 3..5 at reduced T (identical N, K, L), few iterations
    "C_small":  (1024, 4096, 32, 64, "planted", 3, False),
    "C_mid":    (1024, 8192, 32, 64, "planted", 25, False),
    "D_small":  (512, 2048, 16, 256, "planted", 3, False),
    "E_small":  (2048, 2048, 128, 16, "planted", 3, False),
}

# full-size BASELINE configs (bench / property tests; no CPU oracle at this size)
FULL = {
    "A": (100, 250, 3, 20),
    "B": (256, 65536, 8, 32),
    "C": (1024, 1 << 20, 32, 64),
    "D": (512, 1 << 18, 16, 256),
    "E": (2048, 1 << 22, 128, 16),
}


def _predict(W, H):
    """sum_l W[l] @ shift(H, l) in float32 (input synthesis only)."""
    L, N, K = W.shape
    T = H.shape[1]
    est = np.zeros((N, T), dtype=np.float32)
    for l in range(min(L, T)):
        est[:, l:] += W[l] @ H[:, :T - l]
    return est


def make_inputs(N, T, K, L, kind="planted", seed=0):
    """Returns (X, W0, H0) as float32 arrays, X >= 0.

    planted : X = recon(W*, H*) + 0.1 U[0,1), H* 90 % sparse, W* smooth bumps
              (the structure of reference datasets/synthetic.py:7-39).
    uniform : X ~ U[0,1).
    zeros_rows : planted, with some all-zero rows/columns in X and exact zeros
              in W0/H0 (zeros must stay zero under MU; 0/(0+eps) must be 0).
    Init follows reference algs/base.py:78-88: U[0,1) then the alpha rescale.
    """
    rng = np.random.default_rng(seed)
    f32 = np.float32
    if kind in ("planted", "zeros_rows"):
        lags = np.linspace(-3, 3, L, dtype=f32)[:, None, None]
        tau = rng.uniform(-1.5, 1.5, size=(1, N, 1)).astype(f32)
        comp = rng.integers(0, K, size=N)
        Wt = np.exp(-(lags - tau) ** 2).astype(f32) * (np.arange(K)[None, None, :] == comp[None, :, None])
        Ht = (rng.random((K, T), dtype=f32) * (rng.random((K, T), dtype=f32) < 0.1)).astype(f32)
        X = _predict(Wt.astype(f32), Ht) + f32(0.1) * rng.random((N, T), dtype=f32)
    elif kind == "uniform":
        X = rng.random((N, T), dtype=f32)
    else:
        raise ValueError(kind)
    W0 = rng.random((L, N, K), dtype=f32)
    H0 = rng.random((K, T), dtype=f32)
    if kind == "zeros_rows":
        X[::5, :] = 0
        X[:, 10:14] = 0
        W0[:, 1, :] = 0
        W0[0, :, 0] = 0
        H0[:, 20:25] = 0
    est = _predict(W0, H0).astype(np.float64)
    alpha = float((X.astype(np.float64) * est).sum() / (est ** 2).sum())
    s = f32(np.sqrt(alpha))
    return np.ascontiguousarray(X, f32), np.ascontiguousarray(W0 * s, f32), np.ascontiguousarray(H0 * s, f32)


def case_inputs(name):
    N, T, K, L, kind, n_iter, full = CASES[name]
    seed = sum(ord(c) for c in name)
    return make_inputs(N, T, K, L, kind, seed)
