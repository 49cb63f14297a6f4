"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on the seeded cases of
tests/cases.py.  Build-container only; the committed .npz files are what travel.

    python -m oracle.make_golden [case ...]        # default: every case

Each file holds: the float64 loss trajectory of reference MultUpdate
(`loss_hist`, tol=0), the single-step intermediates of iteration 1
(est0, numW, denW, W1, numH, denH, H1 - from the reference's own
_compute_mult_W/_compute_mult_H, algs/mult.py:27-48) for the small cases,
and the final W/H (full for small cases; sums + strided samples otherwise).
Inputs are not stored for the large cases: they are regenerated from the seed.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import import_reference  # noqa: E402
from tests.cases import CASES, case_inputs    # noqa: E402


def run_case(name, out_dir):
    import_reference()
    from cmfpy.algs.mult import MultUpdate
    from cmfpy.model import ModelDimensions

    N, T, K, L, kind, n_iter, full = CASES[name]
    if kind == "ref_synthetic":
        # BASELINE config 1 through the reference's own generator
        # (datasets/synthetic.py:7-39; it also draws from the global RNG).
        from cmfpy.datasets.synthetic import Synthetic
        from tests.cases import make_inputs
        np.random.seed(0)
        X32 = Synthetic(n_components=K, n_features=N, n_lags=L, n_timebins=T,
                        seed=0).generate().astype(np.float32)
        _, W32, H32 = make_inputs(N, T, K, L, "uniform", seed=7)
        est = __import__("cmfpy").common.cmf_predict(W32.astype(np.float64), H32.astype(np.float64))
        s = np.float32(np.sqrt((X32 * est).sum() / (est ** 2).sum()))
        W32, H32 = W32 * s, H32 * s
    else:
        X32, W32, H32 = case_inputs(name)
    X, W0, H0 = (a.astype(np.float64) for a in (X32, W32, H32))
    dims = ModelDimensions(X, maxlag=L, n_components=K)
    out = {"shape": np.array([N, T, K, L]), "n_iter": np.array(n_iter)}

    if full:   # single-step intermediates straight from the reference's methods
        a0 = MultUpdate(X, dims, initW=W0.copy(), initH=H0.copy(), tol=0)
        out["est0"] = a0.est.copy()
        numW, denW = a0._compute_mult_W()
        a0.update()
        a1 = MultUpdate(X, dims, initW=a0.W.copy(), initH=H0.copy(), tol=0)
        numH, denH = a1._compute_mult_H()
        out.update(numW=numW, denW=denW, W1=a0.W.copy(), numH=numH, denH=denH,
                   H1=a0.H.copy(), X=X32, W0=W32, H0=H32)

    alg = MultUpdate(X, dims, initW=W0.copy(), initH=H0.copy(), tol=0)
    hist = [alg.loss]
    t0 = time.time()
    for i in range(n_iter):
        hist.append(alg.update())
    dt = (time.time() - t0) / n_iter
    out["loss_hist"] = np.array(hist)
    out["ref_seconds_per_iter"] = np.array(dt)
    if full:
        out["W_final"], out["H_final"] = alg.W, alg.H
    else:
        out["W_sum"], out["H_sum"] = alg.W.sum(), alg.H.sum()
        out["W_sample"] = alg.W.reshape(-1)[::997].copy()
        out["H_sample"] = alg.H.reshape(-1)[::997].copy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    print("%-8s N=%d T=%d K=%d L=%d  %d it  %.4f s/it  loss %.6f -> %.6f" %
          (name, N, T, K, L, n_iter, dt, hist[0], hist[-1]), flush=True)


if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name, out_dir)
