"""Gradient solvers (reference cmfpy/algs/gradient_descent.py: GradDescent, BlockDescent) and HALS
(cmfpy/algs/hals.py on cmfpy/algs/accelerated.py).

CPU part (`-m "not gpu"`): the NumPy restatement in oracle/cmf_oracle.py against golden trajectories generated from
the reference itself (oracle/make_golden_gd.py), and the identity the device path rests on: the gradients are the
multiplicative-update terms, gW = den_W - num_W and gH = den_H - num_H.

GPU part (`-m gpu`): the device solvers, called through the C ABI, against the same goldens.  Tolerances: loss
trajectory <= 1e-4 relative (the parity bar of the MU path), Lipschitz constant <= 1e-5 relative (power iteration vs
LAPACK), gradients <= 2e-5 of their largest entry on the fp32 path.
"""
import numpy as np
import pytest

from oracle import cmf_oracle as o
from tests.cases import CASES, case_inputs
from tests.conftest import golden

GD_CASES = ["A", "odd_k5", "mid", "tc_k32"]
TRAJ_TOL = 1e-4


def _inputs(name):
    g = golden("gd_" + name)
    X, W0, H0 = case_inputs(name)
    return g, X, W0, H0


def _fit_loop(alg, n_iter):
    """The loop of CMF.fit (model.py:157-172): update, then converged() - which adapts the H step size."""
    hist, steps = [alg.loss], [alg.step_size]
    for _ in range(n_iter):
        hist.append(alg.update())
        alg.converged(hist)
        steps.append(alg.step_size)
    return np.array(hist), np.array(steps)


# ---------------------------------------------------------------- CPU: oracle vs reference goldens
@pytest.mark.parametrize("key", ["gd", "bcd"])
@pytest.mark.parametrize("name", GD_CASES)
def test_oracle_trajectory_against_reference_golden(name, key):
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    alg = o.GradDescentOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64), initH=H0.astype(np.float64),
                              tol=0, block=(key == "bcd"))
    assert abs(alg.lipschitz_W() - float(g[key + "_lipschitz0"])) <= 1e-10 * float(g[key + "_lipschitz0"])
    hist, steps = _fit_loop(alg, int(g["n_iter"]))
    assert np.abs(hist - g[key + "_loss_hist"]).max() <= 1e-12
    assert np.array_equal(steps, g[key + "_step_hist"])
    if key + "_W" in g.files:
        assert np.abs(alg.W - g[key + "_W"]).max() <= 1e-10
        assert np.abs(alg.gW - o.GradDescentOracle(X.astype(np.float64), L, K, initW=alg.W, initH=alg.H).gW).max() <= 1e-9


def test_gradients_are_the_mu_terms():
    """gW[l] = s_T_dot(resids, H, l) = den_W[l] - num_W[l] and gH = den_H - num_H (gradient_descent.py:40-52 vs
    mult.py:27-48): what lets the device solvers reuse the MU contraction kernels."""
    rng = np.random.default_rng(2)
    N, T, K, L = 11, 70, 3, 6
    X, W, H = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    alg = o.GradDescentOracle(X, L, K, initW=W, initH=H)
    numW, denW = o.w_terms(X, alg.est, H, L)
    numH, denH = o.h_terms(X, alg.est, W)
    np.testing.assert_allclose(alg.gW, denW - numW, atol=1e-11)
    np.testing.assert_allclose(alg.gH, denH - numH, atol=1e-11)


def test_step_size_adaptation():
    """converged() divides the H step by step_decrement when the loss went up and never reports convergence then
    (gradient_descent.py:94-113)."""
    rng = np.random.default_rng(0)
    alg = o.GradDescentOracle(rng.random((5, 40)), 3, 2, initW=rng.random((3, 5, 2)), initH=rng.random((2, 40)), tol=1.0)
    assert alg.converged([1.0, 0.9, 0.95]) is False and alg.step_size == 1e-4 / 5.0
    assert alg.converged([1.0, 0.9, 0.85]) is True and alg.step_size == 1e-4 / 5.0


HALS_VARIANTS = {"hals": {}, "hals_acc": dict(max_iter=3, weightW=1, weightH=2, stop_thresh=0.05)}


@pytest.mark.parametrize("key", list(HALS_VARIANTS))
@pytest.mark.parametrize("name", GD_CASES)
def test_hals_oracle_against_reference_golden(name, key):
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    alg = o.HALSOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64), initH=H0.astype(np.float64), tol=0,
                       **HALS_VARIANTS[key])
    ref = g[key + "_loss_hist"]
    n = len(ref) - 1 if X.size <= 40000 else 4                   # the NumPy sweeps are Python loops: keep the CPU suite short
    hist = [alg.loss] + [alg.update() for _ in range(n)]
    assert np.abs(np.array(hist) - ref[:n + 1]).max() <= 1e-12
    if key + "_W" in g.files and n == len(ref) - 1:
        assert np.abs(alg.W - g[key + "_W"]).max() <= 1e-10 and np.abs(alg.H - g[key + "_H"]).max() <= 1e-10


def test_hals_argument_check():
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError):                              # accelerated.py:47-48
        o.HALSOracle(rng.random((5, 40)), 3, 2, initW=rng.random((3, 5, 2)), initH=rng.random((2, 40)), max_iter=1, weightW=0.5)


# ---------------------------------------------------------------- GPU: device solvers vs reference goldens
def _device(cls_name, X, W0, H0, L, K, precision):
    from cmfpy_b200.algs import ALGORITHMS
    from cmfpy_b200.model import ModelDimensions
    return ALGORITHMS[cls_name](X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0,
                                precision=precision)


def _supported(precision, N, K, L):
    from cmfpy_b200 import _lib
    return bool(_lib.load().cmf_precision_supported(_lib.PRECISIONS[precision], N, K, L))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("key", ["gd", "bcd"])
@pytest.mark.parametrize("name", GD_CASES)
def test_device_trajectory_against_reference_golden(built_lib, name, key, precision):
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    alg = _device(key, X, W0, H0, L, K, precision)
    lam = alg.lipschitz_W()
    ref_lam = float(g[key + "_lipschitz0"])
    assert abs(lam - ref_lam) <= (1e-5 if precision != "tf32" else 2e-3) * ref_lam
    settled, iters = alg.lipschitz_state()
    assert settled and 64 <= iters <= 1024 and iters % 64 == 0
    if precision == "fp32" and key + "_gW0" in g.files:
        gW, gH = alg.gW, alg.gH
        assert np.abs(gW - g[key + "_gW0"]).max() <= 2e-5 * np.abs(g[key + "_gW0"]).max()
        assert np.abs(gH - g[key + "_gH0"]).max() <= 2e-5 * np.abs(g[key + "_gH0"]).max()
    hist, steps = _fit_loop(alg, int(g["n_iter"]))
    ref = g[key + "_loss_hist"]
    rel = np.abs(hist - ref) / ref
    print("%s/%s/%s: max rel loss err %.3e, lipschitz rel err %.2e" % (name, key, precision, rel.max(), abs(lam - ref_lam) / ref_lam))
    assert rel.max() <= (TRAJ_TOL if precision != "tf32" else 5e-3)
    assert np.array_equal(steps, g[key + "_step_hist"])
    if key + "_W" in g.files and precision == "fp32":
        assert np.abs(alg.W - g[key + "_W"]).max() <= 5e-4 * np.abs(g[key + "_W"]).max()
        assert np.abs(alg.H - g[key + "_H"]).max() <= 5e-4 * np.abs(g[key + "_H"]).max()
    alg.close()


@pytest.mark.gpu
def test_cmf_fit_with_gradient_solvers(built_lib):
    """CMF(alg_name='gd' | 'bcd') through the model API (model.py:146: the registry plug-in point)."""
    from cmfpy_b200 import CMF
    g, X, W0, H0 = _inputs("A")
    for key in ("gd", "bcd"):
        model = CMF(3, 20, n_iter_max=int(g["n_iter"]), alg_name=key, verbose=False, tol=0, initW=W0, initH=H0)
        model.fit(X)
        ref = g[key + "_loss_hist"]
        assert len(model.loss_hist) == len(ref)
        assert (np.abs(np.array(model.loss_hist) - ref) / ref).max() <= TRAJ_TOL
        assert (model.motifs >= 0).all() and (model.factors >= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("key", list(HALS_VARIANTS))
@pytest.mark.parametrize("name", GD_CASES)
def test_device_hals_against_reference_golden(built_lib, name, key, precision):
    """HALSUpdate on the device (residual kept current by csrc/hals_kernels.cuh) against the reference's own
    trajectories; `precision` only selects the kernel of the closing reconstruction."""
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    from cmfpy_b200.algs import ALGORITHMS
    from cmfpy_b200.model import ModelDimensions
    alg = ALGORITHMS["hals"](X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0, tol=0,
                             precision=precision, **HALS_VARIANTS[key])
    ref = g[key + "_loss_hist"]
    hist = np.array([alg.loss] + alg.update_many(len(ref) - 1))
    rel = np.abs(hist - ref) / ref
    print("%s/%s/%s: max rel loss err %.3e" % (name, key, precision, rel.max()))
    assert rel.max() <= TRAJ_TOL
    if key + "_W" in g.files:
        assert np.abs(alg.W - g[key + "_W"]).max() <= 2e-3 * np.abs(g[key + "_W"]).max()
        assert np.abs(alg.H - g[key + "_H"]).max() <= 2e-3 * np.abs(g[key + "_H"]).max()
    alg.close()


@pytest.mark.gpu
def test_cmf_fit_with_hals(built_lib):
    from cmfpy_b200 import CMF
    g, X, W0, H0 = _inputs("A")
    model = CMF(3, 20, n_iter_max=10, alg_name="hals", verbose=False, tol=0, initW=W0, initH=H0)
    model.fit(X)
    ref = g["hals_loss_hist"][:11]
    assert (np.abs(np.array(model.loss_hist) - ref) / ref).max() <= TRAJ_TOL
    with pytest.raises(ValueError):
        CMF(3, 20, alg_name="hals", verbose=False, max_iter=1, weightH=0.5, initW=W0, initH=H0).fit(X)
