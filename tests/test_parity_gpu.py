"""GPU parity tests: the CUDA path, called through the C ABI, against the golden
fixtures generated from the unmodified reference and against the CPU oracle on
the same seeded inputs.

Tolerances (stated once, used everywhere):
  * loss trajectory: max_i |loss_gpu[i] - loss_ref[i]| / loss_ref[i] <= 1e-4
    over all recorded iterations (BASELINE.json north_star), float64 reference;
  * single contraction, fp32 path: rtol 2e-5 of the largest entry;
  * single contraction, tf32 path: rtol 2e-3 of the largest entry (10-bit
    mantissa operands, fp32 accumulation);
  * loss trajectory, tf32 path: reported per case, asserted <= 5e-3.  A fixed
    1e-4 perturbation of X alone moves the fast-converging planted trajectories
    by 5e-4 at a fixed iteration index (CPU experiment in DESIGN.md), so no
    10-bit-mantissa path can meet 1e-4 there; the fp32 path does.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import cmf_oracle as o
from tests.cases import CASES, case_inputs, make_inputs
from tests.conftest import golden

pytestmark = pytest.mark.gpu

TRAJ_TOL = 1e-4          # fp32 path (the parity bar)
TRAJ_TOL_TF32 = 5e-3     # tf32 path: reported, no bar in the north star (DESIGN.md section 6 lists
                         # the measured values: 4e-8 .. 2.3e-3, worst on config B at iteration ~60)
# tf32x3 (three TF32 MMAs per product, two-level accumulation: short tensor-memory sub-chunks folded in fp32
# registers) is held to the same 1e-4 bar on EVERY golden case, config B included, with no exception; measured
# 8e-10 .. 8e-6, config B 1.1e-5 (direct) / 3.8e-5 (Gram): profiles/r02_trajectory_errors.log.
FULL_CASES = [n for n, c in CASES.items() if c[6]]
ALL_CASES = list(CASES)


PRECISION_MODES = ["fp32", "tf32", "tf32g", "tf32x3", "tf32x3g"]   # ...g = Gram-route denominators;
                                                        # tf32x3 = error-compensated tensor-core path (fp32-grade)
EXACT_MODES = ("fp32", "tf32x3", "tf32x3g")             # modes held to the 1e-4 parity bar


def _split(precision):
    return (precision[:-1], "gram") if precision.endswith("g") else (precision, "direct")


def _supported(precision, N, K, L):
    from cmfpy_b200 import _lib
    return bool(_lib.load().cmf_precision_supported(_lib.PRECISIONS[_split(precision)[0]], N, K, L))


def _inputs(name):
    g = golden(name)
    if "X" in g.files:
        return g, g["X"], g["W0"], g["H0"]
    X, W0, H0 = case_inputs(name)
    return g, X, W0, H0


def _solver(X, W0, H0, L, K, precision):
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    prec, den = _split(precision)
    return MultUpdate(X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0, initH=H0,
                      tol=0, precision=prec, denominators=den)


def _close(a, b, rel):
    scale = np.abs(b).max() + 1e-30
    assert np.abs(a - b).max() <= rel * scale, "max abs err %.3e vs scale %.3e" % (np.abs(a - b).max(), scale)


@pytest.mark.parametrize("precision", PRECISION_MODES)
@pytest.mark.parametrize("name", FULL_CASES)
def test_single_step_kernels(built_lib, name, precision):
    """est, W terms, W update, H terms, H update of iteration 1, one kernel at a
    time, against the reference's own intermediates."""
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    rel = 2e-5 if precision in EXACT_MODES else 2e-3
    alg = _solver(X, W0, H0, L, K, precision)
    if precision == "tf32g":
        assert alg.path_name == "tcgen05-tf32+gram"
    if precision == "tf32x3":
        assert alg.path_name == "tcgen05-tf32x3"
    if precision == "tf32x3g":
        assert alg.path_name == "tcgen05-tf32x3+gram"
    _close(alg.est, g["est0"], rel)
    numW, denW = alg._compute_mult_W()
    _close(numW, g["numW"], rel)
    _close(denW, g["denW"], rel)
    alg2 = _solver(X, g["W1"].astype(np.float32), H0, L, K, precision)
    numH, denH = alg2._compute_mult_H()
    _close(numH, g["numH"], rel)
    _close(denH, g["denH"], rel)
    loss1 = alg.update()
    assert abs(loss1 - g["loss_hist"][1]) / g["loss_hist"][1] < TRAJ_TOL
    _close(alg.W, g["W1"], 10 * rel)
    _close(alg.H, g["H1"], 10 * rel)
    # zeros stay exactly zero under MU (0 * num / (den + eps) == 0)
    assert np.all(alg.W[g["W1"] == 0] == 0)
    assert np.all(alg.H[g["H1"] == 0] == 0)
    assert (alg.W >= 0).all() and (alg.H >= 0).all()
    alg.close(); alg2.close()


@pytest.mark.parametrize("precision", PRECISION_MODES)
@pytest.mark.parametrize("name", ALL_CASES)
def test_loss_trajectory(built_lib, name, precision):
    g, X, W0, H0 = _inputs(name)
    N, T, K, L = (int(v) for v in g["shape"])
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    n_iter = int(g["n_iter"])
    alg = _solver(X, W0, H0, L, K, precision)
    hist = np.array([alg.loss] + alg.update_many(n_iter))
    ref = g["loss_hist"]
    rel = np.abs(hist - ref) / ref
    print("%s/%s: max rel loss err %.3e (final %.6f vs %.6f)" % (name, precision, rel.max(), hist[-1], ref[-1]))
    tol = TRAJ_TOL if precision in EXACT_MODES else TRAJ_TOL_TF32
    assert rel.max() <= tol
    Wg, Hg = alg.W, alg.H
    if "W_final" in g.files:
        ftol = 5e-3 if precision in EXACT_MODES else 5e-2
        _close(Wg, g["W_final"], ftol)
        _close(Hg, g["H_final"], ftol)
    else:
        assert abs(Wg.sum() - g["W_sum"]) / g["W_sum"] < 1e-3
        assert abs(Hg.sum() - g["H_sum"]) / g["H_sum"] < 1e-3
    alg.close()


@pytest.mark.parametrize("precision", PRECISION_MODES)
@pytest.mark.parametrize("shape", [(3, 5, 2, 8), (1, 1, 1, 1), (7, 40, 100, 3), (300, 600, 64, 5), (5, 9, 3, 9),
                                   (130, 257, 17, 31), (64, 3000, 32, 300)])    # the last: lag range > one K1 window
def test_extreme_shapes_against_oracle(built_lib, shape, precision):
    """L > T, single entries, K padded to 128 / 64 / 32, ragged everything: three MU iterations against
    the float64 oracle on the same inputs."""
    N, T, K, L = shape
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    X, W0, H0 = make_inputs(N, T, K, L, "uniform", seed=N + T + K + L)
    ref = o.MultUpdateOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64), initH=H0.astype(np.float64), tol=0)
    ref_hist = [ref.loss] + [ref.update() for _ in range(3)]
    alg = _solver(X, W0, H0, L, K, precision)
    hist = [alg.loss] + alg.update_many(3)
    tol = 1e-5 if precision in EXACT_MODES else 2e-3
    for a, b in zip(hist, ref_hist):
        assert abs(a - b) <= tol * max(b, 0.5), (hist, ref_hist)   # (an exactly-fittable 1x1 problem has loss ~ 0)
    wh_tol = 20 * tol
    _close(alg.W, ref.W, wh_tol)
    _close(alg.H, ref.H, wh_tol)
    alg.close()


def test_update_one_by_one_equals_batched(built_lib):
    g, X, W0, H0 = _inputs("odd_k5")
    N, T, K, L = (int(v) for v in g["shape"])
    a = _solver(X, W0, H0, L, K, "fp32")
    b = _solver(X, W0, H0, L, K, "fp32")
    la = [a.update() for _ in range(6)]
    lb = b.update_many(6)
    assert la == lb                      # same kernels, same order: bit-identical
    assert np.array_equal(a.W, b.W) and np.array_equal(a.H, b.H)
    a.close(); b.close()


def test_caller_arrays_are_not_mutated(built_lib):
    """MU rebinds W/H (reference mult.py:18,22); initW/initH/data stay untouched."""
    g, X, W0, H0 = _inputs("k1")
    N, T, K, L = (int(v) for v in g["shape"])
    Xc, Wc, Hc = X.copy(), W0.copy(), H0.copy()
    alg = _solver(X, W0, H0, L, K, "fp32")
    alg.update_many(3)
    assert np.array_equal(X, Xc) and np.array_equal(W0, Wc) and np.array_equal(H0, Hc)
    alg.close()


# ---- the reference's own unit tests, run through the GPU primitives -------
def test_reference_sdot_vectors_on_gpu(built_lib):
    from cmfpy_b200.common import s_dot
    from tests.test_oracle import OV, B_SDOT, SDOT_EXPECT
    for s, exp in SDOT_EXPECT.items():
        assert_allclose(s_dot(OV, B_SDOT, s), [exp], rtol=1e-6)


def test_reference_sTdot_vectors_on_gpu(built_lib):
    from cmfpy_b200.common import s_T_dot
    from tests.test_oracle import OV, B_STDOT, STDOT_EXPECT
    for s, exp in STDOT_EXPECT.items():
        assert_allclose(s_T_dot(OV, B_STDOT, s), [exp], rtol=1e-6)


def test_primitives_against_oracle(built_lib):
    from cmfpy_b200.common import cmf_predict, tensor_transconv
    rng = np.random.default_rng(3)
    for (N, T, K, L) in [(5, 33, 2, 4), (130, 700, 9, 17), (64, 257, 16, 1)]:
        W, H, X = rng.random((L, N, K)), rng.random((K, T)), rng.random((N, T))
        _close(cmf_predict(W, H), o.cmf_predict(W, H), 1e-5)
        _close(tensor_transconv(W, X), o.tensor_transconv(W, X), 1e-5)


def test_score_on_device_against_oracle(built_lib):
    """CMF.score (reference model.py:202-221) without reading est back."""
    from cmfpy_b200.common import score
    rng = np.random.default_rng(11)
    for (N, T, K, L) in [(7, 90, 2, 5), (130, 1500, 9, 17), (256, 4096, 32, 64)]:
        W, H, X = rng.random((L, N, K)), rng.random((K, T)), rng.random((N, T)) * L * K / 4
        ref = 1 - np.linalg.norm(o.cmf_predict(W, H) - X) ** 2 / np.linalg.norm(X) ** 2
        assert abs(score(W, H, X) - ref) <= 1e-5 * max(1.0, abs(ref))
        if _supported("tf32x3", N, K, L):
            # tensor memory accumulates with round-toward-zero: est carries a relative bias of about
            # -6e-8 per MMA step (256 steps at K=32, L=64 -> 1.5e-5), which this statistic sees directly
            assert abs(score(W, H, X, precision="tf32x3") - ref) <= 1e-4 * max(1.0, abs(ref))


def test_loadings_sort_and_renormalize(built_lib):
    """compute_loadings / sort_components / renormalize (reference model.py:178-189, 278-311)."""
    from cmfpy_b200 import CMF
    from cmfpy_b200.model import compute_loadings, renormalize
    rng = np.random.default_rng(5)
    N, T, K, L = 40, 300, 4, 7
    W, H = rng.random((L, N, K)), rng.random((K, T))
    W[:, :, 2] *= 3.0                                   # one dominant component
    X = o.cmf_predict(W, H)
    ref = [np.linalg.norm(o.cmf_predict(W[:, :, k:k + 1], H[k:k + 1]) - X) / (np.linalg.norm(X) + o.EPSILON) for k in range(K)]
    got = compute_loadings(X, W, H)
    assert_allclose(got, ref, rtol=1e-5)
    model = CMF(K, L)
    model._W, model._H = W.copy(), H.copy()
    ind = model.sort_components(X)
    assert list(ind) == list(np.argsort(ref)) and ind[0] == 2
    assert np.array_equal(model.motifs, W[:, :, ind]) and np.array_equal(model.factors, H[ind])
    W2, H2 = renormalize(W, H)
    assert_allclose(np.linalg.norm(H2, axis=1), 1.0, rtol=1e-12)
    assert_allclose(o.cmf_predict(W2, H2), X, rtol=1e-10)


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("mode", ["l2", "l1", "std"])
def test_device_normalisation(built_lib, mode, precision):
    """normalize= scales the rows on the device (cmf_mu_row_stats / cmf_mu_scale_rows) exactly as the reference's
    dataset classes do on the host, and the solve then equals the solve on the pre-normalised matrix."""
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.common import row_scales
    from cmfpy_b200.model import ModelDimensions
    N, T, K, L = 37, 301, 5, 7
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=21)
    X = X * np.linspace(0.5, 20.0, N, dtype=np.float32)[:, None]
    X64 = X.astype(np.float64)
    Xn = X64 * row_scales(mode, X64.sum(1), (X64 ** 2).sum(1), np.abs(X64).sum(1), T)[:, None]
    dims = ModelDimensions(X, maxlag=L, n_components=K)
    a = MultUpdate(X, dims, initW=W0, initH=H0, tol=0, precision=precision, normalize=mode)
    assert abs(a.normX - np.linalg.norm(Xn)) <= 1e-5 * np.linalg.norm(Xn)
    ref = o.MultUpdateOracle(Xn, L, K, initW=W0.astype(np.float64), initH=H0.astype(np.float64), tol=0)
    hist, ref_hist = [a.loss] + a.update_many(5), [ref.loss] + [ref.update() for _ in range(5)]
    assert np.abs(np.array(hist) - ref_hist).max() <= 2e-5 * max(ref_hist)
    a.close()


# ---- model API --------------------------------------------------------------
def test_cmf_fit_predict_score(built_lib):
    from cmfpy_b200 import CMF
    g, X, W0, H0 = _inputs("A")
    model = CMF(3, 20, n_iter_max=100, verbose=False, tol=0, initW=W0, initH=H0)
    assert model.fit(X) is None                      # reference fit returns None
    assert len(model.loss_hist) == 101 and len(model.time_hist) == 101
    assert model.time_hist[0] == 0.0 and np.all(np.diff(model.time_hist) > 0)
    rel = np.abs(np.array(model.loss_hist) - g["loss_hist"]) / g["loss_hist"]
    assert rel.max() <= TRAJ_TOL
    assert model.motifs.shape == (20, 100, 3) and model.factors.shape == (3, 250)
    assert model.n_features == 100 and model.n_timesteps == 250
    est = model.predict()
    assert est.shape == (100, 250)
    _close(est, o.cmf_predict(model.motifs, model.factors), 2e-6)
    r2 = model.score(X)
    assert abs((1 - r2) - model.loss_hist[-1] ** 2) < 1e-5
    assert sorted(model.argsort_units()) == list(range(100))


def test_cmf_early_stopping_and_random_init(built_lib):
    from cmfpy_b200 import CMF
    X, _, _ = make_inputs(20, 120, 3, 5, "planted", seed=9)
    model = CMF(3, 5, n_iter_max=100000, verbose=False, tol=1e-3, seed=0)
    model.fit(X)
    assert 3 < len(model.loss_hist) < 2000           # converged() fired
    d = np.abs(np.diff(model.loss_hist[-3:]))
    assert np.all(d < 1e-3)
    # the alpha rescale of rand_init makes <X, est0> = ||est0||^2
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    alg = MultUpdate(X, ModelDimensions(X, maxlag=5, n_components=3), seed=1)
    est = alg.est
    assert abs((X * est).sum() / (est ** 2).sum() - 1.0) < 1e-4
    alg.close()


def test_errors_match_reference(built_lib):
    from cmfpy_b200 import CMF
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    X = np.random.default_rng(0).random((6, 50)).astype(np.float32)
    with pytest.raises(ValueError):
        CMF(2, 3, verbose=False).fit(-X)
    with pytest.raises(ValueError):
        MultUpdate(X, ModelDimensions(X, maxlag=3, n_components=2), patience=0)
    with pytest.raises(ValueError):
        MultUpdate(X, ModelDimensions(X, maxlag=3, n_components=2), patience=2.5)
    with pytest.raises(ValueError):
        MultUpdate(X, ModelDimensions(X, maxlag=3, n_components=2), initW=np.ones((3, 6, 3)), initH=np.ones((2, 50)))


# ---- size-independent properties at a larger size -----------------------------
@pytest.mark.parametrize("precision", PRECISION_MODES)
def test_properties_large(built_lib, precision):
    N, T, K, L = 1024, 1 << 15, 32, 64        # config C at T/32
    if not _supported(precision, N, K, L):
        pytest.skip("no %s kernel for this shape" % precision)
    X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=42)
    alg = _solver(X, W0, H0, L, K, precision)
    l0 = alg.loss
    hist = alg.update_many(4)
    assert np.all(np.diff([l0] + hist) < 0), "MU must decrease the loss here"
    W, H = alg.W, alg.H
    assert np.isfinite(W).all() and np.isfinite(H).all() and (W >= 0).all() and (H >= 0).all()
    # linearity of the reconstruction: recon(2W, H) = 2 recon(W, H)
    from cmfpy_b200.common import cmf_predict
    Ws, Hs = W[:, :64].astype(np.float32), H[:, :4096].astype(np.float32)
    e1, e2 = cmf_predict(Ws, Hs, precision=_split(precision)[0]), cmf_predict(2 * Ws, Hs, precision=_split(precision)[0])
    assert_allclose(e2, 2 * e1, rtol=1e-6, atol=1e-6)
    # the loss the solver reports equals the loss of what it returns
    est = alg.est
    loss = np.linalg.norm(est - X) / np.linalg.norm(X)
    assert abs(loss - hist[-1]) / hist[-1] < 1e-4
    alg.close()


@pytest.mark.parametrize("precision", ["tf32", "tf32g", "tf32x3", "tf32x3g"])
def test_fp32_and_tf32_agree_large(built_lib, precision):
    N, T, K, L = 512, 1 << 14, 16, 32
    if not _supported("tf32", N, K, L):
        pytest.skip("no tf32 kernel for this shape")
    X, W0, H0 = make_inputs(N, T, K, L, "planted", seed=17)
    a, b = _solver(X, W0, H0, L, K, "fp32"), _solver(X, W0, H0, L, K, precision)
    ha, hb = np.array(a.update_many(10)), np.array(b.update_many(10))
    assert (np.abs(ha - hb) / ha).max() < TRAJ_TOL
    # the scale split between W and H must not drift apart either
    assert abs(a.W.sum() / b.W.sum() - 1) < 1e-3 and abs(a.H.sum() / b.H.sum() - 1) < 1e-3
    a.close(); b.close()


@pytest.mark.parametrize("kind", ["uniform", "planted"])
def test_one_pass_loss(built_lib, kind):
    """tf32x3 on a large problem computes the loss from a one-pass reconstruction (loss_precision="auto",
    cmf_mu_set_loss_mode); it must agree with the three-pass loss to 1e-6 relative, and W, H must be identical."""
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    N, T, K, L = 1024, 1 << 16, 32, 64             # L N K = 2^21, K T = 2^21 factor entries
    X, W0, H0 = make_inputs(N, T, K, L, kind, seed=17)
    dims = ModelDimensions(X, maxlag=L, n_components=K)
    out = {}
    for mode in ("auto", "full"):
        alg = MultUpdate(X, dims, initW=W0, initH=H0, tol=0, precision="tf32x3", denominators="gram", loss_precision=mode)
        l0 = alg.loss                                 # the host has seen a loss: the next batch may go one-pass
        hist = alg.update_many(3) + alg.update_many(3)
        out[mode] = (np.array([l0] + hist), alg.W, alg.H, alg.launch_table() if False else None)
        alg.close()
    a, f = out["auto"][0], out["full"][0]
    print("one-pass loss vs full (%s): max rel diff %.2e" % (kind, (np.abs(a - f) / f).max()))
    assert (np.abs(a - f) / f).max() <= 1e-6
    assert np.array_equal(out["auto"][1], out["full"][1]) and np.array_equal(out["auto"][2], out["full"][2])


@pytest.mark.parametrize("precision,tol", [("tf32x3", 1e-5), ("tf32", 1e-3)])
@pytest.mark.parametrize("kind", ["uniform", "planted"])
def test_loss_from_w_terms(built_lib, precision, kind, tol):
    """Full Gram route, loss_precision="auto": while loss^2 >= 0.1 the loss comes from the W terms of the updated
    factors (||X||^2 - 2 <W, num_W> + <W, den_W>, cmf_abi.cu decide_loss_mode) and no reconstruction runs; below that
    the residual is reconstructed again.  The losses must agree with loss_precision="full" (every iteration's
    residual formed explicitly, all operand passes) - 1e-5 relative for the fp32-grade mode, a tenth of the parity
    bar - and W, H must be bit-identical: the loss evaluation never feeds the factors.  'uniform' keeps the loss
    near 0.5 (identity throughout), 'planted' falls through the threshold (the switch back is exercised)."""
    from cmfpy_b200.algs.mult import MultUpdate
    from cmfpy_b200.model import ModelDimensions
    N, T, K, L = 256, 1 << 14, 8, 32
    X, W0, H0 = make_inputs(N, T, K, L, kind, seed=23)
    dims = ModelDimensions(X, maxlag=L, n_components=K)
    out = {}
    for mode in ("auto", "full"):
        # (plain TF32 never takes the identity on its own: its W terms are too coarse to amplify; asked for explicitly)
        lp = "wterms" if (mode == "auto" and precision == "tf32") else mode
        alg = MultUpdate(X, dims, initW=W0, initH=H0, tol=0, precision=precision, denominators="gram", loss_precision=lp)
        assert alg.path_name.endswith("+gram")
        l0 = alg.loss
        hist = alg.update_many(5) + [alg.update() for _ in range(3)] + alg.update_many(40)
        alg.set_profiling(2)
        hist += alg.update_many(2)
        table = alg.launch_table()
        alg.set_profiling(0)
        out[mode] = (np.array([l0] + hist), alg.W, alg.H, table)
        alg.close()
    a, f = out["auto"][0], out["full"][0]
    rel = np.abs(a - f) / f
    print("loss from W terms vs residual (%s, %s): loss %.3f -> %.3f, max rel diff %.2e" % (precision, kind, f[0], f[-1], rel.max()))
    assert rel.max() <= tol
    assert np.array_equal(out["auto"][1], out["full"][1]) and np.array_equal(out["auto"][2], out["full"][2])
    assert "loss_identity" not in out["full"][3]
    if f[-3] ** 2 >= 0.1 or precision == "tf32":
        assert "loss_identity" in out["auto"][3] and not any(k.startswith("tc_recon") for k in out["auto"][3])
    else:
        assert "loss_identity" not in out["auto"][3]
    if kind == "uniform":
        assert f[-1] ** 2 >= 0.1              # (the identity was active from the second batch on)


def test_device_buffer_cache(built_lib):
    """The device buffers of a closed solver (1 MiB and more) are reused by the next solver that asks for the same
    sizes (csrc/dev_cache.cuh): results must not depend on what the previous owner left in them, and
    release_cached_memory() must leave the library usable."""
    import cmfpy_b200
    N, T, K, L = 512, 1 << 16, 8, 16                  # 128 MiB per N x T buffer
    hists = []
    for seed, release in ((1, False), (2, False), (1, True), (1, False)):
        X, W0, H0 = make_inputs(N, T, K, L, "uniform", seed=seed)
        alg = _solver(X, W0, H0, L, K, "tf32x3")
        hists.append((seed, [alg.loss] + alg.update_many(3)))
        alg.close()
        if release:
            cmfpy_b200.release_cached_memory()
    same = [h for s, h in hists if s == 1]
    assert same[0] == same[1] == same[2]              # bit-identical whatever the buffers held before
    assert hists[1][1] != same[0]
