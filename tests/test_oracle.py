"""CPU tests of the oracle (oracle/cmf_oracle.py): against the reference's own
known-answer vectors, against the committed golden fixtures generated from the
unmodified reference, and - when /root/reference is present (build container
only) - against the reference itself."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import cmf_oracle as o
from oracle import ref_shim
from tests.cases import CASES, case_inputs
from tests.conftest import golden

# ---- the reference's known-answer vectors (reference tests/test_numeric.py:15-55)
OV = np.array([[1., 1., 1.]])
B_SDOT = np.array([[0., 1., 2., 3.], [4., 5., 6., 7.], [8., 9., 10., 11.]])
SDOT_EXPECT = {2: [0., 0., 12., 15.], 1: [0., 12., 15., 18.], 0: [12., 15., 18., 21.],
               -1: [15., 18., 21., 0.], -2: [18., 21., 0., 0.]}
B_STDOT = np.array([[0., 1., 2.], [3., 4., 5.], [6., 7., 8.], [9., 10., 11.]])
STDOT_EXPECT = {2: [0., 3., 6., 9.], 1: [1., 7., 13., 19.], 0: [3., 12., 21., 30.],
                -1: [3., 9., 15., 21.], -2: [2., 5., 8., 11.]}


@pytest.mark.parametrize("shift", sorted(SDOT_EXPECT))
def test_sdot_known_answers(shift):
    assert_allclose(o.s_dot(OV, B_SDOT, shift), [SDOT_EXPECT[shift]])


@pytest.mark.parametrize("shift", sorted(STDOT_EXPECT))
def test_sTdot_known_answers(shift):
    assert_allclose(o.s_T_dot(OV, B_STDOT, shift), [STDOT_EXPECT[shift]])


def test_stacked_identity():
    rng = np.random.default_rng(0)
    W, H = rng.random((6, 9, 3)), rng.random((3, 40))
    assert_allclose(o.cmf_predict(W, H), o.cmf_predict_stacked(W, H), rtol=1e-12)
    X = rng.random((9, 40))
    num, _ = o.w_terms(X, X, H, 6)
    for l in range(6):
        assert_allclose(num[l], o.s_T_dot(X, H, l), rtol=1e-12)


FULL_CASES = [n for n, c in CASES.items() if c[6]]
ALL_FAST = [n for n, c in CASES.items() if n not in ("B", "C_small", "D_small", "E_small")]


@pytest.mark.parametrize("name", FULL_CASES)
def test_single_step_against_reference_golden(name):
    g = golden(name)
    N, T, K, L = (int(v) for v in g["shape"])
    X, W0, H0 = (g[k].astype(np.float64) for k in ("X", "W0", "H0"))
    est0 = o.cmf_predict(W0, H0)
    assert_allclose(est0, g["est0"], rtol=1e-11, atol=1e-13)
    numW, denW = o.w_terms(X, est0, H0, L)
    assert_allclose(numW, g["numW"], rtol=1e-11, atol=1e-12)
    assert_allclose(denW, g["denW"], rtol=1e-11, atol=1e-12)
    W1 = W0 * numW / (denW + o.EPSILON)
    assert_allclose(W1, g["W1"], rtol=1e-10, atol=1e-13)
    numH, denH = o.h_terms(X, o.cmf_predict(W1, H0), W1)
    assert_allclose(numH, g["numH"], rtol=1e-10, atol=1e-12)
    assert_allclose(denH, g["denH"], rtol=1e-10, atol=1e-12)
    assert_allclose(H0 * numH / (denH + o.EPSILON), g["H1"], rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("name", ALL_FAST)
def test_trajectory_against_reference_golden(name):
    g = golden(name)
    N, T, K, L = (int(v) for v in g["shape"])
    if "X" in g.files:
        X, W0, H0 = g["X"], g["W0"], g["H0"]
    else:
        X, W0, H0 = case_inputs(name)
    n_iter = int(g["n_iter"])
    W, H, hist = o.fit(X.astype(np.float64), L, K, n_iter_max=n_iter, initW=W0.astype(np.float64),
                       initH=H0.astype(np.float64), tol=0, reuse_est=True)
    assert len(hist) == n_iter + 1
    assert_allclose(hist, g["loss_hist"], rtol=1e-9)
    if "W_final" in g.files:
        assert_allclose(W, g["W_final"], rtol=1e-7, atol=1e-12)
        assert_allclose(H, g["H_final"], rtol=1e-7, atol=1e-12)
    else:
        assert_allclose(W.sum(), g["W_sum"], rtol=1e-9)
        assert_allclose(H.reshape(-1)[::997], g["H_sample"], rtol=1e-7, atol=1e-12)


def test_config_B_golden_is_present_and_sane():
    g = golden("B")
    hist = g["loss_hist"]
    assert hist.shape == (101,) and np.all(np.diff(hist) <= 1e-12)


def test_float32_oracle_is_within_bar():
    """Calibration: the same algorithm in fp32 stays within the 1e-4 bar."""
    g = golden("A")
    X, W0, H0 = g["X"], g["W0"], g["H0"]
    _, _, hist = o.fit(X, 20, 3, n_iter_max=100, initW=W0, initH=H0, tol=0, dtype=np.float32)
    rel = np.abs(np.array(hist) - g["loss_hist"]) / g["loss_hist"]
    assert rel.max() < 1e-5


@pytest.mark.parametrize("n_shards", [2, 4])
def test_sharded_restatement_matches(n_shards):
    rng = np.random.default_rng(5)
    N, T, K, L = 12, 96, 3, 7
    X, W, H = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    alg = o.MultUpdateOracle(X, L, K, initW=W, initH=H)
    loss = alg.update()
    Wn, Hn, ls = o.sharded_update(X, W, H, n_shards)
    assert_allclose(Wn, alg.W, rtol=1e-12)
    assert_allclose(Hn, alg.H, rtol=1e-12)
    assert_allclose(ls, loss, rtol=1e-12)


def test_fit_bookkeeping_and_errors():
    rng = np.random.default_rng(2)
    X = rng.random((8, 60))
    with pytest.raises(ValueError):
        o.fit(-X, 3, 2)
    with pytest.raises(ValueError):
        o.MultUpdateOracle(X, 3, 2, patience=0)
    W, H, hist = o.fit(X, 3, 2, n_iter_max=100000, tol=1e-3, rng=np.random.default_rng(0))
    assert 2 < len(hist) < 1000          # early stop through converged()
    assert W.shape == (3, 8, 2) and H.shape == (2, 60)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout only exists in the build container")
def test_oracle_equals_unmodified_reference():
    ref_shim.import_reference()
    from cmfpy.algs.mult import MultUpdate
    from cmfpy.model import ModelDimensions
    from cmfpy import common as rc
    rng = np.random.default_rng(11)
    N, T, K, L = 14, 90, 3, 6
    X, W0, H0 = rng.random((N, T)), rng.random((L, N, K)), rng.random((K, T))
    assert_allclose(o.cmf_predict(W0, H0), rc.cmf_predict(W0, H0), rtol=1e-13)
    assert_allclose(o.tensor_transconv(W0, X), rc.tensor_transconv(W0, X), rtol=1e-13)
    assert_allclose(o.shift_and_stack(H0, L), rc.shift_and_stack(H0, L))
    for s in (-3, -1, 0, 2, 5):
        assert_allclose(o.s_dot(W0[0], H0, s), rc.s_dot(W0[0], H0, s), rtol=1e-13)
        assert_allclose(o.s_T_dot(X, H0, s), rc.s_T_dot(X, H0, s), rtol=1e-13)
    ref = MultUpdate(X, ModelDimensions(X, maxlag=L, n_components=K), initW=W0.copy(), initH=H0.copy(), tol=0)
    orc = o.MultUpdateOracle(X, L, K, initW=W0, initH=H0, tol=0)
    assert ref.loss == pytest.approx(orc.loss, rel=1e-14)
    for _ in range(10):
        assert ref.update() == pytest.approx(orc.update(), rel=1e-13)
    assert_allclose(orc.W, ref.W, rtol=1e-11)
    assert_allclose(orc.H, ref.H, rtol=1e-11)


def test_row_scales_match_the_reference_normalisations():
    """cmfpy_b200.common.row_scales against the expressions of the reference's dataset classes."""
    from cmfpy_b200.common import row_scales
    rng = np.random.default_rng(4)
    X = rng.random((7, 50)) * rng.random((7, 1)) * 10
    X[3] = 2.5                                                   # a constant feature: zero variance
    s1, s2, sa = X.sum(1), (X ** 2).sum(1), np.abs(X).sum(1)
    np.testing.assert_allclose(X * row_scales("l2", s1, s2, sa, 50)[:, None],
                               X / (1e-6 + np.linalg.norm(X, axis=1, keepdims=True)), rtol=1e-12)    # songbird.py:18-19
    np.testing.assert_allclose(X * row_scales("l1", s1, s2, sa, 50)[:, None],
                               X / (1e-8 + np.linalg.norm(X, ord=1, axis=1, keepdims=True)), rtol=1e-12)  # maze.py:71-72
    from sklearn import preprocessing                             # vox_celeb.py:100-102
    ref = preprocessing.StandardScaler(with_mean=False).fit_transform(X.T).T
    np.testing.assert_allclose(X * row_scales("std", s1, s2, sa, 50)[:, None], ref, rtol=1e-9)


def test_model_checkpoint_round_trip(tmp_path):
    """save_model / load_model (.npz with cmf.jl's dataset names) and the h5py-gated cmf.jl loader."""
    from cmfpy_b200.model import CMF, load_cmfjl_model, load_model, save_model
    rng = np.random.default_rng(1)
    m = CMF(3, 4)
    m._W, m._H = rng.random((4, 6, 3)), rng.random((3, 20))
    m.loss_hist, m.time_hist = [0.9, 0.5, 0.4], [0.0, 0.1, 0.2]
    X = rng.random((6, 20))
    path = str(tmp_path / "ckpt.npz")
    save_model(path, m, X)
    data, m2 = load_model(path, verbose=False)
    assert np.array_equal(data, X) and np.array_equal(m2.motifs, m.motifs) and np.array_equal(m2.factors, m.factors)
    assert m2.loss_hist == m.loss_hist and m2.n_components == 3 and m2.maxlag == 4
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            load_cmfjl_model(path)


class _FakeH5File(dict):
    """A dict-backed stand-in for h5py.File: datasets by name, context manager, `[]` assignment.  The image has no
    h5py; this executes the cmf.jl loader / saver and pins their axis conventions."""
    store = {}

    def __init__(self, path, mode="r"):
        super().__init__()
        self.path, self.mode = path, mode
        if "w" not in mode:
            self.update(_FakeH5File.store[path])

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if "w" in self.mode:
            _FakeH5File.store[self.path] = {k: np.array(v) for k, v in self.items()}
        return False


def test_cmfjl_files_round_trip_and_match_reference_loader(monkeypatch):
    """save_cmfjl_model -> load_cmfjl_model, and (in the build container) the reference's own load_cmfjl_model
    (model.py:346-363) on the very same file object."""
    import sys
    import types
    from cmfpy_b200.model import CMF, load_cmfjl_model, save_cmfjl_model
    fake = types.ModuleType("h5py")
    fake.File = _FakeH5File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    rng = np.random.default_rng(3)
    m = CMF(3, 5)
    m._W, m._H = rng.random((5, 7, 3)), rng.random((3, 40))
    m.loss_hist, m.time_hist = [0.8, 0.6, 0.55], [0.0, 0.01, 0.02]
    X = rng.random((7, 40))
    save_cmfjl_model("mem://model.h5", m, X)
    raw = _FakeH5File.store["mem://model.h5"]
    # what a row-major reader sees of Julia's column-major arrays: data T x N, W K x N x L, H T x K
    assert raw["data"].shape == (40, 7) and raw["W"].shape == (3, 7, 5) and raw["H"].shape == (40, 3)
    data, m2 = load_cmfjl_model("mem://model.h5")
    assert np.array_equal(data, X) and np.array_equal(m2.motifs, m.motifs) and np.array_equal(m2.factors, m.factors)
    assert list(m2.loss_hist) == m.loss_hist and list(m2.time_hist) == m.time_hist
    assert (m2.n_components, m2.maxlag) == (3, 5)
    if ref_shim.available():
        ref_shim.import_reference()
        import cmfpy.model as ref_model
        monkeypatch.setattr(ref_model, "h5py", fake)
        rdata, rm = ref_model.load_cmfjl_model("mem://model.h5")
        assert np.array_equal(rdata, data) and np.array_equal(rm._W, m2.motifs) and np.array_equal(rm._H, m2.factors)
        assert np.array_equal(rm.loss_hist, m2.loss_hist)


def test_cmfjl_files_with_real_h5py(tmp_path):
    h5py = pytest.importorskip("h5py")
    if getattr(h5py, "__file__", None) is None or not hasattr(h5py, "File"):
        pytest.skip("h5py here is the import shim of oracle/ref_shim.py, not the library")
    from cmfpy_b200.model import CMF, load_cmfjl_model, save_cmfjl_model
    rng = np.random.default_rng(4)
    m = CMF(2, 3)
    m._W, m._H = rng.random((3, 4, 2)), rng.random((2, 9))
    m.loss_hist, m.time_hist = [0.5, 0.4], [0.0, 0.1]
    path = str(tmp_path / "m.h5")
    save_cmfjl_model(path, m, rng.random((4, 9)))
    with h5py.File(path, "r") as f:
        assert f["W"].shape == (2, 4, 3)
    _, m2 = load_cmfjl_model(path)
    assert np.array_equal(m2.motifs, m.motifs) and np.array_equal(m2.factors, m.factors)
