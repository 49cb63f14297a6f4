"""The device-side input pipeline: `Synthetic` (reference datasets/synthetic.py:7-46) and the
spectrogram step (reference datasets/vox_celeb.py:58-104).

CPU part: the NumPy restatement of the generator against the reference's own class (structure exactly, random
draws at distribution level - the reference uses NumPy's global generator and cannot be reproduced), the
spectrogram restatement against scikit-learn's scaler.  GPU part (-m gpu): the device generator against the
restatement (random streams bit for bit), time sharding, device matrices through `CMF.fit`, the device spectrogram
against scipy.
"""
import os

import numpy as np
import pytest

from oracle import cmf_oracle as o
from oracle import synthetic_oracle as so

REFERENCE = "/root/reference"


# ----------------------------------------------------------------------------- CPU: the restatement
def test_counter_stream_known_answers():
    """mix64 is the splitmix64 finaliser: compare the vectorised uint64 code with exact Python integers."""
    def mix_py(z):
        m = (1 << 64) - 1
        z = (z + 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        return z ^ (z >> 31)
    xs = [0, 1, 2, 12345, (1 << 64) - 1, 0x0123456789ABCDEF]
    got = so.mix64(np.array(xs, dtype=np.uint64))
    assert [int(v) for v in got] == [mix_py(x) for x in xs]
    assert int(so.mix64(np.uint64(0))) == 0xE220A8397B1DCDAF          # first output of splitmix64 seeded with 0
    key = so.stream_key(7, so.STREAM_NOISE)
    assert int(key) == mix_py((mix_py(7) + so.STREAM_NOISE) & ((1 << 64) - 1))
    r = so.draw(key, np.arange(4, dtype=np.uint64))
    assert [int(v) for v in r] == [mix_py(int(key) ^ mix_py(i)) for i in range(4)]


def test_counter_stream_is_uniform():
    from scipy import stats
    r = so.draw(so.stream_key(3, so.STREAM_H), np.arange(200000, dtype=np.uint64))
    for u in (so.u_hi(r), so.u_lo(r)):
        assert u.min() >= 0.0 and u.max() < 1.0
        assert stats.kstest(u.astype(np.float64), "uniform").pvalue > 1e-3
    # the two halves of a draw, neighbouring counters and neighbouring streams are uncorrelated
    r2 = so.draw(so.stream_key(3, so.STREAM_NOISE), np.arange(200000, dtype=np.uint64))
    for a, b in ((so.u_hi(r), so.u_lo(r)), (so.u_hi(r)[1:], so.u_hi(r)[:-1]), (so.u_hi(r), so.u_hi(r2))):
        assert abs(np.corrcoef(a, b)[0, 1]) < 0.01


def test_synthetic_restatement_structure():
    """Every statement of synthetic.py:21-39 on the oracle's own draws."""
    s = so.SyntheticOracle(n_components=4, n_features=30, n_lags=12, n_timebins=500, H_sparsity=0.8,
                           noise_scale=0.5, seed=11)
    assert s.W.shape == (12, 30, 4) and s.H.shape == (4, 500) and s.data.shape == (30, 500)
    assert abs((s.H > 0).mean() - 0.2) < 0.03 and s.H.max() < 1.0
    assert s.noise.min() >= 0 and s.noise.max() < 0.5 and abs(s.noise.mean() - 0.25) < 0.01
    for n in range(30):                                     # one bump per feature, maximum one, on one component
        nz = np.flatnonzero(s.W[:, n, :].sum(axis=0))
        assert list(nz) == [s.component[n]]
        assert s.W[:, n, s.component[n]].max() == pytest.approx(1.0)
        np.testing.assert_allclose(s.W[:, n, s.component[n]], so.gauss_plus_delay(12, s.tau[n]), rtol=1e-6)
    np.testing.assert_allclose(s.data, o.cmf_predict(s.W.astype(float), s.H.astype(float)) + s.noise, rtol=1e-12)
    np.testing.assert_allclose(s.generate(), s.data + s.noise, rtol=1e-12)
    assert -1.5 <= s.tau.min() and s.tau.max() < 1.5


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree only exists in the build container")
def test_synthetic_matches_reference_distribution():
    """The reference's own class (global NumPy generator: not reproducible) against the restatement, at the level of
    distributions: sparsity and moments of H, noise range, one unit-height bump per feature with the reference's
    own `_gauss_plus_delay` shape, moments of the data."""
    from oracle import ref_shim
    ref_shim.import_reference()                      # the unmodified reference behind its import shims
    from cmfpy.datasets.synthetic import Synthetic, _gauss_plus_delay
    kw = dict(n_components=5, n_features=60, n_lags=25, n_timebins=4000, H_sparsity=0.9, noise_scale=0.7)
    np.random.seed(5)
    ref = Synthetic(seed=5, **kw)
    mine = so.SyntheticOracle(seed=5, **kw)
    assert ref.W.shape == mine.W.shape and ref.H.shape == mine.H.shape and ref.data.shape == mine.data.shape
    assert abs((ref.H > 0).mean() - (mine.H > 0).mean()) < 0.02
    assert abs(ref.H[ref.H > 0].mean() - mine.H[mine.H > 0].mean()) < 0.03          # U[0,1) where non-zero
    assert abs(ref.H[ref.H > 0].std() - mine.H[mine.H > 0].std()) < 0.03
    assert abs(ref.noise.mean() - mine.noise.mean()) < 0.01 and abs(ref.noise.std() - mine.noise.std()) < 0.01
    for s in (ref, mine):
        per_feature = (s.W.sum(axis=0) > 0).sum(axis=1)
        assert (per_feature == 1).all()                      # exactly one component carries each feature
        assert np.allclose(s.W.max(axis=(0, 2)), 1.0)
    # a bump of the restatement is the reference's function for the same delay: reproduce it through the global RNG
    class FixedTau:
        def __init__(self, tau):
            self.tau = tau
        def uniform(self, lo, hi):
            assert (lo, hi) == (-1.5, 1.5)
            return self.tau
    import cmfpy.datasets.synthetic as mod
    saved = mod.np.random
    try:
        for n in range(5):
            mod.np = type("np_proxy", (), {"random": FixedTau(mine.tau[n]), "linspace": np.linspace, "exp": np.exp})
            np.testing.assert_allclose(mine.W[:, n, mine.component[n]], _gauss_plus_delay(25), rtol=1e-6)
    finally:
        mod.np = np
        assert mod.np.random is saved
    # first two moments of the data agree to sampling error
    assert abs(ref.data.mean() - mine.data.mean()) / ref.data.mean() < 0.1
    assert abs(ref.generate().mean() - mine.generate().mean()) / ref.generate().mean() < 0.1


def test_spectrogram_restatement_scaler():
    from sklearn import preprocessing
    rng = np.random.default_rng(0)
    audio = rng.standard_normal(8000)
    S = so.spectrogram_oracle(audio, 8000, normalize=False)
    assert S.shape == (81, (8000 - 48) // (160 - 48))
    ref = preprocessing.StandardScaler(with_mean=False).fit_transform(S.T).T       # vox_celeb.py:100-102
    np.testing.assert_allclose(so.spectrogram_oracle(audio, 8000, normalize=True), ref, rtol=1e-10)


# ----------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_device_synthetic_against_restatement(precision):
    from cmfpy_b200.datasets import Synthetic
    kw = dict(n_components=4, n_features=70, n_lags=16, n_timebins=3000, H_sparsity=0.85, noise_scale=0.3, seed=21)
    dev = Synthetic(precision=precision, **kw)
    ref = so.SyntheticOracle(**kw)
    assert dev.name == "synthetic"
    assert np.array_equal(dev.H, ref.H.astype(np.float64))             # the random streams are bit-exact
    assert np.array_equal(dev.noise, ref.noise.astype(np.float64))
    np.testing.assert_allclose(dev.W, ref.W, rtol=2e-6, atol=1e-12)
    assert np.array_equal(dev.W > 0, ref.W > 0)
    scale = np.abs(ref.data).max()
    np.testing.assert_allclose(dev.data, ref.data, atol=2e-6 * scale)
    np.testing.assert_allclose(dev.generate(), ref.generate(), atol=2e-6 * scale)
    m = dev.device_data()
    assert m.shape == (70, 3000)
    assert np.array_equal(m.to_host(), dev.data)
    g = dev.device_generate()
    assert np.array_equal(g.to_host(), dev.generate())
    dev.close()


@gpu
def test_device_synthetic_time_shards_agree():
    """Columns [t0, t0 + n) generated as a shard equal the same columns of the whole data set."""
    from cmfpy_b200.datasets import Synthetic
    kw = dict(n_components=3, n_features=40, n_lags=20, n_timebins=2048, H_sparsity=0.9, noise_scale=1.0, seed=4,
              precision="fp32")
    whole = Synthetic(**kw)
    for t0, n in ((0, 700), (700, 648), (5, 30), (1348, 700)):
        part = Synthetic(t_offset=t0, t_local=n, **kw)
        assert np.array_equal(part.H, whole.H[:, t0:t0 + n])
        assert np.array_equal(part.noise, whole.noise[:, t0:t0 + n])
        np.testing.assert_allclose(part.data, whole.data[:, t0:t0 + n], rtol=1e-6, atol=1e-7)
        part.close()
    whole.close()


@gpu
def test_fit_from_device_matrix():
    """CMF.fit on a device matrix = CMF.fit on its host copy; the negativity check runs on the device."""
    from cmfpy_b200 import CMF
    from cmfpy_b200.datasets import Synthetic
    ds = Synthetic(n_components=3, n_features=50, n_lags=10, n_timebins=1500, H_sparsity=0.9, noise_scale=0.1, seed=2)
    rng = np.random.default_rng(0)
    W0, H0 = rng.random((10, 50, 3)), rng.random((3, 1500))
    a = CMF(3, 10, n_iter_max=8, verbose=False, tol=0, initW=W0, initH=H0)
    a.fit(ds.device_data())
    b = CMF(3, 10, n_iter_max=8, verbose=False, tol=0, initW=W0, initH=H0)
    b.fit(ds.data.astype(np.float32))
    np.testing.assert_allclose(a.loss_hist, b.loss_hist, rtol=1e-7)
    np.testing.assert_allclose(a.motifs, b.motifs, rtol=1e-6)
    ref = o.MultUpdateOracle(ds.data, 10, 3, initW=W0, initH=H0, tol=0)
    hist = [ref.loss] + [ref.update() for _ in range(8)]
    np.testing.assert_allclose(a.loss_hist, hist, rtol=1e-4)
    neg = Synthetic(n_components=3, n_features=50, n_lags=10, n_timebins=1500, noise_scale=-5.0, seed=2)
    with pytest.raises(ValueError, match="Negative values"):
        CMF(3, 10, n_iter_max=2, verbose=False).fit(neg.device_data())


@gpu
@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("fs,seconds", [(16000, 1.5), (8000, 2.0), (11025, 1.0)])
def test_device_spectrogram_against_scipy(normalize, fs, seconds):
    from cmfpy_b200.datasets import spectrogram
    rng = np.random.default_rng(fs)
    t = np.arange(int(fs * seconds)) / fs
    audio = (np.sin(2 * np.pi * 440 * t) + 0.5 * np.sin(2 * np.pi * 1800 * t * (1 + 0.2 * t)) +
             0.1 * rng.standard_normal(t.size) + 0.3)
    ref = so.spectrogram_oracle(audio, fs, normalize=normalize)
    for src in (audio, audio.astype(np.float32)):
        S = spectrogram(src, fs, normalize=normalize)
        assert S.shape == ref.shape
        got = S.to_host()
        # float32 DFT of a few hundred points against scipy's float64 FFT; bins are compared relative to the
        # largest bin of their segment (a power spectrum spans many decades)
        tol = 2e-5 * np.abs(ref).max(axis=0, keepdims=True) if not normalize else 2e-4 * np.abs(ref).max()
        assert np.all(np.abs(got - ref) <= tol), float(np.abs(got - ref).max())
        assert got.min() >= 0.0


@gpu
def test_spectrogram_feeds_the_solver():
    from cmfpy_b200 import CMF
    from cmfpy_b200.datasets import spectrogram
    rng = np.random.default_rng(1)
    audio = np.abs(rng.standard_normal(16000 * 2))
    S = spectrogram(audio, 16000)
    m = CMF(2, 5, n_iter_max=5, verbose=False, tol=0, seed=0)
    m.fit(S)
    assert len(m.loss_hist) == 6 and all(np.diff(m.loss_hist) <= 1e-6)
    assert m.motifs.shape == (5, S.shape[0], 2)
    with pytest.raises(ValueError):
        spectrogram(audio[:100], 16000)
