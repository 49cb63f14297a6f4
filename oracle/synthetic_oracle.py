"""TEST INFRASTRUCTURE ONLY - CPU restatement of the device-side input pipeline.

Only tests/ may import this module; nothing under cmfpy_b200/ does.

* `SyntheticOracle` follows reference cmfpy/datasets/synthetic.py:7-46 statement by statement (sparse H :21-25,
  Gaussian-bump motifs :27-30 and :42-46, noise :33, data = cmf_predict(W, H) + noise :36, generate() = data + noise
  :38-39).  The reference draws from NumPy's Mersenne Twister - partly the GLOBAL unseeded generator (:28, :43) - so
  its random stream is not reproducible by construction.  The device generator therefore uses its own counter-based
  stream (splitmix64 finaliser keyed by seed, stream and global element index; cmfpy_b200/csrc/dataset_kernels.cuh),
  which this module restates BIT FOR BIT with NumPy uint64 arithmetic.  Parity with the reference is pinned at two
  levels: structure (every statement above, checked exactly given the same uniform draws) and distribution
  (tests/test_datasets.py compares moments / sparsity / bump shapes with the reference's own class in this container).
* `spectrogram_oracle` is the `generate` step of reference cmfpy/datasets/vox_celeb.py:58-104: it calls
  scipy.signal.spectrogram exactly as the reference does (:89-98) and restates StandardScaler(with_mean=False)
  (:100-102) in NumPy.
"""
import numpy as np

from . import cmf_oracle

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
STREAM_H, STREAM_NOISE, STREAM_MOTIF = 0, 1, 2


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream_key(seed, stream):
    with np.errstate(over="ignore"):
        return mix64(mix64(np.uint64(seed)) + np.uint64(stream))


def draw(key, idx):
    return mix64(key ^ mix64(np.asarray(idx, dtype=np.uint64)))


def u_hi(r):
    return (r >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def u_lo(r):
    return (r & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(2.0 ** -24)


def gauss_plus_delay(n_steps, tau):
    """reference synthetic.py:42-46 with the delay passed in."""
    x = np.linspace(-3 - tau, 3 - tau, n_steps)
    y = np.exp(-x ** 2)
    return y / y.max()


class SyntheticOracle:
    """reference synthetic.py:7-39 on the counter-based stream of the device generator (float32 storage, as on the
    device; the reconstruction in float64)."""

    def __init__(self, n_components=3, n_features=100, n_lags=100, n_timebins=10000, H_sparsity=0.9,
                 noise_scale=1.0, seed=0):
        self.name = "synthetic"
        K, N, L, T = n_components, n_features, n_lags, n_timebins
        # :21-25  H = rand * binomial(1, 1 - H_sparsity)
        r = draw(stream_key(seed, STREAM_H), np.arange(K * T, dtype=np.uint64).reshape(K, T))
        keep = u_lo(r) < np.float32(1.0 - H_sparsity)
        self.H = np.where(keep, u_hi(r), np.float32(0)).astype(np.float32)
        # :27-30  one bump per feature on a random component
        r = draw(stream_key(seed, STREAM_MOTIF), np.arange(N, dtype=np.uint64))
        self.component = np.minimum((u_hi(r) * np.float32(K)).astype(np.int64), K - 1)
        self.tau = -1.5 + 3.0 * u_lo(r).astype(np.float64)
        W = np.zeros((L, N, K))
        for i, j in enumerate(self.component):
            W[:, i, j] += gauss_plus_delay(L, self.tau[i])
        self.W = W.astype(np.float32)
        # :33  noise
        r = draw(stream_key(seed, STREAM_NOISE), np.arange(N * T, dtype=np.uint64).reshape(N, T))
        self.noise = (np.float32(noise_scale) * u_hi(r)).astype(np.float32)
        # :36
        self.data = cmf_oracle.cmf_predict(self.W.astype(np.float64), self.H.astype(np.float64)) + self.noise

    def generate(self):
        return self.data + self.noise            # :38-39


def spectrogram_oracle(audio, sampling_rate, seg_length=20e-3, overlap=0.3, normalize=True):
    """reference vox_celeb.py:58-104 (`VoxCeleb.generate`) without the file handling."""
    from scipy import signal
    nperseg = round(int(seg_length * sampling_rate))          # :89
    noverlap = round(int(nperseg * overlap))                  # :90
    _, _, S = signal.spectrogram(np.asarray(audio, dtype=np.float64), fs=sampling_rate, nperseg=nperseg,
                                 noverlap=noverlap)           # :92-98
    if normalize:                                             # :100-102 StandardScaler(with_mean=False) on S.T
        sd = np.sqrt(S.var(axis=1))
        sd[sd == 0.0] = 1.0
        S = S / sd[:, None]
    return S
