// The step before the solver, on the device: the reference's synthetic data set
// (cmfpy/datasets/synthetic.py:7-46) and the spectrogram front end of its audio loader
// (cmfpy/datasets/vox_celeb.py:58-104).  All HBM-bound except the per-segment DFT.
//
// Random numbers come from a counter-based generator: element `idx` of stream `s` under seed `seed` is
//     r = mix64(key ^ mix64(idx)),  key = mix64(mix64(seed) + s),   mix64 = the splitmix64 finaliser,
// so a value depends only on (seed, stream, GLOBAL index): time shards of any width generate the very same
// data set, and the test suite restates the stream bit for bit in NumPy (tests/test_datasets.py).
#pragma once
#include "common.cuh"

namespace cmf {
namespace ds {

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t stream_key(uint64_t seed, uint32_t stream) { return mix64(mix64(seed) + stream); }
__host__ __device__ inline uint64_t draw(uint64_t key, uint64_t idx) { return mix64(key ^ mix64(idx)); }
// two independent U[0,1) values with 24 bits each from one draw
__host__ __device__ inline float u_hi(uint64_t r) { return (float)(uint32_t)(r >> 40) * 5.9604644775390625e-8f; }
__host__ __device__ inline float u_lo(uint64_t r) { return (float)(uint32_t)(r & 0xFFFFFFu) * 5.9604644775390625e-8f; }

enum : uint32_t { kStreamH = 0, kStreamNoise = 1, kStreamMotif = 2 };

// H[k][c] = U * Bernoulli(1 - H_sparsity) for global column t0 + c (synthetic.py:22-25); zeros outside [0, T).
// H is K x ld row-major, c < ncols.
__global__ void __launch_bounds__(256)
synth_h_kernel(float* __restrict__ H, int K, long long ld, long long ncols, long long t0, long long T, uint64_t key,
               float keep) {
  const long long total = (long long)K * ncols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long k = i / ncols, c = i % ncols, t = t0 + c;
    float v = 0.f;
    if (t >= 0 && t < T) {
      const uint64_t r = draw(key, (uint64_t)(k * T + t));
      v = u_lo(r) < keep ? u_hi(r) : 0.f;
    }
    H[k * ld + c] = v;
  }
}

// W[:, n, j_n] = gauss_plus_delay(L) (synthetic.py:27-30, 42-46): one component j_n per feature, a Gaussian bump
// exp(-x^2) on x = linspace(-3 - tau, 3 - tau, L), tau ~ U(-1.5, 1.5), scaled to a maximum of one.  W is L x N x K.
__global__ void __launch_bounds__(128)
synth_w_kernel(float* __restrict__ W, int L, int N, int K, uint64_t key) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const uint64_t r = draw(key, (uint64_t)n);
  int j = (int)(u_hi(r) * (float)K);
  if (j >= K) j = K - 1;
  const double tau = -1.5 + 3.0 * (double)u_lo(r);
  const double a = -3.0 - tau, b = 3.0 - tau;
  const double step = L > 1 ? (b - a) / (double)(L - 1) : 0.0;
  double ymax = 0.0;
  for (int l = 0; l < L; ++l) {
    const double x = (l == L - 1 && L > 1) ? b : a + step * l;
    const double y = exp(-x * x);
    ymax = y > ymax ? y : ymax;
  }
  for (int l = 0; l < L; ++l) {
    const double x = (l == L - 1 && L > 1) ? b : a + step * l;
    const double y = exp(-x * x) / ymax;
    float* w = W + ((size_t)l * N + n) * K;
    for (int k = 0; k < K; ++k) w[k] = k == j ? (float)y : 0.f;
  }
}

// D[n][c] = (base ? base[n][c] : 0) + times * scale * U(n, t0 + c)   (synthetic.py:33, 36, 39); N x ld row-major
__global__ void __launch_bounds__(256)
synth_noise_kernel(float* __restrict__ D, const float* __restrict__ base, int N, long long ld, long long ncols,
                   long long t0, long long T, uint64_t key, float scale, float times) {
  const long long total = (long long)N * ncols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long n = i / ncols, c = i % ncols;
    const float u = scale * u_hi(draw(key, (uint64_t)(n * T + t0 + c)));
    float v = base ? base[n * ld + c] : 0.f;
    for (float k = 0.f; k < times; k += 1.f) v += u;       // `generate()` adds the same noise a second time
    D[n * ld + c] = v;
  }
}

// dst (rows x cols, ld ldd, double or float) <- src (float, ld lds)
template <class TO>
__global__ void __launch_bounds__(256)
copy_convert_kernel(TO* __restrict__ dst, long long ldd, const float* __restrict__ src, long long lds, long long rows,
                    long long cols) {
  const long long total = rows * cols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / cols, c = i % cols;
    dst[r * ldd + c] = (TO)src[r * lds + c];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Spectrogram (scipy.signal.spectrogram as vox_celeb.py:92-98 calls it: mode 'psd', scaling 'density',
// detrend 'constant', one-sided, the window passed in).  Segment s covers samples [s hop, s hop + nperseg);
// S[f][s] = c_f |sum_n w[n] (x[n] - mean_s) e^{-2 pi i f n / nperseg}|^2 / (fs sum w^2), c_f = 2 except at DC and
// (even nperseg) Nyquist.  One block transforms kSegs consecutive segments (a direct DFT from a shared twiddle
// table: nperseg is a few hundred, the whole spectrogram a few ms) and writes S in 128-byte rows.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSegs = 32;

__global__ void __launch_bounds__(256)
spectrogram_kernel(const float* __restrict__ audio, long long n_samples, const float* __restrict__ window, int nperseg,
                   int segld, int hop, long long n_seg, int n_freq, float scale, float* __restrict__ S, long long ld) {
  extern __shared__ float sm[];
  float* cs = sm;                          // cos(2 pi m / nperseg)
  float* sn = cs + nperseg;                // sin(2 pi m / nperseg)
  float* seg = sn + nperseg;               // [kSegs][segld] windowed, detrended samples (segld odd: no bank conflicts)
  float* out = seg + (size_t)kSegs * segld;     // [n_freq][kSegs + 1]
  const long long s0 = (long long)blockIdx.x * kSegs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int m = tid; m < nperseg; m += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)m / (float)nperseg, &s, &c);
    cs[m] = c; sn[m] = s;
  }
  // one warp per segment: mean, then window * (x - mean)
  for (int j = warp; j < kSegs; j += blockDim.x / 32) {
    const long long s = s0 + j;
    float* dst = seg + (size_t)j * segld;
    if (s >= n_seg) { for (int m = lane; m < nperseg; m += 32) dst[m] = 0.f; continue; }
    const float* x = audio + s * hop;
    float acc = 0.f;
    for (int m = lane; m < nperseg; m += 32) acc += x[m];
    const float mean = warp_sum(acc) / (float)nperseg;
    for (int m = lane; m < nperseg; m += 32) dst[m] = window[m] * (x[m] - mean);
  }
  __syncthreads();
  // (frequency, segment) pairs: the 32 lanes of a warp take the 32 segments of one frequency
  for (int f = warp; f < n_freq; f += blockDim.x / 32) {
    const float* x = seg + (size_t)lane * segld;
    float re = 0.f, im = 0.f;
    int ph = 0;                            // (f m) mod nperseg
    for (int m = 0; m < nperseg; ++m) {
      const float v = x[m];
      re = fmaf(v, cs[ph], re);
      im = fmaf(v, sn[ph], im);
      ph += f; if (ph >= nperseg) ph -= nperseg;
    }
    const bool edge = (f == 0) || ((nperseg & 1) == 0 && f == nperseg / 2);
    out[(size_t)f * (kSegs + 1) + lane] = (re * re + im * im) * scale * (edge ? 1.f : 2.f);
  }
  __syncthreads();
  for (int i = tid; i < n_freq * kSegs; i += blockDim.x) {
    const int f = i / kSegs, j = i % kSegs;
    if (s0 + j < n_seg) S[(size_t)f * ld + s0 + j] = out[(size_t)f * (kSegs + 1) + j];
  }
}

// Per-row sums of a row-major rows x cols matrix: part[(r * nchunk + chunk) * 2 + {0, 1}] = sum x, sum x^2 (double)
__global__ void __launch_bounds__(256)
row_moments_kernel(const float* __restrict__ X, long long ld, long long cols, long long chunk_cols, int nchunk,
                   double* __restrict__ part) {
  const long long r = blockIdx.y;
  const int chunk = blockIdx.x;
  const long long c0 = (long long)chunk * chunk_cols;
  const long long c1 = c0 + chunk_cols < cols ? c0 + chunk_cols : cols;
  double s1 = 0.0, s2 = 0.0;
  for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    const double v = (double)X[r * ld + c];
    s1 += v; s2 += v * v;
  }
  __shared__ double sh[2][8];
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
    part[((size_t)r * nchunk + chunk) * 2] = a;
    part[((size_t)r * nchunk + chunk) * 2 + 1] = b;
  }
}

// StandardScaler(with_mean=False) (vox_celeb.py:100-102): rows divided by their population standard deviation
// (a zero deviation scales by one, as scikit-learn does)
__global__ void __launch_bounds__(256)
row_std_scale_kernel(float* __restrict__ X, long long ld, long long cols, const double* __restrict__ part, int nchunk) {
  const long long r = blockIdx.y;
  __shared__ float inv;
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int c = 0; c < nchunk; ++c) { a += part[((size_t)r * nchunk + c) * 2]; b += part[((size_t)r * nchunk + c) * 2 + 1]; }
    const double mean = a / (double)cols;
    double var = b / (double)cols - mean * mean;
    if (var < 0.0) var = 0.0;
    const double sd = sqrt(var);
    inv = sd > 0.0 ? (float)(1.0 / sd) : 1.f;
  }
  __syncthreads();
  const float s = inv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += stride) X[r * ld + c] *= s;
}

}  // namespace ds
}  // namespace cmf
