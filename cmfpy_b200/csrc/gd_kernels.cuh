// Projected gradient descent / block coordinate descent for CMF
// (reference cmfpy/algs/gradient_descent.py:15-159) on top of the MU contractions:
//   gW[l] = s_T_dot(resids, H, l) = den_W[l] - num_W[l]     (gradient_descent.py:41-45; mult.py:37-38)
//   gH    = sum_l W[l]^T shift(resids, -l) = den_H - num_H   (gradient_descent.py:47-52; mult.py:44-46)
// so the gradients are the W-term / H-term buffers the MU kernels already fill, and the step
//   x <- max(x - ss * g, 0)                                   (gradient_descent.py:148-159)
// is one HBM-bound pass.  The W step size is 1 / lambda_max of the (K L) x (K L) block-Toeplitz matrix of the
// lag autocorrelations of H (lipschitz_W, gradient_descent.py:54-69): the autocorrelations come from the W-terms
// kernel run on H^T itself and lambda_max from a power iteration that stays on the device.
#pragma once
#include "common.cuh"

namespace cmf {
namespace gd {

// P <- max(P - ss * (den - num), 0);  ss = *step_dev when given (1 / lipschitz_W, left on the device), else step_host.
// P_op (optional): TF32-rounded operand copy, as in ew::mu_update_kernel.  16 B / element algorithmic.
__global__ void __launch_bounds__(256)
projected_step_kernel(float4* __restrict__ P, const float4* __restrict__ num, const float4* __restrict__ den, long long n4,
                      const float* __restrict__ step_dev, float step_host, float4* __restrict__ P_op) {
  const float ss = step_dev ? *step_dev : step_host;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = P[i];
    const float4 a = __ldcs(num + i), d = __ldcs(den + i);
    p.x = fmaxf(p.x - ss * (d.x - a.x), 0.f);
    p.y = fmaxf(p.y - ss * (d.y - a.y), 0.f);
    p.z = fmaxf(p.z - ss * (d.z - a.z), 0.f);
    p.w = fmaxf(p.w - ss * (d.w - a.w), 0.f);
    P[i] = p;
    if (P_op) {
      p.x = round_tf32(p.x); p.y = round_tf32(p.y); p.z = round_tf32(p.z); p.w = round_tf32(p.w);
      P_op[i] = p;
    }
  }
}

// A operand of the autocorrelation pass on the FFMA path: rows of H^T in the role of the data
// (P[d][a][b] = sum_t H[a][t] H[b][t-d] = s_T_dot(H, H, d)[a][b], gradient_descent.py:56)
struct AutoA {
  static constexpr bool kAlongR = false;
  const float* Ht_owned; int Kp;
  __device__ const float* ptr(long long a, long long tau, int) const {
    if (a >= Kp) return nullptr;
    return Ht_owned + tau * Kp + a;
  }
};
// part[split][d][a][b] = v
struct AutoEpi {
  static constexpr bool kReduce = false;
  float* part; int Kp, LKp; long long per_split;
  __device__ float store(long long a, long long c, float4 v, int split, int) const {
    if (a >= Kp || c >= LKp) return 0.f;
    const int d = (int)(c / Kp), b = (int)(c % Kp);
    *reinterpret_cast<float4*>(part + (long long)split * per_split + ((long long)d * Kp + a) * Kp + b) = v;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};

// ---- lambda_max of hW, hW[(i,a),(j,b)] = P[j-i][a][b] (i <= j), P[i-j][b][a] (i > j) ------------------------
// (gradient_descent.py:58-69) by power iteration on the device: a multi-block matrix-vector product with the
// block-Toeplitz structure (the matrix itself is never formed) and a one-block Rayleigh quotient / normalisation.
// State: [0] lambda, [1] previous lambda, [2] settled count, [3] done flag (doubles); v persists across calls
// (warm start: H moves little between iterations).  Once `done` is set the remaining launches return at once.
struct PowerState { double lam, lam_prev, settled, done; };

// Pt[d][a][b] = P[d][b][a]: both triangles of hW then read contiguous rows
__global__ void __launch_bounds__(256)
transpose_blocks_kernel(const float* __restrict__ P, float* __restrict__ Pt, int L, int Kp) {
  const long long total = (long long)L * Kp * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % Kp), a = (int)((i / Kp) % Kp);
    const long long d = i / ((long long)Kp * Kp);
    Pt[i] = P[(d * Kp + b) * Kp + a];
  }
}

__global__ void power_reset_kernel(PowerState* st, float* v, int n) {
  // keep a usable stored vector (warm start); otherwise start from ones
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  double s2 = 0.0;
  for (int o = threadIdx.x; o < n; o += blockDim.x) { const float x = v[o]; s2 += (double)x * x; if (!isfinite(x)) bad = 1; }
  s2 = warp_sum(s2);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    if (!(t > 1e-30)) bad = 1;
    red[0] = t;
    st->lam = 0.0; st->lam_prev = -1.0; st->settled = 0.0; st->done = 0.0;
  }
  __syncthreads();
  const float scale = bad ? 0.f : (float)(1.0 / sqrt(red[0]));
  const float fill = (float)(1.0 / sqrt((double)n));
  for (int o = threadIdx.x; o < n; o += blockDim.x) v[o] = bad ? fill : v[o] * scale;
}

// y = hW v: one warp per output (i,a); the lanes stride over the (j,b) pairs of the reduction
__global__ void __launch_bounds__(256)
toeplitz_matvec_kernel(const float* __restrict__ P, const float* __restrict__ Pt, const float* __restrict__ v,
                       float* __restrict__ y, int L, int Kp, const PowerState* __restrict__ st) {
  if (st->done != 0.0) return;
  const int n = L * Kp;
  const int lane = threadIdx.x & 31;
  const int o = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (o >= n) return;
  const int i = o / Kp, a = o - i * Kp;
  float acc = 0.f;
  for (int q = lane; q < n; q += 32) {
    const int j = q / Kp, b = q - j * Kp;
    const float m = (i <= j) ? P[((size_t)(j - i) * Kp + a) * Kp + b] : Pt[((size_t)(i - j) * Kp + a) * Kp + b];
    acc = fmaf(m, v[q], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) y[o] = acc;
}

// lambda = v.y (||v|| = 1), v <- y / ||y||; sets done after two consecutive relative changes <= tol
__global__ void __launch_bounds__(1024)
power_normalize_kernel(float* __restrict__ v, const float* __restrict__ y, int n, PowerState* st, double tol,
                       double* __restrict__ lam_out, float* __restrict__ inv_out) {
  if (st->done != 0.0) return;
  __shared__ double red[2][32];
  __shared__ double vy_s, yy_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double vy = 0.0, yy = 0.0;
  for (int o = tid; o < n; o += blockDim.x) { const double a = v[o], b = y[o]; vy += a * b; yy += b * b; }
  vy = warp_sum(vy); yy = warp_sum(yy);
  if (lane == 0) { red[0][warp] = vy; red[1][warp] = yy; }
  __syncthreads();
  if (warp == 0) {
    double x = lane < nw ? red[0][lane] : 0.0, z = lane < nw ? red[1][lane] : 0.0;
    x = warp_sum(x); z = warp_sum(z);
    if (lane == 0) { vy_s = x; yy_s = z; }
  }
  __syncthreads();
  const double lam = vy_s, nrm2 = yy_s;
  if (nrm2 > 0.0) {
    const float inv = (float)(1.0 / sqrt(nrm2));
    for (int o = tid; o < n; o += blockDim.x) v[o] = y[o] * inv;
  }
  if (tid == 0) {
    const bool close = fabs(lam - st->lam_prev) <= tol * fabs(lam);
    st->settled = close ? st->settled + 1.0 : 0.0;
    st->lam_prev = lam;
    st->lam = lam;
    if (st->settled >= 2.0 || !(nrm2 > 0.0)) st->done = 1.0;
    *lam_out = lam;
    *inv_out = lam > 0.0 ? (float)(1.0 / lam) : 0.f;
  }
}

}  // namespace gd
}  // namespace cmf
