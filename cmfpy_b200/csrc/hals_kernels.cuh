// HALS for convolutive NMF (reference cmfpy/algs/hals.py on top of cmfpy/algs/accelerated.py): one coordinate
// block at a time, with the residual R = est - X kept current after every block.
//
//   W sweep (hals.py:78-105), for k, then l:   h = H[k] shifted right by l
//       W[l,:,k] <- max((W[l,:,k] ||h||^2 - R h) / (||h||^2 + eps), FACTOR_MIN);   R += (new - old) (x) h
//   H sweep (hals.py:113-181), for k, then l:  Wk = W[:, :, k] (N x L)
//       batch  t = l, l+L, ... < T-L (disjoint windows => independent):
//       H[k,t] <- max((H[k,t] ||Wk||^2 - <Wk, R[:, t:t+L]>) / (||Wk||^2 + eps), FACTOR_MIN);  R[:, t:t+L] += (new - old) Wk
//       then the entry t = T-L+l with the motif cut to its first T-t lags.
//
// The reference removes a block's contribution from R, solves, and adds the new one back; the forms above are the
// same arithmetic with one pass less (the CPU restatement used by the tests agrees with the reference to 1e-16).
// The sweeps are sequential in (k, l) by definition, so each block update is a pass over R: HBM-bound work.
//   * W: "apply the previous column's correction" and "dot with this column's h" are ONE pass over R^T
//     (8 B / element per column); a small solve kernel sits between two passes.
//   * H: one CTA per batch entry reads its L x N window twice (dot, then update); W[:, :, k] is gathered once
//     per component into a contiguous panel so those reads coalesce.
// R^T is time-major like X^T ([t][n]): feature n is a column, so a warp reads 32 features of one time step as one
// 128-byte line, and a lag is a row offset of H^T.
#pragma once
#include "common.cuh"

namespace cmf {
namespace hals {

constexpr float kFactorMin = 0.f;                 // reference common.py:10

// R^T = est^T - X^T   (3xTF32 storage: est = Et + Elo, X = Xt + Xlo)
__global__ void __launch_bounds__(256)
resid_kernel(float4* __restrict__ Rt, const float4* __restrict__ Et, const float4* __restrict__ Xt,
             const float4* __restrict__ Elo, const float4* __restrict__ Xlo, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 e = Et[i], x = Xt[i];
    float4 r = make_float4(e.x - x.x, e.y - x.y, e.z - x.z, e.w - x.w);
    if (Elo) {
      const float4 el = Elo[i], xl = Xlo[i];
      r.x += el.x - xl.x; r.y += el.y - xl.y; r.z += el.z - xl.z; r.w += el.w - xl.w;
    }
    Rt[i] = r;
  }
}

// One pass over R^T for column c = (k, l) of the W sweep:
//   R[t][n] += delta_prev[n] * hprev(t)        (the correction of the previous column; skipped when hprev == null)
//   part[chunk][n]  = sum_{t in chunk} R[t][n] * hcur(t),   part_h[chunk] = sum_{t in chunk} hcur(t)^2
// h(t) = H^T[h + t - l][k]; the rows before the data are zeros, so t < l needs no branch.
// grid (ceil(Np / 256), n_chunks), 256 threads: thread = feature.
__global__ void __launch_bounds__(256)
w_pass_kernel(float* __restrict__ Rt, int Np, long long T, int rows_per_chunk, const float* __restrict__ hprev,
              const float* __restrict__ delta_prev, const float* __restrict__ hcur, int Kp, float* __restrict__ part,
              float* __restrict__ part_h) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const long long t0 = (long long)blockIdx.y * rows_per_chunk;
  const long long t1 = t0 + rows_per_chunk < T ? t0 + rows_per_chunk : T;
  const float dp = (hprev && n < Np) ? delta_prev[n] : 0.f;
  float acc = 0.f, hh = 0.f;
  if (n < Np) {
    for (long long t = t0; t < t1; ++t) {
      float r = Rt[t * Np + n];
      if (hprev) {
        r = fmaf(dp, hprev[t * Kp], r);
        Rt[t * Np + n] = r;
      }
      if (hcur) {
        const float hv = hcur[t * Kp];
        acc = fmaf(r, hv, acc);
        hh = fmaf(hv, hv, hh);
      }
    }
    if (hcur) part[(size_t)blockIdx.y * Np + n] = acc;
  }
  if (hcur && blockIdx.x == 0 && threadIdx.x == 0) part_h[blockIdx.y] = hh;
}

// Solve column (k, l): g[n] = sum_chunk part, ||h||^2 = sum_chunk part_h,
//   W[l][n][k] <- max((w ||h||^2 - g) / (||h||^2 + eps), FACTOR_MIN),  delta[n] = new - old
__global__ void __launch_bounds__(256)
w_solve_kernel(float* __restrict__ Wcol /* &W[l][0][k] */, int Np, int Kp, const float* __restrict__ part,
               const float* __restrict__ part_h, int n_chunks, float* __restrict__ delta) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Np) return;
  double hn2 = 0.0, g = 0.0;
  for (int c = 0; c < n_chunks; ++c) { hn2 += part_h[c]; g += part[(size_t)c * Np + n]; }
  const float old = Wcol[(size_t)n * Kp];
  const float nw = fmaxf((float)(((double)old * hn2 - g) / (hn2 + (double)kEpsilon)), kFactorMin);
  Wcol[(size_t)n * Kp] = nw;
  delta[n] = nw - old;
}

// Wk[l][n] = W[l][n][k] (contiguous panel of one component) and w2[l] = sum_n W[l][n][k]^2.  grid = L blocks.
__global__ void __launch_bounds__(256)
gather_component_kernel(const float* __restrict__ W, int Np, int Kp, int k, float* __restrict__ Wk, double* __restrict__ w2) {
  const int l = blockIdx.x;
  __shared__ double red[8];
  double s = 0.0;
  for (int n = threadIdx.x; n < Np; n += blockDim.x) {
    const float v = W[((size_t)l * Np + n) * Kp + k];
    Wk[(size_t)l * Np + n] = v;
    s += (double)v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    w2[l] = t;
  }
}

// One CTA per entry of H[k]: entry b is t = t_first + b * t_stride, with the first n_lags lags of the motif.
//   dot = <Wk[:n_lags], R[t : t + n_lags]>,  ||Wk||^2 = sum_{l' < n_lags} w2[l']
//   H[k][t] <- max((H ||Wk||^2 - dot) / (||Wk||^2 + eps), FACTOR_MIN);  R[t + l'][n] += (new - old) Wk[l'][n]
__global__ void __launch_bounds__(256)
h_entries_kernel(float* __restrict__ Rt, int Np, const float* __restrict__ Wk, const double* __restrict__ w2, int n_lags,
                 float* __restrict__ Hcol /* &H^T[h][k] */, int Kp, long long t_first, long long t_stride) {
  const long long t = t_first + (long long)blockIdx.x * t_stride;
  __shared__ double red[8];
  __shared__ float delta_s;
  const long long cnt = (long long)n_lags * Np;            // the window is n_lags contiguous rows of R^T
  float* win = Rt + t * Np;
  double s = 0.0;
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) s += (double)(Wk[i] * win[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double dot = 0.0, n2 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) dot += red[w];
    for (int l = 0; l < n_lags; ++l) n2 += w2[l];
    const float old = Hcol[(size_t)t * Kp];
    const float nw = fmaxf((float)(((double)old * n2 - dot) / (n2 + (double)kEpsilon)), kFactorMin);
    Hcol[(size_t)t * Kp] = nw;
    delta_s = nw - old;
  }
  __syncthreads();
  const float d = delta_s;
  if (d != 0.f)
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) win[i] = fmaf(d, Wk[i], win[i]);
}

// out[0] = sum (a - b)^2 over n floats (one block; for the inner-iteration stop rule, accelerated.py:57-69)
__global__ void __launch_bounds__(1024)
diff_sumsq_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double d = (double)a[i] - (double)b[i]; s += d * d; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v;
  }
}

}  // namespace hals
}  // namespace cmf
