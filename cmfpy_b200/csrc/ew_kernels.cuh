// HBM-bound kernels of the MU iteration: the multiplicative update itself,
// split-sum reduction, loss finalisation, layout/dtype conversion at the ABI
// boundary.  All are coalesced, 128-bit vectorised where the layout allows,
// warp-shuffle reduced, and launched on grids that are multiples of the SM
// count (grid-stride loops).
#pragma once
#include "common.cuh"

namespace cmf {
namespace ew {

// P <- P * num / (den + eps)   (reference cmfpy/algs/mult.py:18 and :22)
// n4 = number of float4 elements.  Optionally rounds the result to TF32 so the
// tensor-core path consumes exactly what is stored.
// Algorithmic bytes: 16 per element (3 reads + 1 write).
// P_op (optional) receives the TF32-rounded copy the tensor-core kernels read;
// the fp32 master P keeps full precision so rounding never accumulates in the state.
// Algorithmic bytes: 16 per element (3 reads + 1 write), +4 with P_op.
__global__ void __launch_bounds__(256)
mu_update_kernel(float4* __restrict__ P, const float4* __restrict__ num,
                 const float4* __restrict__ den, long long n4, float4* __restrict__ P_op) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = P[i];
    const float4 a = __ldcs(num + i), d = __ldcs(den + i);
    p.x = p.x * a.x / (d.x + kEpsilon);
    p.y = p.y * a.y / (d.y + kEpsilon);
    p.z = p.z * a.z / (d.z + kEpsilon);
    p.w = p.w * a.w / (d.w + kEpsilon);
    P[i] = p;
    if (P_op) {
      p.x = round_tf32(p.x); p.y = round_tf32(p.y); p.z = round_tf32(p.z); p.w = round_tf32(p.w);
      P_op[i] = p;
    }
  }
}

// dst = round_tf32(src)
__global__ void __launch_bounds__(256)
round_copy_kernel(float4* __restrict__ dst, const float4* __restrict__ src, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = src[i];
    p.x = round_tf32(p.x); p.y = round_tf32(p.y); p.z = round_tf32(p.z); p.w = round_tf32(p.w);
    dst[i] = p;
  }
}

// out[i] = sum_s part[s * stride_s + i]   (deterministic split-K reduction)
__global__ void __launch_bounds__(256)
sum_splits_kernel(float4* __restrict__ out, const float4* __restrict__ part,
                  long long n4, long long stride4, int nsplit) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 s = __ldcs(part + i);
    for (int k = 1; k < nsplit; ++k) {
      const float4 v = __ldcs(part + (long long)k * stride4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = s;
  }
}

// sum of n doubles -> out[0]  (one block; fixed order => deterministic)
__global__ void __launch_bounds__(1024)
sum_doubles_kernel(const double* __restrict__ in, long long n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v;
  }
}

// The squared residual through the W terms (the loss identity of the Gram route):
//   ||X - est||^2 = ||X||^2 - 2 <X, est> + ||est||^2,   <X, est> = sum W (.) num_W,   ||est||^2 = sum W (.) den_W
// because num_W[l] = X[:, l:] H[:, :T-l]^T and den_W[l] = est[:, l:] H[:, :T-l]^T (mult.py:35-38) are exactly the
// derivatives of the two inner products with respect to W[l].  partial[b] = sum over block b of W (den - 2 num), in
// double; sum_doubles_offset_kernel adds ||X||^2.  Time shards: local num / den partials and the local ||X||^2 give
// local values whose sum over the shards is the global squared residual.
__global__ void __launch_bounds__(256)
wterms_dot_kernel(const float4* __restrict__ W, const float4* __restrict__ num, const float4* __restrict__ den,
                  long long n4, double* __restrict__ partial) {
  __shared__ double red[8];
  double s = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 w = W[i], a = num[i], d = den[i];
    s += (double)w.x * ((double)d.x - 2.0 * (double)a.x) + (double)w.y * ((double)d.y - 2.0 * (double)a.y) +
         (double)w.z * ((double)d.z - 2.0 * (double)a.z) + (double)w.w * ((double)d.w - 2.0 * (double)a.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w];
    partial[blockIdx.x] = v;
  }
}

// out[0] = offset + sum_i in[i]   (fixed order: deterministic)
__global__ void __launch_bounds__(1024)
sum_doubles_offset_kernel(const double* __restrict__ in, long long n, double* __restrict__ out, double offset) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = offset + v;
  }
}

// loss_out[slot] = sqrt(sumsq[0]) / norm_x    (reference base.py:90-97)
__global__ void loss_from_sumsq_kernel(const double* sumsq, double norm_x, double* loss_out, int slot) {
  loss_out[slot] = sqrt(fmax(sumsq[0], 0.0)) / norm_x;
}

// same, slot taken from (and advancing) a device counter: the form a captured CUDA graph replays
__global__ void loss_from_sumsq_counter_kernel(const double* sumsq, double norm_x, double* loss_out, int* counter) {
  const int slot = *counter;
  loss_out[slot] = sqrt(fmax(sumsq[0], 0.0)) / norm_x;
  *counter = slot + 1;
}

// sum of squares + any-negative flag over n floats, per-block partials
// (reference base.py:25 la.norm(data); model.py:138 (data < 0).any())
__global__ void __launch_bounds__(256)
sumsq_neg_kernel(const float4* __restrict__ x, long long n4, double* __restrict__ partial,
                 int* __restrict__ neg_flag) {
  __shared__ double red[8];
  double s = 0.0;
  bool neg = false;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = x[i];
    s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    neg |= (v.x < 0.f) | (v.y < 0.f) | (v.z < 0.f) | (v.w < 0.f);
  }
  if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0) atomicOr(neg_flag, 1);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

// per-block partials of <x, e> and <e, e> (reference rand_init alpha,
// cmfpy/algs/base.py:86-87); partial[2*b] and partial[2*b+1]
__global__ void __launch_bounds__(256)
dot_sumsq_kernel(const float4* __restrict__ x, const float4* __restrict__ e, long long n4,
                 double* __restrict__ partial, const float4* __restrict__ x_lo = nullptr,
                 const float4* __restrict__ e_lo = nullptr) {
  __shared__ double red[2][8];
  double sxe = 0.0, see = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = x[i], b = e[i];
    if (x_lo) { const float4 t = x_lo[i]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    if (e_lo) { const float4 t = e_lo[i]; b.x += t.x; b.y += t.y; b.z += t.z; b.w += t.w; }
    sxe += (double)(a.x * b.x + a.y * b.y) + (double)(a.z * b.z + a.w * b.w);
    see += (double)(b.x * b.x + b.y * b.y) + (double)(b.z * b.z + b.w * b.w);
  }
  sxe = warp_sum(sxe);
  see = warp_sum(see);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sxe; red[1][threadIdx.x >> 5] = see; }
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = (threadIdx.x < 8) ? red[0][threadIdx.x] : 0.0;
    double b = (threadIdx.x < 8) ? red[1][threadIdx.x] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b; }
  }
}

// p[i] *= s  (rescale of the random initialisation, base.py:88)
__global__ void __launch_bounds__(256)
scale_kernel(float4* __restrict__ p, long long n4, float s, int round_out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = p[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    p[i] = v;
  }
}

// dst[c][r] = (float) src[r][c]  for r < rows, c < cols; 32x32 smem tiles.
// src: rows x cols, leading dimension lds.  dst: cols x ldd.
// Used for X (N x T -> T x Np) and H (K x T -> T x Kp) on the way in, and the
// reverse on the way out.  round_out rounds to TF32 on the way in.
template <class TI, class TO>
__global__ void __launch_bounds__(256)
transpose_convert_kernel(const TI* __restrict__ src, long long lds, TO* __restrict__ dst,
                         long long ldd, long long rows, long long cols, int round_out,
                         const TI* __restrict__ src2 = nullptr) {
  __shared__ float tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const long long r = r0 + ty + j, c = c0 + tx;
    float v = (r < rows && c < cols) ? (float)src[r * lds + c] : 0.f;
    if (src2 && r < rows && c < cols) v += (float)src2[r * lds + c];      // hi + lo pair
    tile[ty + j][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const long long c = c0 + ty + j, r = r0 + tx;
    if (c < cols && r < rows) {
      float v = tile[tx][ty + j];
      if (round_out) v = round_tf32(v);
      dst[c * ldd + r] = (TO)v;
    }
  }
}

// W between the reference layout L x N x K (dtype TI) and the padded device
// layout L x Np x Kp (fp32).  One thread per padded element.
template <class TI>
__global__ void __launch_bounds__(256)
w_pad_in_kernel(const TI* __restrict__ src, float* __restrict__ dst, int L, int N, int K,
                int Np, int Kp, int round_out) {
  const long long total = (long long)L * Np * Kp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % Kp);
    const int n = (int)((i / Kp) % Np);
    const int l = (int)(i / ((long long)Kp * Np));
    float v = (k < K && n < N) ? (float)src[((long long)l * N + n) * K + k] : 0.f;
    if (round_out) v = round_tf32(v);
    dst[i] = v;
  }
}
template <class TO>
__global__ void __launch_bounds__(256)
w_pad_out_kernel(const float* __restrict__ src, TO* __restrict__ dst, int L, int N, int K,
                 int Np, int Kp) {
  const long long total = (long long)L * N * K;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % K);
    const int n = (int)((i / K) % N);
    const int l = (int)(i / ((long long)K * N));
    dst[i] = (TO)src[((long long)l * Np + n) * Kp + k];
  }
}

// ---- per-feature statistics / scaling of X^T (time-major: feature n is a column, so a warp reads 32 features of
// one time step as one 128-byte line).  Used for the dataset normalisations of the reference
// (datasets/songbird.py:18-19 L2 rows, datasets/maze.py:71-72 L1 rows, datasets/vox_celeb.py:100-102 unit variance).
// part[chunk][stat][n], stat = sum x, sum x^2, sum |x| over the rows of the chunk; x = hi + lo when lo is given.
__global__ void __launch_bounds__(256)
row_stats_kernel(const float* __restrict__ Xt, const float* __restrict__ Xlo, long long rows, int Np, int rows_per_chunk,
                 double* __restrict__ part) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Np) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  double s1 = 0.0, s2 = 0.0, sa = 0.0;
  for (long long r = r0; r < r1; ++r) {
    float x = Xt[r * Np + n];
    if (Xlo) x += Xlo[r * Np + n];
    s1 += x; s2 += (double)x * x; sa += fabsf(x);
  }
  double* o = part + (size_t)blockIdx.y * 3 * Np;
  o[n] = s1; o[Np + n] = s2; o[2 * Np + n] = sa;
}
// out[stat][n] = sum_chunk part[chunk][stat][n]
__global__ void __launch_bounds__(256)
row_stats_sum_kernel(const double* __restrict__ part, int nchunks, int Np, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * Np) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += part[(size_t)c * 3 * Np + i];
  out[i] = s;
}
// X^T[r][n] *= scale[n]   (all rows, halo included; hi + lo pairs are recombined first and left unsplit in Xt)
__global__ void __launch_bounds__(256)
scale_rows_kernel(float* __restrict__ Xt, float* __restrict__ Xlo, long long rows, int Np, const float* __restrict__ scale,
                  int round_out) {
  const long long total = rows * Np;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float x = Xt[i];
    if (Xlo) { x += Xlo[i]; Xlo[i] = 0.f; }
    x *= scale[i % Np];
    Xt[i] = round_out ? round_tf32(x) : x;
  }
}

// elementwise dtype conversion (staging of float64 host data)
template <class TI, class TO>
__global__ void __launch_bounds__(256)
convert_kernel(const TI* __restrict__ src, TO* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = (TO)src[i];
}

}  // namespace ew
}  // namespace cmf
