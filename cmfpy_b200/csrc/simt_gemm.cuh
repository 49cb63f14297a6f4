// Exact-fp32 shift-GEMM on the FFMA pipe.
//
// One register-tiled, double-buffered SGEMM skeleton, C[m][c] = sum_r A(m,r)*B(r,c),
// whose operands are *functors*: the three contractions of the MU iteration
// (reconstruction, W terms, H terms) differ only in how (row, reduction index)
// maps to an address in the time-major device arrays, and a lag is just an
// offset in that mapping - no shifted copy of H, X or est is ever materialised.
//
// This is the CMF_PREC_FP32 path, and the general-shape path (any N, K, L, T).
// The tcgen05 path in tc_*.cuh replaces it for CMF_PREC_TF32.
#pragma once
#include "common.cuh"

namespace cmf {
namespace simt {

// Operand functor concept
//   static constexpr bool kAlongR;
//       true : the 4 elements (i, r..r+3) are contiguous in memory
//       false: the 4 elements (i..i+3, r) are contiguous in memory
//   __device__ const float* ptr(long long i, long long r, int src) const;
//       address of element (i, r) - 16-byte aligned, 4 valid floats - or
//       nullptr when the quad lies outside the operand (reads as zeros).
// Epilogue functor concept
//   static constexpr bool kReduce;
//   __device__ float store(long long m, long long c, float4 v, int split, int src) const;
//   __device__ void  block_sum(double s, long long block_linear) const;   (if kReduce)

template <int BM, int BN, int BK, int TM, int TN, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
shift_gemm_kernel(const AOp a, const BOp b, const Epi epi, long long R, long long r_chunk, int nsrc) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TXN = BN / TN;          // threads along the column dimension
  constexpr int NV = TN / 4;            // float4 vectors per thread row
  constexpr int MV = TM / 4;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  static_assert(TM % 4 == 0 && TN % 4 == 0 && BK % 4 == 0, "tile shape");
  constexpr int QA = (BM * BK / 4 + NT - 1) / NT;
  constexpr int QB = (BN * BK / 4 + NT - 1) / NT;

  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const long long m0 = (long long)blockIdx.x * BM;
  const long long c0 = (long long)blockIdx.y * BN;
  const int split = blockIdx.z / nsrc, src = blockIdx.z % nsrc;
  const long long r_begin = (long long)split * r_chunk;
  const long long r_end = min(R, r_begin + r_chunk);
  const int ntiles = (int)((r_end - r_begin + BK - 1) / BK);

  // Two-level accumulation: `acc` collects kFlush reduction tiles, then is
  // folded into `tot`.  Keeps the fp32 rounding error of a length-R sum near
  // sqrt(kFlush*BK) + sqrt(R/(kFlush*BK)) ulps instead of sqrt(R) (the reference
  // sums in float64; this is what holds the fp32 path inside the 1e-4 bar with margin).
  constexpr int kFlush = 8;
  float acc[TM][TN], tot[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }

  float4 ra[QA], rb[QB];

  auto fetch = [&](int t) {
    const long long r0 = r_begin + (long long)t * BK;
#pragma unroll
    for (int it = 0; it < QA; ++it) {
      const int q = tid + it * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < BM * BK / 4) {
        const float* p;
        if (AOp::kAlongR) {
          const int i = q / (BK / 4), rq = q % (BK / 4);
          const long long r = r0 + rq * 4;
          p = (r < r_end) ? a.ptr(m0 + i, r, src) : nullptr;
        } else {
          const int r = q / (BM / 4), iq = q % (BM / 4);
          p = (r0 + r < r_end) ? a.ptr(m0 + iq * 4, r0 + r, src) : nullptr;
        }
        if (p) v = __ldg(reinterpret_cast<const float4*>(p));
      }
      ra[it] = v;
    }
#pragma unroll
    for (int it = 0; it < QB; ++it) {
      const int q = tid + it * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < BN * BK / 4) {
        const float* p;
        if (BOp::kAlongR) {
          const int i = q / (BK / 4), rq = q % (BK / 4);
          const long long r = r0 + rq * 4;
          p = (r < r_end) ? b.ptr(c0 + i, r, src) : nullptr;
        } else {
          const int r = q / (BN / 4), iq = q % (BN / 4);
          p = (r0 + r < r_end) ? b.ptr(c0 + iq * 4, r0 + r, src) : nullptr;
        }
        if (p) v = __ldg(reinterpret_cast<const float4*>(p));
      }
      rb[it] = v;
    }
  };

  auto stash = [&](int buf) {
#pragma unroll
    for (int it = 0; it < QA; ++it) {
      const int q = tid + it * NT;
      if (q < BM * BK / 4) {
        if (AOp::kAlongR) {
          const int i = q / (BK / 4), rq = q % (BK / 4);
          As[buf][rq * 4 + 0][i] = ra[it].x;
          As[buf][rq * 4 + 1][i] = ra[it].y;
          As[buf][rq * 4 + 2][i] = ra[it].z;
          As[buf][rq * 4 + 3][i] = ra[it].w;
        } else {
          const int r = q / (BM / 4), iq = q % (BM / 4);
          *reinterpret_cast<float4*>(&As[buf][r][iq * 4]) = ra[it];
        }
      }
    }
#pragma unroll
    for (int it = 0; it < QB; ++it) {
      const int q = tid + it * NT;
      if (q < BN * BK / 4) {
        if (BOp::kAlongR) {
          const int i = q / (BK / 4), rq = q % (BK / 4);
          Bs[buf][rq * 4 + 0][i] = rb[it].x;
          Bs[buf][rq * 4 + 1][i] = rb[it].y;
          Bs[buf][rq * 4 + 2][i] = rb[it].z;
          Bs[buf][rq * 4 + 3][i] = rb[it].w;
        } else {
          const int r = q / (BN / 4), iq = q % (BN / 4);
          *reinterpret_cast<float4*>(&Bs[buf][r][iq * 4]) = rb[it];
        }
      }
    }
  };

  if (ntiles > 0) {
    fetch(0);
    stash(0);
  }
  __syncthreads();

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) fetch(t + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float af[TM], bf[TN];
#pragma unroll
      for (int v = 0; v < MV; ++v) {
        const float4 x = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + v * 4]);
        af[v * 4 + 0] = x.x; af[v * 4 + 1] = x.y; af[v * 4 + 2] = x.z; af[v * 4 + 3] = x.w;
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 x = *reinterpret_cast<const float4*>(&Bs[buf][kk][v * (BN / NV) + tx * 4]);
        bf[v * 4 + 0] = x.x; bf[v * 4 + 1] = x.y; bf[v * 4 + 2] = x.z; bf[v * 4 + 3] = x.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    if (t + 1 < ntiles) stash(buf ^ 1);
    if ((t % kFlush) == kFlush - 1) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] += tot[i][j];

  float part = 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 o = make_float4(acc[i][v * 4], acc[i][v * 4 + 1], acc[i][v * 4 + 2], acc[i][v * 4 + 3]);
      part += epi.store(m0 + ty * TM + i, c0 + v * (BN / NV) + tx * 4, o, split, src);
    }

  if (Epi::kReduce) {
    __shared__ double red[NT / 32];
    double s = warp_sum((double)part);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid < 32) {
      double v = (tid < NT / 32) ? red[tid] : 0.0;
      v = warp_sum(v);
      if (tid == 0)
        epi.block_sum(v, ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    }
  }
}

// ------------------------------------------------------------------------
// Operand / epilogue functors over the time-major device arrays
//   Ht : (h + RT) x Kp, row index tau + h      (tau = local time, h = L-1)
//   Xt, Et : RT x Np
//   W  : L x Np x Kp
// ------------------------------------------------------------------------

// K1 reconstruction  Et[tau][n] = sum_{l,k} Ht[tau-l][k] * W[l][n][k]
// (reference cmf_predict, cmfpy/common.py:50-58; r = l*Kp + k)
struct ReconA {
  static constexpr bool kAlongR = true;
  const float* Ht; int Kp, h;
  __device__ const float* ptr(long long tau, long long r, int) const {
    const int l = (int)(r / Kp), k = (int)(r % Kp);
    return Ht + (tau - l + h) * Kp + k;
  }
};
struct ReconB {
  static constexpr bool kAlongR = true;
  const float* W; int Np, Kp;
  __device__ const float* ptr(long long n, long long r, int) const {
    if (n >= Np) return nullptr;
    const int l = (int)(r / Kp), k = (int)(r % Kp);
    return W + ((long long)l * Np + n) * Kp + k;
  }
};
// writes est (zero past the valid range) and accumulates sum (est - X)^2 over
// owned rows (reference cache_resids + loss, cmfpy/algs/base.py:57-62, 90-97)
struct ReconEpi {
  static constexpr bool kReduce = true;
  float* Et; const float* Xt; double* block_partials;
  int Np; long long t_own, t_valid; int round_out;
  __device__ float store(long long tau, long long n, float4 v, int, int) const {
    if (n >= Np) return 0.f;
    if (tau >= t_valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
    if (tau < t_own) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(Xt + tau * Np + n));
      const float d0 = v.x - x.x, d1 = v.y - x.y, d2 = v.z - x.z, d3 = v.w - x.w;
      s = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    *reinterpret_cast<float4*>(Et + tau * Np + n) = v;
    return s;
  }
  __device__ void block_sum(double s, long long b) const { block_partials[b] = s; }
};

// K2 W terms  out[src][l][n][k] = sum_tau S[tau][n] * Ht[tau-l][k],  S = Xt | Et
// (reference _compute_mult_W, cmfpy/algs/mult.py:27-40; c = l*Kp + k)
struct WTermsA {
  static constexpr bool kAlongR = false;
  const float* Xt; const float* Et; int Np;
  __device__ const float* ptr(long long n, long long tau, int src) const {
    if (n >= Np) return nullptr;
    return (src ? Et : Xt) + tau * Np + n;
  }
};
struct WTermsB {
  static constexpr bool kAlongR = false;
  const float* Ht; int Kp, h, LKp;
  __device__ const float* ptr(long long c, long long tau, int) const {
    if (c >= LKp) return nullptr;
    const int l = (int)(c / Kp), k = (int)(c % Kp);
    return Ht + (tau - l + h) * Kp + k;
  }
};
struct WTermsEpi {
  static constexpr bool kReduce = false;
  float* part; int Np, Kp, LKp; long long per_src;   // per_src = L*Np*Kp
  __device__ float store(long long n, long long c, float4 v, int split, int src) const {
    if (n >= Np || c >= LKp) return 0.f;
    const int l = (int)(c / Kp), k = (int)(c % Kp);
    float* o = part + ((long long)split * 2 + src) * per_src + ((long long)l * Np + n) * Kp + k;
    *reinterpret_cast<float4*>(o) = v;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};

// K3 H terms  out[src][tau][k] = sum_{l,n} S[tau+l][n] * W[l][n][k],  S = Xt | Et
// (reference tensor_transconv, cmfpy/common.py:61-86; r = l*Np + n)
struct HTermsA {
  static constexpr bool kAlongR = true;
  const float* Xt; const float* Et; int Np;
  __device__ const float* ptr(long long tau, long long r, int src) const {
    const int l = (int)(r / Np), n = (int)(r % Np);
    return (src ? Et : Xt) + (tau + l) * Np + n;
  }
};
struct HTermsB {
  static constexpr bool kAlongR = false;
  const float* W; int Kp;
  __device__ const float* ptr(long long k, long long r, int) const {
    if (k >= Kp) return nullptr;
    return W + r * Kp + k;
  }
};
struct HTermsEpi {
  static constexpr bool kReduce = false;
  float* out; int Kp; long long per_src;              // per_src = TO*Kp
  __device__ float store(long long tau, long long k, float4 v, int, int src) const {
    if (k >= Kp) return 0.f;
    *reinterpret_cast<float4*>(out + (long long)src * per_src + tau * Kp + k) = v;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};

// ---- tail corrections of the Gram route (rows of est past the end of the data) ----
// plain store out[m][c]
struct PlainEpi {
  static constexpr bool kReduce = false;
  float* out; long long ld; long long m_rows; int c_cols;
  __device__ float store(long long m, long long c, float4 v, int, int) const {
    if (m < m_rows && c < c_cols) *reinterpret_cast<float4*>(out + m * ld + c) = v;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};
// A(m, r=(l,n)) = Etail[m + m_off + l - t0][n] when that row exists (rows of the untruncated est at t >= t0)
struct TailHA {
  static constexpr bool kAlongR = true;
  const float* Etail; int Np; long long m_off, t0, ntail;
  __device__ const float* ptr(long long m, long long r, int) const {
    const int l = (int)(r / Np), n = (int)(r % Np);
    const long long row = m + m_off + l - t0;
    if (row < 0 || row >= ntail) return nullptr;
    return Etail + row * Np + n;
  }
};
// part[split][m][c] = v   (split-K partials; summed by tail_sub_kernel)
struct PartEpi {
  static constexpr bool kReduce = false;
  float* part; long long ld, m_rows; int c_cols; long long per_split;
  __device__ float store(long long m, long long c, float4 v, int split, int) const {
    if (m < m_rows && c < c_cols) *reinterpret_cast<float4*>(part + split * per_split + m * ld + c) = v;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};
// out[(m + m_off)][c] -= sum_split part[split][m][c]
// One warp per float4 of the output: the lanes stride over the splits and a shuffle tree adds them (fixed
// order => deterministic); a serial loop over a few hundred splits per thread was latency-bound.
__global__ void __launch_bounds__(256)
tail_sub_kernel(float* __restrict__ out, const float* __restrict__ part, int nsplit, long long per_split, long long m_rows,
                int ld, long long m_off, long long row_end) {
  const long long total4 = m_rows * ld / 4;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i4 = warp0; i4 < total4; i4 += nwarps) {
    const long long i = i4 * 4;
    const long long row = i / ld + m_off;
    if (row < 0 || row >= row_end) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = lane; sp < nsplit; sp += 32) {
      const float4 v = *reinterpret_cast<const float4*>(part + (size_t)sp * per_split + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (lane == 0) {
      float4* o = reinterpret_cast<float4*>(out + row * ld + (i % ld));
      float4 c = *o;
      c.x -= acc.x; c.y -= acc.y; c.z -= acc.z; c.w -= acc.w;
      *o = c;
    }
  }
}
// out[(m + m_off)][k] -= v    (rows m + m_off < row_end)
struct SubEpi {
  static constexpr bool kReduce = false;
  float* out; int Kp; long long m_off, row_end;
  __device__ float store(long long m, long long k, float4 v, int, int) const {
    const long long row = m + m_off;
    if (k >= Kp || row < 0 || row >= row_end) return 0.f;
    float4* o = reinterpret_cast<float4*>(out + row * Kp + k);
    float4 c = *o;
    c.x -= v.x; c.y -= v.y; c.z -= v.z; c.w -= v.w;
    *o = c;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};

// out[l][n][k] -= v   with c = l*Kp + k   (W-layout subtract, tail correction of the W step)
struct SubWEpi {
  static constexpr bool kReduce = false;
  float* out; int Np, Kp, LKp;
  __device__ float store(long long n, long long c, float4 v, int, int) const {
    if (n >= Np || c >= LKp) return 0.f;
    const int l = (int)(c / Kp), k = (int)(c % Kp);
    float4* o = reinterpret_cast<float4*>(out + ((long long)l * Np + n) * Kp + k);
    float4 x = *o;
    x.x -= v.x; x.y -= v.y; x.z -= v.z; x.w -= v.w;
    *o = x;
    return 0.f;
  }
  __device__ void block_sum(double, long long) const {}
};

}  // namespace simt
}  // namespace cmf
