// Host side of the tcgen05 path: folding parameters, tensor maps, scratch
// buffers, launches.
#pragma once
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "ew_kernels.cuh"
#include "tc_kernels.cuh"

namespace cmf {
namespace tc {

struct Dims {
  int N, K, L, Np, Kp, h;
  long long Tloc, TO, RT, RH, t_valid;
  int num_sms;
};

// Padded component count of the tensor-core path: 8, 16, 32, 64 or 128 (0 = unsupported).
inline int padded_k(int K) {
  if (K <= 8) return 8;
  if (K <= 16) return 16;
  if (K <= 32) return 32;
  if (K <= 64) return 64;
  if (K <= 128) return 128;
  return 0;
}

// How Kp is brought to 128-byte operand rows (see tc_kernels.cuh).
struct Fold {
  int Kp = 32, s = 1, CB = 1, Lv = 1, KW = 32;   // lag stride, column blocks, virtual lags, virtual row width
  int J = 1, n_glag = 4;                          // H terms: virtual lags per lag group, lag groups (4 / CB)
  int recon_wrows = 320, hterms_wrows = 288;
};

inline Fold make_fold(int Kp, int L) {
  Fold f;
  f.Kp = Kp;
  f.s = Kp < 32 ? 32 / Kp : 1;
  f.CB = Kp > 32 ? Kp / 32 : 1;
  f.Lv = (L + f.s - 1) / f.s;
  f.KW = 32 * f.CB;
  f.n_glag = 4 / f.CB;
  f.J = (f.Lv + f.n_glag - 1) / f.n_glag;
  f.recon_wrows = round_up(256 + f.s * (f.Lv - 1), 64);
  f.hterms_wrows = round_up(256 + f.s * (f.J - 1), 32);
  return f;
}

constexpr size_t kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100

inline bool shape_supported(int N, int K, int L) {
  const int Kp = padded_k(K);
  if (N < 1 || K < 1 || L < 1 || Kp == 0) return false;
  const Fold f = make_fold(Kp, L);
  return recon_smem_bytes(f.recon_wrows) <= kMaxSmem && hterms_smem_bytes(f.hterms_wrows) <= kMaxSmem &&
         wterms_smem_bytes(f.s) <= kMaxSmem;
}

struct TcState {
  bool ready = false;
  Dims d{};
  Fold f{};
  int mask = 7;                    // bit0 recon, bit1 w terms, bit2 h terms run on tensor cores
  float *Xt = nullptr, *Et = nullptr, *Ht = nullptr, *W = nullptr, *numden = nullptr, *hterms = nullptr;  // masters
  float *Wv = nullptr, *Hv = nullptr;   // TF32-rounded (and, for Kp < 32, lag-folded) operand copies
  double *loss_partials = nullptr, *d_sumsq = nullptr;
  float* wpart = nullptr;
  float* hscratch = nullptr;       // [2][4][32][TO + 256] lag-group partials of the H terms
  int* d_err = nullptr;
  int n_chunks = 1, n_lag_groups = 1;
  int recon_grid = 1, wterms_grid = 1, hterms_grid = 1;
  long long wcount = 0, wv_count = 0, hv_count = 0;
  CUtensorMap tmW_k1, tmH_k1, tmX_k2, tmE_k2, tmH_k2, tmW_k3, tmX_k3, tmE_k3;
};

constexpr int kReconLaunches = 2, kWTermsLaunches = 2, kHTermsLaunches = 2;   // kernels per phase

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 map over a row-major [rows][cols] array (row pitch = cols floats)
inline int make_map(CUtensorMap* m, const float* base, long long rows, long long cols, int box_cols, int box_rows,
                    CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  CMF_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CMF_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for rows=%lld cols=%lld box=%dx%d", (int)r, rows,
            cols, box_cols, box_rows);
  return 0;
}

inline void destroy(TcState& s) {
  cudaFree(s.wpart); cudaFree(s.hscratch); cudaFree(s.d_err); cudaFree(s.Wv); cudaFree(s.Hv);
  s.wpart = s.hscratch = s.Wv = s.Hv = nullptr;
  s.d_err = nullptr;
  s.ready = false;
}

inline int launch_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

inline int ew_blocks(const TcState& s, long long items) {
  long long b = ceil_div_ll(items, 256);
  if (b > s.d.num_sms * 8ll) b = s.d.num_sms * 8ll;
  return (int)(b < 1 ? 1 : b);
}

// When no folding is needed the MU update kernels write the rounded copy
// themselves (fused); these return the pointers they should write, or null.
inline float* fused_w_op(TcState& s) { return (s.ready && s.f.s == 1) ? s.Wv : nullptr; }
inline float* fused_h_op(TcState& s) { return (s.ready && s.f.s == 1) ? s.Hv : nullptr; }

// Rebuild the operand copies from the fp32 masters.
inline int refresh_w(TcState& s, cudaStream_t stream, bool after_fused_update = false) {
  if (!s.ready) return 0;
  const Dims& d = s.d;
  if (s.f.s == 1) {
    if (after_fused_update) return 0;
    ew::round_copy_kernel<<<ew_blocks(s, s.wcount / 4), 256, 0, stream>>>((float4*)s.Wv, (const float4*)s.W, s.wcount / 4);
    return launch_ok("round_w");
  }
  fold_w_kernel<<<ew_blocks(s, s.wv_count), 256, 0, stream>>>(s.Wv, s.W, d.L, s.f.Lv, d.Np, d.Kp, s.f.s);
  return launch_ok("fold_w");
}
inline int refresh_h(TcState& s, cudaStream_t stream, long long row0, long long nrows, bool after_fused_update = false) {
  if (!s.ready || nrows <= 0) return 0;
  const Dims& d = s.d;
  if (s.f.s == 1) {
    if (after_fused_update) return 0;
    const long long n4 = nrows * d.Kp / 4;
    ew::round_copy_kernel<<<ew_blocks(s, n4), 256, 0, stream>>>((float4*)(s.Hv + row0 * d.Kp), (const float4*)(s.Ht + row0 * d.Kp), n4);
    return launch_ok("round_h");
  }
  // a folded row depends on the s-1 rows before it
  long long r1 = row0 + nrows + s.f.s - 1;
  if (r1 > d.RH) r1 = d.RH;
  fold_h_kernel<<<ew_blocks(s, (r1 - row0) * 32), 256, 0, stream>>>(s.Hv, s.Ht, row0, r1 - row0, d.Kp);
  return launch_ok("fold_h");
}

inline int init(TcState& s, const Dims& d, float* Xt, float* Et, float* Ht, float* W, float* numden, float* hterms,
                double* loss_partials, long long n_loss_partials, double* d_sumsq, cudaStream_t stream) {
  s.d = d;
  s.f = make_fold(d.Kp, d.L);
  const Fold& f = s.f;
  s.Xt = Xt; s.Et = Et; s.Ht = Ht; s.W = W; s.numden = numden; s.hterms = hterms;
  s.loss_partials = loss_partials; s.d_sumsq = d_sumsq;
  s.wcount = (long long)d.L * d.Np * d.Kp;
  s.wv_count = (long long)f.Lv * d.Np * f.KW;
  s.hv_count = d.RH * f.KW;
  if (const char* e = getenv("CMF_TC_MASK")) s.mask = atoi(e);
  CMF_CHECK(d.Kp == padded_k(d.K) && shape_supported(d.N, d.K, d.L), "shape not supported by the tensor-core path");
  CMF_CHECK(n_loss_partials >= d.num_sms, "loss partial buffer too small");

  CMF_CUDA(cudaMalloc((void**)&s.d_err, 4));
  CMF_CUDA(cudaMemsetAsync(s.d_err, 0, 4, stream));
  CMF_CUDA(cudaMalloc((void**)&s.Wv, (size_t)s.wv_count * 4));
  CMF_CUDA(cudaMalloc((void**)&s.Hv, (size_t)s.hv_count * 4));
  CMF_CUDA(cudaMemsetAsync(s.Wv, 0, (size_t)s.wv_count * 4, stream));
  CMF_CUDA(cudaMemsetAsync(s.Hv, 0, (size_t)s.hv_count * 4, stream));

  // ---- K1 -------------------------------------------------------------
  {
    const long long tiles = ceil_div_ll(d.Np, 128) * (d.RT / 256);
    s.recon_grid = (int)(tiles < d.num_sms ? tiles : d.num_sms);
  }
  CMF_TRY(make_map(&s.tmW_k1, s.Wv, (long long)f.Lv * d.Np, f.KW, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_TRY(make_map(&s.tmH_k1, s.Hv, d.RH, f.KW, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_CUDA(cudaFuncSetAttribute(tc_recon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)recon_smem_bytes(f.recon_wrows)));

  // ---- K2 -------------------------------------------------------------
  s.n_lag_groups = (int)ceil_div_ll(f.Lv, 16);
  {
    const long long units = ceil_div_ll(d.Np, 128) * s.n_lag_groups * f.CB * 2;
    const long long stages_total = ceil_div_ll(d.Tloc, 32);
    long long cmax = stages_total / 4; if (cmax < 1) cmax = 1;
    const long long cmem = (2ll << 30) / (2 * s.wcount * 4); if (cmem < cmax) cmax = cmem < 1 ? 1 : cmem;
    if (cmax > 256) cmax = 256;
    // number of time chunks: fill every SM with equal work (whole waves)
    double best = -1.0; int bestc = 1;
    for (long long c = 1; c <= cmax; ++c) {
      const long long items = units * c;
      const long long waves = ceil_div_ll(items, d.num_sms);
      double eff = (double)items / (double)(waves * d.num_sms);
      if (items >= 2ll * d.num_sms) eff += 1e-3;         // prefer at least two items per SM
      if (eff > best + 1e-9) { best = eff; bestc = (int)c; }
    }
    s.n_chunks = bestc;
    const long long items = units * s.n_chunks;
    s.wterms_grid = (int)(items < d.num_sms ? items : d.num_sms);
  }
  CMF_CUDA(cudaMalloc((void**)&s.wpart, (size_t)s.n_chunks * 2 * s.wcount * 4));
  CMF_TRY(make_map(&s.tmX_k2, Xt, d.Tloc, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_TRY(make_map(&s.tmE_k2, Et, d.Tloc, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_TRY(make_map(&s.tmH_k2, s.Hv, d.RH, f.KW, 32, wterms_brows(f.s), CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_CUDA(cudaFuncSetAttribute(tc_wterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wterms_smem_bytes(f.s)));

  // ---- K3 -------------------------------------------------------------
  {
    const long long tiles = d.TO / 256 + 1;
    s.hterms_grid = (int)(tiles < d.num_sms ? tiles : d.num_sms);
  }
  CMF_TRY(make_map(&s.tmW_k3, s.Wv, (long long)f.Lv * d.Np, f.KW, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_TRY(make_map(&s.tmX_k3, Xt, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_TRY(make_map(&s.tmE_k3, Et, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_CUDA(cudaMalloc((void**)&s.hscratch, (size_t)2 * 4 * kKp * (d.TO + 256) * 4));
  CMF_CUDA(cudaMemsetAsync(s.hscratch, 0, (size_t)2 * 4 * kKp * (d.TO + 256) * 4, stream));
  CMF_CUDA(cudaFuncSetAttribute(tc_hterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)hterms_smem_bytes(f.hterms_wrows)));
  if (d.Kp * 33 * 4 > 48 * 1024)
    CMF_CUDA(cudaFuncSetAttribute(combine_groups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d.Kp * 33 * 4));
  s.ready = true;
  return 0;
}

inline int recon(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  ReconParams p;
  p.Np = d.Np; p.L = f.Lv; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.wrows = f.recon_wrows;
  p.s = f.s; p.CB = f.CB; p.h_shift = d.h - f.s * (f.Lv - 1);
  p.n_tiles = (long long)p.n_tiles_n * (d.RT / 256);
  p.t_own = d.Tloc; p.t_valid = d.t_valid;
  p.Et = s.Et; p.Xt = s.Xt; p.loss_partials = s.loss_partials; p.round_out = 1; p.err = s.d_err;
  tc_recon_kernel<<<s.recon_grid, kReconThreads, recon_smem_bytes(f.recon_wrows), stream>>>(s.tmW_k1, s.tmH_k1, p);
  CMF_TRY(launch_ok("tc_recon"));
  ew::sum_doubles_kernel<<<1, 1024, 0, stream>>>(s.loss_partials, s.recon_grid, s.d_sumsq);
  return launch_ok("loss_sum");
}

inline int w_terms(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  WTermsParams p;
  p.Np = d.Np; p.L = d.L; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.n_lag_groups = s.n_lag_groups;
  p.n_chunks = s.n_chunks; p.h = d.h;
  p.Lv = f.Lv; p.Kp = d.Kp; p.s = f.s; p.CB = f.CB; p.brows = wterms_brows(f.s);
  p.n_items = (long long)p.n_tiles_n * p.n_lag_groups * f.CB * 2 * p.n_chunks;
  p.stages_total = ceil_div_ll(d.Tloc, 32);
  p.part = (s.n_chunks == 1) ? s.numden : s.wpart;
  p.per_src = s.wcount; p.err = s.d_err;
  tc_wterms_kernel<<<s.wterms_grid, kWtThreads, wterms_smem_bytes(f.s), stream>>>(s.tmX_k2, s.tmE_k2, s.tmH_k2, p);
  CMF_TRY(launch_ok("tc_wterms"));
  if (s.n_chunks > 1) {
    const long long n4 = 2 * s.wcount / 4;
    ew::sum_splits_kernel<<<ew_blocks(s, n4), 256, 0, stream>>>((float4*)s.numden, (const float4*)s.wpart, n4, n4, s.n_chunks);
    CMF_TRY(launch_ok("w_terms_sum"));
  }
  return 0;
}

inline int h_terms(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  HTermsParams p;
  p.Np = d.Np; p.J = f.J; p.n_chunks_n = (int)ceil_div_ll(d.Np, 32); p.wrows = f.hterms_wrows;
  p.s = f.s; p.CB = f.CB;
  p.n_tiles = d.TO / 256 + 1; p.ts = d.TO + 256; p.scratch = s.hscratch; p.err = s.d_err;
  tc_hterms_kernel<<<s.hterms_grid, kHtThreads, hterms_smem_bytes(f.hterms_wrows), stream>>>(s.tmW_k3, s.tmX_k3, s.tmE_k3, p);
  CMF_TRY(launch_ok("tc_hterms"));
  combine_groups_kernel<<<(unsigned)(d.TO / 32), 256, d.Kp * 33 * 4, stream>>>(s.hscratch, s.hterms, p.ts, d.TO, f.J,
                                                                            f.s, f.CB, d.Kp);
  return launch_ok("combine_groups");
}

// Device-side pipeline errors (bounded waits that expired) surface here.
inline int check(TcState& s, cudaStream_t stream) {
  if (!s.ready) return 0;
  int e = 0;
  CMF_CUDA(cudaMemcpyAsync(&e, s.d_err, 4, cudaMemcpyDeviceToHost, stream));
  CMF_CUDA(cudaStreamSynchronize(stream));
  CMF_CHECK(e == 0, "tensor-core kernel pipeline error %d (a barrier wait timed out)", e);
  return 0;
}

}  // namespace tc
}  // namespace cmf
