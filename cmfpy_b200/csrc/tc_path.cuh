// Host side of the tcgen05 path: tensor maps, scratch buffers, launches.
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

struct TcState {
  bool ready = false;
  Dims d{};
  int mask = 7;                    // bit0 recon, bit1 w terms, bit2 h terms run on tensor cores
  float *Xt = nullptr, *Et = nullptr, *Ht = nullptr, *W = nullptr, *numden = nullptr, *hterms = nullptr;
  double *loss_partials = nullptr, *d_sumsq = nullptr;
  float* wpart = nullptr;
  float* hscratch = nullptr;        // [2][4][Kp][TO + 256] lag-group partials of the H terms
  int* d_err = nullptr;
  int n_chunks = 1, n_lag_groups = 1, J = 1;
  int recon_wrows = 320, hterms_wrows = 288;
  int recon_grid = 1, wterms_grid = 1, hterms_grid = 1;
  long long wcount = 0;
  CUtensorMap tmW_k1, tmH_k1, tmX_k2, tmE_k2, tmH_k2, tmW_k3, tmX_k3, tmE_k3;
};

constexpr int kReconLaunches = 2, kWTermsLaunches = 2, kHTermsLaunches = 2;   // kernels per phase

// The tensor-core kernels in tc_kernels.cuh are specialised for one 128-byte
// row of H^T (25 <= K <= 32) and lag windows that fit shared memory.
inline bool shape_supported(int N, int K, int L) {
  return N >= 1 && K > 24 && K <= 32 && L >= 1 && L <= 256;
}

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
  cudaFree(s.wpart);
  cudaFree(s.hscratch);
  s.hscratch = nullptr;
  cudaFree(s.d_err);
  s.wpart = nullptr;
  s.d_err = nullptr;
  s.ready = false;
}

inline int init(TcState& s, const Dims& d, float* Xt, float* Et, float* Ht, float* W, float* numden, float* hterms,
                double* loss_partials, long long n_loss_partials, double* d_sumsq, cudaStream_t stream) {
  s.d = d;
  s.Xt = Xt; s.Et = Et; s.Ht = Ht; s.W = W; s.numden = numden; s.hterms = hterms;
  s.loss_partials = loss_partials; s.d_sumsq = d_sumsq;
  s.wcount = (long long)d.L * d.Np * d.Kp;
  if (const char* e = getenv("CMF_TC_MASK")) s.mask = atoi(e);
  CMF_CHECK(d.Kp == kKp, "tensor-core path needs Kp == 32");
  CMF_CHECK(n_loss_partials >= d.num_sms, "loss partial buffer too small");

  CMF_CUDA(cudaMalloc((void**)&s.d_err, 4));
  CMF_CUDA(cudaMemsetAsync(s.d_err, 0, 4, stream));

  // ---- K1 -------------------------------------------------------------
  s.recon_wrows = round_up(256 + d.L - 1, 64);
  {
    const long long tiles = ceil_div_ll(d.Np, 128) * (d.RT / 256);
    s.recon_grid = (int)(tiles < d.num_sms ? tiles : d.num_sms);
  }
  CMF_TRY(make_map(&s.tmW_k1, W, (long long)d.L * d.Np, d.Kp, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_TRY(make_map(&s.tmH_k1, Ht, d.RH, d.Kp, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_CUDA(cudaFuncSetAttribute(tc_recon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)recon_smem_bytes(s.recon_wrows)));

  // ---- K2 -------------------------------------------------------------
  s.n_lag_groups = (int)ceil_div_ll(d.L, 16);
  {
    const long long units = ceil_div_ll(d.Np, 128) * s.n_lag_groups * 2;
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
  CMF_TRY(make_map(&s.tmH_k2, Ht, d.RH, d.Kp, 32, kWtBRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_CUDA(cudaFuncSetAttribute(tc_wterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wterms_smem_bytes()));

  // ---- K3 -------------------------------------------------------------
  s.J = (int)ceil_div_ll(d.L, 4);
  s.hterms_wrows = round_up(256 + s.J - 1, 32);
  {
    const long long tiles = d.TO / 256 + 1;
    s.hterms_grid = (int)(tiles < d.num_sms ? tiles : d.num_sms);
  }
  CMF_TRY(make_map(&s.tmW_k3, W, (long long)d.L * d.Np, d.Kp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_TRY(make_map(&s.tmX_k3, Xt, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_TRY(make_map(&s.tmE_k3, Et, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_CUDA(cudaMalloc((void**)&s.hscratch, (size_t)2 * 4 * kKp * (d.TO + 256) * 4));
  CMF_CUDA(cudaMemsetAsync(s.hscratch, 0, (size_t)2 * 4 * kKp * (d.TO + 256) * 4, stream));
  CMF_CUDA(cudaFuncSetAttribute(tc_hterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)hterms_smem_bytes(s.hterms_wrows)));
  s.ready = true;
  return 0;
}

inline int launch_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

inline int recon(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  ReconParams p;
  p.Np = d.Np; p.L = d.L; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.wrows = s.recon_wrows;
  p.n_tiles = (long long)p.n_tiles_n * (d.RT / 256);
  p.t_own = d.Tloc; p.t_valid = d.t_valid;
  p.Et = s.Et; p.Xt = s.Xt; p.loss_partials = s.loss_partials; p.round_out = 1; p.err = s.d_err;
  tc_recon_kernel<<<s.recon_grid, kReconThreads, recon_smem_bytes(s.recon_wrows), stream>>>(s.tmW_k1, s.tmH_k1, p);
  CMF_TRY(launch_ok("tc_recon"));
  ew::sum_doubles_kernel<<<1, 1024, 0, stream>>>(s.loss_partials, s.recon_grid, s.d_sumsq);
  return launch_ok("loss_sum");
}

inline int w_terms(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  WTermsParams p;
  p.Np = d.Np; p.L = d.L; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.n_lag_groups = s.n_lag_groups;
  p.n_chunks = s.n_chunks; p.h = d.h;
  p.n_items = (long long)p.n_tiles_n * p.n_lag_groups * 2 * p.n_chunks;
  p.stages_total = ceil_div_ll(d.Tloc, 32);
  p.part = (s.n_chunks == 1) ? s.numden : s.wpart;
  p.per_src = s.wcount; p.err = s.d_err;
  tc_wterms_kernel<<<s.wterms_grid, kWtThreads, wterms_smem_bytes(), stream>>>(s.tmX_k2, s.tmE_k2, s.tmH_k2, p);
  CMF_TRY(launch_ok("tc_wterms"));
  if (s.n_chunks > 1) {
    const long long n4 = 2 * s.wcount / 4;
    long long blocks = ceil_div_ll(n4, 256);
    if (blocks > d.num_sms * 8ll) blocks = d.num_sms * 8ll;
    ew::sum_splits_kernel<<<(int)blocks, 256, 0, stream>>>((float4*)s.numden, (const float4*)s.wpart, n4, n4, s.n_chunks);
    CMF_TRY(launch_ok("w_terms_sum"));
  }
  return 0;
}

inline int h_terms(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  HTermsParams p;
  p.Np = d.Np; p.J = s.J; p.n_chunks_n = (int)ceil_div_ll(d.Np, 32); p.wrows = s.hterms_wrows;
  p.n_tiles = d.TO / 256 + 1; p.ts = d.TO + 256; p.scratch = s.hscratch; p.err = s.d_err;
  tc_hterms_kernel<<<s.hterms_grid, kHtThreads, hterms_smem_bytes(s.hterms_wrows), stream>>>(s.tmW_k3, s.tmX_k3, s.tmE_k3, p);
  CMF_TRY(launch_ok("tc_hterms"));
  combine_groups_kernel<<<(unsigned)(d.TO / 32), 256, 0, stream>>>(s.hscratch, s.hterms, p.ts, d.TO, s.J);
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
