// Host side of the tcgen05 path: folding parameters, tensor maps, scratch
// buffers, launches.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "dev_cache.cuh"
#include "ew_kernels.cuh"
#include "simt_gemm.cuh"
#include "tc_strict_kernels.cuh"

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
  int recon_LB = 0;                               // K1: lags per window (0: all of them in one window)
  int h_hd = 0;                                   // H terms: columns before a time tile that its lag groups reach
  int h_stages = 0;                               // H terms: W ring depth that fits shared memory (0: none does)
  int h_staged = 0;                               // H terms: folded lags reduced through the staged gather
};

constexpr size_t kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100

// W ring depth and reduction form of the H-terms kernel: the staged gather (folded lags, K < 32) is taken when it
// costs no ring stage
inline void choose_hterms_ring(Fold& f, bool direct) {
  f.h_stages = 0;
  f.h_staged = 0;
  static const bool staging_enabled = [] { const char* e = getenv("CMF_HT_STAGED"); return !e || atoi(e) != 0; }();
  for (int st = 3; st >= 2 && !f.h_stages; --st) {
    if (f.s > 1 && staging_enabled && hterms_smem_bytes(st, f.hterms_wrows, f.Kp, f.h_hd, direct, true) <= kMaxSmem) {
      f.h_stages = st; f.h_staged = 1;
    } else if (hterms_smem_bytes(st, f.hterms_wrows, f.Kp, f.h_hd, direct, false) <= kMaxSmem) {
      f.h_stages = st;
    }
  }
}

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
  if (recon_smem_bytes(f.recon_wrows) > kMaxSmem) {
    // the lag range does not fit one window of H^T: blocks of recon_LB lags (tc_recon_kernel, ReconParams::LB)
    const int max_rows = (int)((kMaxSmem - recon_smem_bytes(0)) / (2 * kKp * 4) / 64) * 64;
    f.recon_LB = ((max_rows - 256) / f.s + 1) & ~1;
    f.recon_wrows = round_up(256 + f.s * (f.recon_LB - 1), 64);
  }
  f.hterms_wrows = round_up(256 + f.s * (f.J - 1), 32);
  f.h_hd = (f.n_glag - 1) * f.s * f.J + f.s - 1;
  const bool direct = f.n_glag == 1 && f.s == 1;
  choose_hterms_ring(f, direct);
  return f;
}

inline bool shape_supported(int N, int K, int L) {
  const int Kp = padded_k(K);
  if (N < 1 || K < 1 || L < 1 || Kp == 0) return false;
  const Fold f = make_fold(Kp, L);
  return recon_smem_bytes(f.recon_wrows) <= kMaxSmem && f.h_stages > 0 && wterms_smem_bytes(f.s) <= kMaxSmem;
}

struct TcState {
  bool ready = false;
  Dims d{};
  Fold f{};
  int mask = 7;                    // bit0 recon, bit1 w terms, bit2 h terms run on tensor cores
  float *Xt = nullptr, *Et = nullptr, *Ht = nullptr, *W = nullptr, *numden = nullptr, *hterms = nullptr;  // masters
  float *Wv = nullptr, *Hv = nullptr;   // TF32-rounded (and, for Kp < 32, lag-folded) operand copies
  // 3xTF32 (CMF_PREC_TF32X3): every operand is a TF32 pair.  Wv / Hv rows are [hi (KW) | lo (KW)] (KWs = 2 KW
  // columns); Xt / Et hold the hi halves and Xlo / Elo the lo halves.
  int x3 = 0, KWs = 32;
  int loss_fast = 0;               // 3xTF32, loss-only reconstruction: the hi x hi operand pass alone (see recon())
  float *Xlo = nullptr, *Elo = nullptr;
  CUtensorMap tmXlo_k2, tmElo_k2, tmXlo_k3, tmElo_k3;
  double *loss_partials = nullptr, *d_sumsq = nullptr;
  float* wpart = nullptr;
  float* hcarry = nullptr;         // [2][time tiles][h_hd][Kp]: the part of each K3 tile that belongs to the tile before it
  float* hparts = nullptr;         // K3 with split items: [n_src][h_split][TO][Kp] partial outputs
  int h_split = 1;                 // K3: items per (time tile, source) (HTermsParams::n_split)
  int* d_err = nullptr;
  int n_chunks = 1, n_lag_groups = 1;    // K2: time chunks; lag groups of 16 (tf32) or 8 (3xTF32) virtual lags
  int h_sub = 0;                         // K3: units per tensor-memory sub-chunk (0: one chain per item)
  int recon_grid = 1, wterms_grid = 1, hterms_grid = 1;
  long long wcount = 0, wv_count = 0, hv_count = 0;
  CUtensorMap tmW_k1, tmH_k1, tmX_k2, tmE_k2, tmH_k2, tmW_k3, tmX_k3, tmE_k3;

  // ---- Gram route for the denominators (gram & 1: H step, gram & 2: W step) ----
  int gram = 0;
  int gram_request = 0;             // from cmf_mu_params.denominators (CMF_GRAM in the environment overrides)
  int LK = 0, Lr = 0, Lrv = 0;
  // den_H = R (*) H runs on the H-terms kernel with R in the role of W, H^T in the role of the data and the
  // K components in the role of the features (Fold of that problem: J, head columns, window rows, ring depth)
  Fold fr{};
  float* hcarry_r = nullptr;
  CUtensorMap tmRw_k3, tmHs_k3, tmHslo_k3;
  int NpA = 0;                      // row half-width of Wt (Np rounded up to 32: the lo half starts on a TMA box boundary)
  long long g_rows = 0;             // rows allocated for G (LK rounded up to 256)
  float *Wt = nullptr, *G = nullptr, *Rw = nullptr, *Rwv = nullptr, *Etail = nullptr;
  long long ntail = 0;              // rows of est past the end of the data that fall inside this shard's window
  CUtensorMap tmWt_a, tmWt_b;
  // W step
  float *P = nullptr, *Ppart = nullptr, *Mt = nullptr;
  float* Pw = nullptr;             // the lag autocorrelation of H as the shift operand of den_W (see den_w_gram)
  long long pw_rows = 0;
  int pw_LB = 0, pw_wrows = 0, pw_srows = 32;
  CUtensorMap tmPw_b;
  int p_chunks = 1, p_grid = 1;
  CUtensorMap tmHx_k2, tmHxlo_k2, tmMt_b;
  int p_quad = 0, p_lag_groups = 1;     // autocorrelation pass: quad mode (Kp <= 32), its lag groups
  long long p_stage0 = 0;
  bool autocorr_ready = false;
};


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
                    CUtensorMapSwizzle swz, long long pitch_cols = 0) {
  if (pitch_cols <= 0) pitch_cols = cols;
  EncodeTiledFn enc = get_encode();
  CMF_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CMF_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for rows=%lld cols=%lld box=%dx%d", (int)r, rows,
            cols, box_cols, box_rows);
  return 0;
}

// fp32 map of rank <= 5: dims (innermost first; dim 0 contiguous), byte strides of dims 1.., box
inline int make_map_nd(CUtensorMap* m, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  CMF_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CMF_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a rank-%d map", (int)r, rank);
  return 0;
}

// K2 data operand: a 32-row stage of S^T (rows x cols, row pitch `pitch` floats) as ONE box of four 32-feature
// regions: dims (feature in block, time row, 32-feature block) -> shared memory [block][row][32 features].
// (The last block of a ragged feature count reads on into the next row: those accumulator rows are never stored.)
inline int make_map_k2src(CUtensorMap* m, const float* base, long long rows, long long cols, long long pitch) {
  cuuint64_t dims[3] = {(cuuint64_t)(cols < 32 ? cols : 32), (cuuint64_t)rows, (cuuint64_t)ceil_div_ll(cols, 32)};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, 128};
  cuuint32_t box[3] = {32, 32, 4};
  return make_map_nd(m, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

// K3 motif operand: both lags of a stage and all four (lag group, column block) regions as ONE box:
// dims (k in block, feature, 32-column block, lag group, lag in group) -> shared memory [lag][region][32 n][32 k].
// Wv must be allocated for n_glag * J lags (zeros past Lv).
inline int make_map_k3w(CUtensorMap* m, const float* Wv, const Fold& f, long long Np, long long KWs) {   // f.J, f.n_glag, f.CB
  cuuint64_t dims[5] = {32, (cuuint64_t)Np, (cuuint64_t)(KWs / 32), (cuuint64_t)f.n_glag, (cuuint64_t)f.J};
  cuuint64_t strides[4] = {(cuuint64_t)KWs * 4, 128, (cuuint64_t)f.J * Np * KWs * 4, (cuuint64_t)Np * KWs * 4};
  cuuint32_t box[5] = {32, 32, (cuuint32_t)f.CB, (cuuint32_t)f.n_glag, (cuuint32_t)kHtLagsPerStage};
  return make_map_nd(m, Wv, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

inline void destroy(TcState& s) {
  cached_free(s.wpart); cached_free(s.hcarry); cached_free(s.hparts); cached_free(s.d_err); cached_free(s.Wv); cached_free(s.Hv);
  cached_free(s.Wt); cached_free(s.G); cached_free(s.Rw); cached_free(s.Rwv); cached_free(s.Etail); cached_free(s.hcarry_r);
  cached_free(s.P); cached_free(s.Ppart); cached_free(s.Mt); cached_free(s.Pw);
  s.Wt = s.G = s.Rw = s.Rwv = s.Etail = s.P = s.Ppart = s.Mt = s.hcarry_r = s.Pw = nullptr;
  s.wpart = s.hcarry = s.hparts = s.Wv = s.Hv = nullptr;
  s.d_err = nullptr;
  s.ready = false;
}

// every kernel launch of this path goes through launch_ok(); the ABI reads the counter for
// cmf_mu_launch_count()
inline long long& launch_counter() {
  static thread_local long long n = 0;
  return n;
}

// Per-launch device times (cmf_mu_set_profiling(h, 2)): while a log is installed, every launch that goes through
// launch_ok() or the ABI's launch_check() is followed by an event on the solver's stream; consecutive events
// bracket one kernel (all launches of a solver are on one stream).
struct LaunchLog {
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> pool;
  std::vector<const char*> labels;         // labels[i]: the launch that ends at event i (labels[0]: the start mark)
};
inline LaunchLog*& launch_log() {
  static thread_local LaunchLog* p = nullptr;
  return p;
}
inline void log_launch(const char* what) {
  LaunchLog* l = launch_log();
  if (!l) return;
  if (l->labels.size() == l->pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    l->pool.push_back(e);
  }
  cudaEventRecord(l->pool[l->labels.size()], l->stream);
  l->labels.push_back(what);
}

inline int launch_ok(const char* what) {
  ++launch_counter();
  log_launch(what);
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
inline float* fused_w_op(TcState& s) { return (s.ready && s.f.s == 1 && !s.x3) ? s.Wv : nullptr; }
inline float* fused_h_op(TcState& s) { return (s.ready && s.f.s == 1 && !s.x3) ? s.Hv : nullptr; }

// Rebuild the operand copies from the fp32 masters.
inline int refresh_w(TcState& s, cudaStream_t stream, bool after_fused_update = false) {
  if (!s.ready) return 0;
  const Dims& d = s.d;
  if (s.x3) {
    fold_w_x3_kernel<<<ew_blocks(s, s.wv_count), 256, 0, stream>>>(s.Wv, s.W, d.L, s.f.Lv, d.Np, d.Kp, s.f.s, s.f.KW);
    return launch_ok("split_w");
  }
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
  if (s.x3) {
    long long r1 = row0 + nrows + s.f.s - 1;       // a folded row depends on the s-1 rows before it
    if (r1 > d.RH) r1 = d.RH;
    fold_h_x3_kernel<<<ew_blocks(s, (r1 - row0) * s.f.KW), 256, 0, stream>>>(s.Hv, s.Ht, row0, r1 - row0, d.Kp, s.f.s, s.f.KW);
    return launch_ok("split_h");
  }
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

// units per tensor-memory sub-chunk of the 3xTF32 kernels (CMF_X3_SUB overrides; 0: one chain per work item)
inline int strict_sub_units() {
  static const int v = [] { const char* e = getenv("CMF_X3_SUB"); return e ? atoi(e) : kStrictSubUnits; }();
  return v;
}

// 3xTF32 bookkeeping of a launch on the recon kernel: p.CB holds the reduction blocks of ONE operand pass on entry
inline void set_x3(const TcState& s, ReconParams& p, int lo_a, int lo_b) {
  if (!s.x3) return;
  p.x3 = 1; p.cbx = p.CB; p.CB = 3 * p.cbx; p.lo_off = lo_a; p.lo_off_b = lo_b;
  p.sub_units = strict_sub_units();
}
// 3xTF32 runs the two-level-accumulation kernel (tc_strict_kernels.cuh), plain TF32 the single-chain one
inline void launch_recon(const TcState& s, int grid, size_t smem, cudaStream_t stream, const CUtensorMap& a,
                         const CUtensorMap& b, const ReconParams& p) {
  static const bool force = [] { const char* e = getenv("CMF_FORCE_STRICT_K1"); return e && atoi(e); }();   // A/B timing only
  if (s.x3 || force) tc_recon_x3_kernel<<<grid, kSThreads, smem, stream>>>(a, b, p);
  else tc_recon_kernel<<<grid, kReconThreads, smem, stream>>>(a, b, p);
}
inline int set_recon_smem(size_t bytes) {
  CMF_CUDA(cudaFuncSetAttribute(tc_recon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  CMF_CUDA(cudaFuncSetAttribute(tc_recon_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
inline void launch_wterms(const TcState& s, int grid, cudaStream_t stream, const CUtensorMap& x, const CUtensorMap& e,
                          const CUtensorMap& xlo, const CUtensorMap& elo, const WTermsParams& p) {
  if (s.x3) tc_wterms_x3_kernel<<<grid, kSThreads, wterms8_smem_bytes(s.f.s), stream>>>(x, e, s.tmH_k2, xlo, elo, p);
  else tc_wterms_kernel<<<grid, kWtThreads, wterms_smem_bytes(s.f.s), stream>>>(x, e, s.tmH_k2, p);
}

// ---- lag autocorrelation of H:  P[d][a][b] = sum_t H[a][t] H[b][t-d]  (the W-terms kernel run on H^T itself) ----
// Used by the Gram route (den_w_gram) and by lipschitz_W of the gradient solvers (gradient_descent.py:54-57).
inline int ensure_autocorr(TcState& s) {
  if (s.autocorr_ready) return 0;
  const Dims& d = s.d;
  const Fold& f = s.f;
  const long long pcount = (long long)d.L * d.Kp * d.Kp;
  const int lpr = s.x3 ? 8 : 16;                    // virtual lags per accumulator quarter / per plain item
  // quad mode (WTermsParams::quad): H^T has at most 32 "features", so the four row quarters of the accumulator
  // take four consecutive lag blocks instead of 96 idle rows
  // (3xTF32 only: plain TF32 keeps the time chunks of the numerator pass - equal chains, equal truncation bias -
  // and with the chunk count fixed a wider item buys nothing)
  s.p_quad = (d.Kp <= 32 && s.x3) ? 1 : 0;
  if (const char* e = getenv("CMF_P_QUAD")) s.p_quad = atoi(e) && d.Kp <= 32;
  s.p_lag_groups = (int)ceil_div_ll(f.Lv, s.p_quad ? 4 * lpr : lpr);
  s.p_stage0 = s.p_quad ? -ceil_div_ll(3ll * lpr * f.s, 32) : 0;
  {   // same time chunks as the numerator pass: equal accumulation chains, equal truncation bias
    const long long units = (long long)s.p_lag_groups * f.CB;
    const long long stages_total = ceil_div_ll(d.Tloc, 32) - s.p_stage0;
    long long c = s.n_chunks;
    if (s.p_quad && c * units < d.num_sms) c = ceil_div_ll(d.num_sms, units);     // fewer items: keep the SMs busy
    if (c > stages_total) c = stages_total;
    if (c < 1) c = 1;
    s.p_chunks = (int)c;
    const long long items = units * c;
    s.p_grid = (int)(items < d.num_sms ? items : d.num_sms);
  }
  CMF_TRY(cached_malloc((void**)&s.P, (size_t)pcount * 4));
  CMF_TRY(cached_malloc((void**)&s.Ppart, (size_t)pcount * s.p_chunks * 4));
  // H^T itself as the "data" operand: the first Kp columns of Hv are the unfolded, rounded H^T
  const float* base = s.Hv + (long long)d.h * s.KWs;
  if (s.p_quad) {
    CMF_TRY(make_map(&s.tmHx_k2, base, d.Tloc, d.Kp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, s.KWs));
    s.tmHxlo_k2 = s.tmHx_k2;
    if (s.x3) CMF_TRY(make_map(&s.tmHxlo_k2, base + f.KW, d.Tloc, d.Kp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, s.KWs));
  } else {
    CMF_TRY(make_map_k2src(&s.tmHx_k2, base, d.Tloc, d.Kp, s.KWs));
    s.tmHxlo_k2 = s.tmHx_k2;
    if (s.x3) CMF_TRY(make_map_k2src(&s.tmHxlo_k2, base + f.KW, d.Tloc, d.Kp, s.KWs));
  }
  s.autocorr_ready = true;
  return 0;
}

inline int autocorr(TcState& s, cudaStream_t stream) {
  CMF_TRY(ensure_autocorr(s));
  const Dims& d = s.d;
  const Fold& f = s.f;
  const long long pcount = (long long)d.L * d.Kp * d.Kp;
  WTermsParams p{};
  p.Np = d.Kp; p.L = d.L; p.n_tiles_n = 1; p.n_lag_groups = s.p_lag_groups; p.n_chunks = s.p_chunks; p.h = d.h;
  p.Lv = f.Lv; p.Kp = d.Kp; p.s = f.s; p.CB = f.CB; p.brows = s.x3 ? wterms8_brows(f.s) : wterms_brows(f.s); p.n_src = 1;
  p.n_items = (long long)p.n_lag_groups * f.CB * p.n_chunks;
  p.stages_total = ceil_div_ll(d.Tloc, 32);
  p.quad = s.p_quad; p.stage0 = s.p_stage0;
  p.part = (s.p_chunks == 1) ? s.P : s.Ppart;
  p.per_src = pcount; p.err = s.d_err;
  p.x3 = s.x3; p.lo_off = f.KW; p.sub_units = s.x3 ? strict_sub_units() : 0;
  launch_wterms(s, s.p_grid, stream, s.tmHx_k2, s.tmHx_k2, s.tmHxlo_k2, s.tmHxlo_k2, p);
  CMF_TRY(launch_ok("autocorr_H"));
  if (s.p_chunks > 1) {
    ew::sum_splits_kernel<<<ew_blocks(s, pcount / 4), 256, 0, stream>>>((float4*)s.P, (const float4*)s.Ppart, pcount / 4,
                                                                       pcount / 4, s.p_chunks);
    CMF_TRY(launch_ok("autocorr_H_sum"));
  }
  return 0;
}

inline int init(TcState& s, const Dims& d, float* Xt, float* Et, float* Ht, float* W, float* numden, float* hterms,
                double* loss_partials, long long n_loss_partials, double* d_sumsq, cudaStream_t stream,
                float* Xlo = nullptr) {
  s.d = d;
  s.f = make_fold(d.Kp, d.L);
  const Fold& f = s.f;
  s.x3 = Xlo != nullptr ? 1 : 0;
  s.Xlo = Xlo;
  s.KWs = (s.x3 ? 2 : 1) * f.KW;
  s.Xt = Xt; s.Et = Et; s.Ht = Ht; s.W = W; s.numden = numden; s.hterms = hterms;
  s.loss_partials = loss_partials; s.d_sumsq = d_sumsq;
  s.wcount = (long long)d.L * d.Np * d.Kp;
  s.wv_count = (long long)f.Lv * d.Np * f.KW;      // per half
  s.hv_count = d.RH * f.KW;
  const int halves = s.x3 ? 2 : 1;
  if (const char* e = getenv("CMF_TC_MASK")) s.mask = atoi(e);
  if (s.x3) s.mask = 7;
  CMF_CHECK(d.Kp == padded_k(d.K) && shape_supported(d.N, d.K, d.L), "shape not supported by the tensor-core path");
  CMF_CHECK(n_loss_partials >= d.num_sms, "loss partial buffer too small");

  CMF_TRY(cached_malloc((void**)&s.d_err, 4));
  CMF_CUDA(cudaMemsetAsync(s.d_err, 0, 4, stream));
  // (K3 addresses Wv as n_glag x J lags: the lags past Lv exist and stay zero)
  const long long lv_alloc = (long long)f.n_glag * f.J > f.Lv ? (long long)f.n_glag * f.J : f.Lv;
  const size_t wv_bytes = (size_t)lv_alloc * d.Np * f.KW * halves * 4;
  CMF_TRY(cached_malloc((void**)&s.Wv, wv_bytes));
  CMF_TRY(cached_malloc((void**)&s.Hv, (size_t)s.hv_count * halves * 4));
  CMF_CUDA(cudaMemsetAsync(s.Wv, 0, wv_bytes, stream));
  CMF_CUDA(cudaMemsetAsync(s.Hv, 0, (size_t)s.hv_count * halves * 4, stream));

  // ---- K1 -------------------------------------------------------------
  {
    const long long tiles = ceil_div_ll(d.Np, 128) * (d.RT / 256);
    s.recon_grid = (int)(tiles < d.num_sms ? tiles : d.num_sms);
  }
  CMF_TRY(make_map(&s.tmW_k1, s.Wv, (long long)f.Lv * d.Np, s.KWs, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_TRY(make_map(&s.tmH_k1, s.Hv, d.RH, s.KWs, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_TRY(set_recon_smem(recon_smem_bytes(f.recon_wrows)));

  // ---- K2 -------------------------------------------------------------
  s.n_lag_groups = (int)ceil_div_ll(f.Lv, s.x3 ? 8 : 16);
  {
    const long long units = ceil_div_ll(d.Np, 128) * s.n_lag_groups * f.CB * 2;
    const long long stages_total = ceil_div_ll(d.Tloc, 32);
    long long cmax = stages_total / 4; if (cmax < 1) cmax = 1;
    const long long cmem = (2ll << 30) / (2 * s.wcount * 4); if (cmem < cmax) cmax = cmem < 1 ? 1 : cmem;
    if (cmax > 256) cmax = 256;
    // Number of time chunks, from a cost model of the pass (microseconds): whole waves of equally long items
    // (a stage = 8 MMAs of 128 cycles, three passes with 3xTF32), a tensor-memory drain per item, and the
    // write + re-read of one partial [src][L][Np][Kp] per chunk by sum_splits_kernel.  Long shards end up with
    // the chunk count that fills the machine in whole waves (37 at config C on one GPU); short shards (T split
    // over 8 GPUs) take fewer, larger chunks because the partial traffic no longer amortises.  Among counts
    // within 1 % of the best the largest wins: shorter accumulation chains in tensor memory.
    const double t_stage = s.x3 ? 0.3 * 3.0 : 0.6, t_drain = s.x3 ? 1.0 : 4.0;   // 3xTF32: 4 MMAs per stage, overlapped folds
    const double t_partial = 2.0 * 2.0 * (double)s.wcount * 4.0 / 5.0e6;     // both sources; 5 TB/s
    double best = 1e300;
    std::vector<double> cost((size_t)cmax + 1, 1e300);
    // On the Gram route the autocorrelation pass of H (den_w_gram) uses the same chunks - equal accumulation
    // chains - but has only n_lag_groups * CB units: too few chunks would leave most SMs idle there.
    const long long p_units = (s.gram_request & 2) ? (long long)s.n_lag_groups * f.CB : 0;
    for (long long c = 1; c <= cmax; ++c) {
      const double item = (double)ceil_div_ll(stages_total, c) * t_stage + t_drain;
      cost[c] = (double)ceil_div_ll(units * c, d.num_sms) * item + (c > 1 ? c * t_partial : 0.0);
      if (p_units) cost[c] += (double)ceil_div_ll(p_units * c, d.num_sms) * item;
      if (cost[c] < best) best = cost[c];
    }
    int bestc = 1;
    for (long long c = 1; c <= cmax; ++c)
      if (cost[c] <= 1.01 * best) bestc = (int)c;
    s.n_chunks = bestc;
    if (const char* e = getenv("CMF_WCHUNKS")) { long long c = atoll(e); s.n_chunks = (int)(c < 1 ? 1 : (c > cmax ? cmax : c)); }
    const long long items = units * s.n_chunks;
    s.wterms_grid = (int)(items < d.num_sms ? items : d.num_sms);
  }
  CMF_TRY(make_map_k2src(&s.tmX_k2, Xt, d.Tloc, d.Np, d.Np));
  s.tmXlo_k2 = s.tmX_k2;
  if (s.x3) CMF_TRY(make_map_k2src(&s.tmXlo_k2, Xlo, d.Tloc, d.Np, d.Np));
  CMF_TRY(make_map(&s.tmH_k2, s.Hv, d.RH, s.KWs, 32, s.x3 ? wterms8_brows(f.s) : wterms_brows(f.s),
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  CMF_CUDA(cudaFuncSetAttribute(tc_wterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wterms_smem_bytes(f.s)));
  CMF_CUDA(cudaFuncSetAttribute(tc_wterms_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wterms8_smem_bytes(f.s)));

  // ---- K3 -------------------------------------------------------------
  {
    const long long tiles = d.TO / 256 + 1;
    s.hterms_grid = (int)(tiles < d.num_sms ? tiles : d.num_sms);
  }
  CMF_TRY(make_map_k3w(&s.tmW_k3, s.Wv, f, d.Np, s.KWs));
  CMF_TRY(make_map(&s.tmX_k3, Xt, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  s.tmXlo_k3 = s.tmX_k3;
  if (s.x3) CMF_TRY(make_map(&s.tmXlo_k3, Xlo, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  CMF_CUDA(cudaFuncSetAttribute(tc_hterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)hterms_smem_bytes(f.h_stages, f.hterms_wrows, d.Kp, f.h_hd, f.n_glag == 1 && f.s == 1, f.h_staged != 0)));

  // ---- Gram route ---------------------------------------------------------
  s.gram = s.gram_request;
  if (const char* e = getenv("CMF_GRAM")) s.gram = atoi(e);
  s.LK = d.L * d.Kp;
  s.Lr = 2 * d.L - 1;
  s.Lrv = (s.Lr + f.s - 1) / f.s;
  s.g_rows = round_up_ll(s.LK, 256);
  s.ntail = d.Tloc + d.h - d.t_valid;
  {   // the fold of den_H = R (*) H on the H-terms kernel: 2L-1 lags, Kp "features"
    s.fr = f;
    s.fr.Lv = s.Lrv;
    s.fr.J = (s.Lrv + f.n_glag - 1) / f.n_glag;
    s.fr.hterms_wrows = round_up(256 + f.s * (s.fr.J - 1), 32);
    s.fr.h_hd = (f.n_glag - 1) * f.s * s.fr.J + f.s - 1;
    choose_hterms_ring(s.fr, f.n_glag == 1 && f.s == 1);
    if (!s.fr.h_stages) s.gram &= ~1;
  }
  if ((long long)s.g_rows * s.LK * 4 > (1ll << 30)) s.gram &= ~1;
  if (s.gram & 1) {
    // 3xTF32: Wt rows are [hi (NpA) | lo (NpA)], Rwv rows [hi (KW) | lo (KW)]
    s.NpA = s.x3 ? round_up(d.Np, 32) : d.Np;
    const long long wt_ld = (long long)halves * s.NpA;
    CMF_TRY(cached_malloc((void**)&s.Wt, (size_t)s.LK * wt_ld * 4));
    CMF_CUDA(cudaMemsetAsync(s.Wt, 0, (size_t)s.LK * wt_ld * 4, stream));
    CMF_TRY(cached_malloc((void**)&s.G, (size_t)s.g_rows * s.LK * 4));
    CMF_TRY(cached_malloc((void**)&s.Rw, (size_t)s.Lr * d.Kp * d.Kp * 4));
    // Rwv: R as a W-like operand [lag][k' (feature)][k], folded like Wv; allocated for n_glag * J lags (zeros past Lrv)
    const long long lr_alloc = (long long)f.n_glag * s.fr.J > s.Lrv ? (long long)f.n_glag * s.fr.J : s.Lrv;
    CMF_TRY(cached_malloc((void**)&s.Rwv, (size_t)lr_alloc * d.Kp * s.KWs * 4));
    CMF_CUDA(cudaMemsetAsync(s.Rwv, 0, (size_t)lr_alloc * d.Kp * s.KWs * 4, stream));
    CMF_TRY(make_map(&s.tmWt_a, s.Wt, s.LK, wt_ld, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B));
    CMF_TRY(make_map(&s.tmWt_b, s.Wt, s.LK, wt_ld, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    CMF_TRY(make_map_k3w(&s.tmRw_k3, s.Rwv, s.fr, d.Kp, s.KWs));
    // H^T itself as the data operand: the first Kp columns of Hv (row 0 of Hv is time -(L-1): lag l of R is d = l - (L-1))
    CMF_TRY(make_map(&s.tmHs_k3, s.Hv, d.RH, d.Kp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B, s.KWs));
    s.tmHslo_k3 = s.tmHs_k3;
    if (s.x3) CMF_TRY(make_map(&s.tmHslo_k3, s.Hv + f.KW, d.RH, d.Kp, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B, s.KWs));
    if (s.fr.h_hd > 0) CMF_TRY(cached_malloc((void**)&s.hcarry_r, (size_t)(d.TO / 256 + 1) * s.fr.h_hd * d.Kp * 4));
    // one attribute for both uses of the H-terms kernel
    const bool direct = f.n_glag == 1 && f.s == 1;
    const size_t a = hterms_smem_bytes(f.h_stages, f.hterms_wrows, d.Kp, f.h_hd, direct, f.h_staged != 0);
    const size_t b = hterms_smem_bytes(s.fr.h_stages, s.fr.hterms_wrows, d.Kp, s.fr.h_hd, direct, s.fr.h_staged != 0);
    CMF_CUDA(cudaFuncSetAttribute(tc_hterms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(a > b ? a : b)));
  }
  // H terms.  3xTF32: short tensor-memory sub-chunks folded in fp32 registers (no truncation bias to speak of).
  // Plain TF32 with the Gram denominator: sub-chunks as long as the chain of R (*) H, so that numerator and
  // denominator carry the same truncation bias and the W / H scale split does not drift; direct route: numerator
  // and denominator share one chain per item anyway.
  {
    // (the Gram denominator is one chain of ceil(Kp / 32) * J_r units per time tile on the same kernel)
    s.h_sub = s.x3 ? strict_sub_units() : ((s.gram & 1) ? (int)ceil_div_ll(d.Kp, 32) * s.fr.J : 0);
    if (const char* e = getenv("CMF_HSUB")) s.h_sub = atoi(e);
    const long long tt = d.TO / 256 + 1;
    const int n_src = (s.gram & 1) ? 1 : 2;
    long long items = tt * n_src;
    // Work items are whole time tiles.  On a short shard (config C on 8 GPUs: 513 tiles = 3.47 rounds of 148 CTAs, so
    // 4) the last round runs half empty; when halving the items (two CTAs share a tile's feature chunks and write
    // partial outputs, summed afterwards) fills the rounds better by 4 % or more, do that.
    s.h_split = 1;
    {
      const long long cn = ceil_div_ll(d.Np, 32);
      const auto eff = [&](long long n) { return (double)n / (double)(ceil_div_ll(n, d.num_sms) * d.num_sms); };
      // (only where an item is long enough to pay for the partial-sum pass and the extra carry launches)
      // 3xTF32 only: its two-level accumulation makes the truncation bias of a result independent of the chain length;
      // in plain TF32 the half chains of a split numerator would no longer carry the bias of the Gram denominator's
      // chain and the W / H scale split would drift (measured: 1e-3 of sum(W) after 100 iterations of `mid`)
      if (s.x3 && cn % 2 == 0 && cn * f.J >= 128 && items > d.num_sms && eff(2 * items) > eff(items) + 0.04) s.h_split = 2;
      if (const char* e = getenv("CMF_HT_SPLIT")) {
        const int v = atoi(e);
        s.h_split = (v >= 1 && v <= 4 && cn % v == 0) ? v : 1;
      }
    }
    if (f.h_hd > 0) CMF_TRY(cached_malloc((void**)&s.hcarry, (size_t)2 * s.h_split * tt * f.h_hd * d.Kp * 4));
    if (s.h_split > 1) CMF_TRY(cached_malloc((void**)&s.hparts, (size_t)2 * s.h_split * d.TO * d.Kp * 4));
    items *= s.h_split;
    s.hterms_grid = (int)(items < d.num_sms ? items : d.num_sms);
  }
  if ((long long)s.g_rows * f.Lv * f.KW * halves * 4 > (1ll << 30)) s.gram &= ~2;
  CMF_TRY(cached_malloc((void**)&s.wpart, (size_t)s.n_chunks * ((s.gram & 2) ? 1 : 2) * s.wcount * 4));
  if (s.gram && s.ntail > 0) CMF_TRY(cached_malloc((void**)&s.Etail, (size_t)round_up_ll(s.ntail, 256) * d.Np * 4));
  if (s.gram & 2) {
    CMF_TRY(ensure_autocorr(s));
    static const bool shift_form = [] { const char* e = getenv("CMF_DENW_SHIFT"); return !e || atoi(e) != 0; }();
    if (shift_form) {
      // den_W as a shift-GEMM (autocorr_shift_operand_kernel): lag stride s_rows rows, lags walked in blocks of pw_LB
      s.pw_srows = d.Kp > 32 ? d.Kp : 32;
      const int max_rows = (int)((kMaxSmem - recon_smem_bytes(0)) / (2 * kKp * 4) / 64) * 64;
      int LB = (max_rows - 256) / s.pw_srows + 1;    // lags whose shifted windows fit one shared-memory window
      const bool fits = LB >= 2 || LB >= f.Lv;
      if (LB >= f.Lv) LB = 0;                        // every lag in one window
      else LB &= ~1;                                 // (two lags per pipeline stage)
      s.pw_LB = LB;
      s.pw_wrows = round_up(256 + s.pw_srows * ((LB > 0 ? LB : f.Lv) - 1), 64);
      s.pw_rows = round_up_ll(s.g_rows + (long long)s.pw_srows * f.Lv + s.pw_wrows, 64);
      if (!fits || recon_smem_bytes(s.pw_wrows) > kMaxSmem) {
        s.pw_rows = 0;                               // (no room for two lags per window: keep the Toeplitz GEMM)
      } else {
        CMF_TRY(cached_malloc((void**)&s.Pw, (size_t)s.pw_rows * s.KWs * 4));
        CMF_TRY(make_map(&s.tmPw_b, s.Pw, s.pw_rows, s.KWs, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
        const size_t a = recon_smem_bytes(f.recon_wrows), b = recon_smem_bytes(s.pw_wrows);
        CMF_TRY(set_recon_smem(a > b ? a : b));
      }
    }
    if (s.pw_rows == 0) {
      CMF_TRY(cached_malloc((void**)&s.Mt, (size_t)s.g_rows * f.Lv * f.KW * halves * 4));
      CMF_TRY(make_map(&s.tmMt_b, s.Mt, s.g_rows, (long long)f.Lv * f.KW * halves, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    }
  }
  s.ready = true;
  return 0;
}

// The est buffer is only needed when some MU step contracts est (direct routes) or est is read back;
// with both denominators on the Gram route it is allocated on first demand (cmf_abi.cu).
inline int attach_est(TcState& s, float* Et, float* Elo = nullptr) {
  const Dims& d = s.d;
  s.Et = Et;
  s.Elo = Elo;
  s.tmE_k2 = s.tmX_k2;
  s.tmE_k3 = s.tmX_k3;
  s.tmElo_k2 = s.tmX_k2;
  s.tmElo_k3 = s.tmX_k3;
  if (!Et) return 0;
  CMF_TRY(make_map_k2src(&s.tmE_k2, Et, d.Tloc, d.Np, d.Np));
  CMF_TRY(make_map(&s.tmE_k3, Et, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  if (s.x3) {
    CMF_CHECK(Elo != nullptr, "3xTF32 needs the lo half of est");
    CMF_TRY(make_map_k2src(&s.tmElo_k2, Elo, d.Tloc, d.Np, d.Np));
    CMF_TRY(make_map(&s.tmElo_k3, Elo, d.RT, d.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  return 0;
}

// untruncated est rows t_valid .. t_valid + ntail into Etail (whole time tiles of the recon kernel, no tail mask)
inline int tail_est(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  ReconParams p{};
  p.Np = d.Np; p.L = f.Lv; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.wrows = f.recon_wrows;
  p.s = f.s; p.CB = f.CB; p.cb_cols = f.CB; p.h_shift = (int)(d.h - f.s * (f.Lv - 1) + d.t_valid);
  p.n_rows = d.Np; p.ld_out = d.Np; p.store_mode = 0; p.w_kp = d.Kp; p.w_np = d.Np;
  const long long tail_tiles = ceil_div_ll(s.ntail, 256);          // L - 1 rows of est lie past the end of the data
  p.n_tiles = p.n_tiles_n * tail_tiles;
  p.t_own = 0; p.t_valid = 256 * tail_tiles;
  p.Et = s.Etail; p.Xt = nullptr; p.loss_partials = s.loss_partials + d.num_sms; p.round_out = 0; p.err = s.d_err;
  p.LB = f.recon_LB;
  set_x3(s, p, f.KW, f.KW);
  launch_recon(s, (int)(p.n_tiles < d.num_sms ? p.n_tiles : d.num_sms), recon_smem_bytes(f.recon_wrows), stream, s.tmW_k1,
               s.tmH_k1, p);
  return launch_ok("tail_est");
}

// ---- Gram route, W step: den_W = W (*) A - tail, written into the den half of numden ----
inline int den_w_gram(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  // (a) P[d][k'][k] = sum_t H[k'][t] H[k][t-d] over the owned columns: the W-terms kernel on H^T
  CMF_TRY(autocorr(s, stream));
  float* den = s.numden + s.wcount;
  if (s.pw_rows > 0) {
    // (b) the autocorrelation as a shift operand, (c) den_W = W (*) A as a shift-GEMM on the recon kernel:
    //     "time" = (l,k), lag stride pw_srows rows, output in the W layout (store mode 2)
    autocorr_shift_operand_kernel<<<ew_blocks(s, s.pw_rows * f.KW), 256, 0, stream>>>(s.P, s.Pw, d.L, f.Lv, d.Kp, f.s, f.KW,
                                                                                   s.pw_rows, s.x3);
    CMF_TRY(launch_ok("autocorr_operand"));
    ReconParams p{};
    p.Np = d.Np; p.L = f.Lv; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.wrows = s.pw_wrows;
    p.s = s.pw_srows; p.CB = f.CB; p.cb_cols = f.CB; p.h_shift = 0;
    p.n_rows = d.Np; p.ld_out = d.Np; p.store_mode = 2; p.w_kp = d.Kp; p.w_np = d.Np;
    p.n_tiles = (long long)p.n_tiles_n * (s.g_rows / 256);
    p.t_own = 0; p.t_valid = s.LK;
    p.Et = den; p.Xt = nullptr; p.loss_partials = s.loss_partials + d.num_sms; p.round_out = 0; p.err = s.d_err;
    p.LB = s.pw_LB;
    const int grid = (int)(p.n_tiles < d.num_sms ? p.n_tiles : d.num_sms);
    set_x3(s, p, f.KW, f.KW);
    launch_recon(s, grid, recon_smem_bytes(s.pw_wrows), stream, s.tmW_k1, s.tmPw_b, p);
    CMF_TRY(launch_ok("gram_den_w"));
  } else {
    // (b) block-Toeplitz operand
    toeplitz_kernel<<<ew_blocks(s, s.g_rows * f.Lv * f.KW), 256, 0, stream>>>(s.P, s.Mt, d.L, f.Lv, d.Kp, f.s, f.KW, s.g_rows, s.x3);
    CMF_TRY(launch_ok("toeplitz"));
    // (c) den_W[n][(l,k)] = sum_{(l'v,c)} Wv[l'v][n][c] Mt[(l,k)][(l'v,c)]  (plain GEMM on the recon kernel)
    ReconParams p{};
    p.Np = d.Np; p.L = 1; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.wrows = 256;
    p.s = 0; p.CB = f.Lv * f.CB; p.cb_cols = f.CB; p.h_shift = 0;
    p.n_rows = d.Np; p.ld_out = d.Np; p.store_mode = 2; p.w_kp = d.Kp; p.w_np = d.Np;
    p.n_tiles = (long long)p.n_tiles_n * (s.g_rows / 256);
    p.t_own = 0; p.t_valid = s.LK;
    p.Et = den; p.Xt = nullptr; p.loss_partials = s.loss_partials + d.num_sms; p.round_out = 0; p.err = s.d_err;
    const int grid = (int)(p.n_tiles < d.num_sms ? p.n_tiles : d.num_sms);
    set_x3(s, p, f.KW, f.Lv * f.KW);
    launch_recon(s, grid, recon_smem_bytes(256), stream, s.tmW_k1, s.tmMt_b, p);
    CMF_TRY(launch_ok("gram_den_w"));
  }
  // (d) remove the terms of est that lie past the end of the data
  if (s.ntail > 0) {
    CMF_TRY(tail_est(s, stream));
    simt::WTermsA a{s.Etail, s.Etail, d.Np};
    simt::WTermsB b{s.Ht, d.Kp, (int)(d.h + d.t_valid), d.L * d.Kp};
    simt::SubWEpi e{den, d.Np, d.Kp, d.L * d.Kp};
    dim3 grid((unsigned)ceil_div_ll(d.Np, 128), (unsigned)ceil_div_ll((long long)d.L * d.Kp, 128), 1);
    simt::shift_gemm_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, stream>>>(a, b, e, s.ntail, round_up_ll(s.ntail, 16), 1);
    CMF_TRY(launch_ok("tail_den_w"));
  }
  return 0;
}

// need_loss = false: the reconstruction in the middle of an iteration (between the W and the H step) - nobody reads
// its residual, so the epilogue does not load X (hi and lo halves: half of the launch's HBM traffic on problems whose
// K1 is bound by its epilogue, config B) and the partial sums are not added up.
inline int recon(TcState& s, cudaStream_t stream, bool store_est = true, bool need_loss = true) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  ReconParams p{};
  p.skip_store = store_est ? 0 : 1;
  p.Np = d.Np; p.L = f.Lv; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.wrows = f.recon_wrows;
  p.s = f.s; p.CB = f.CB; p.h_shift = d.h - f.s * (f.Lv - 1);
  p.cb_cols = f.CB; p.n_rows = d.Np; p.ld_out = d.Np; p.store_mode = 0; p.w_kp = d.Kp; p.w_np = d.Np;
  p.n_tiles = (long long)p.n_tiles_n * (d.RT / 256);
  p.t_own = d.Tloc; p.t_valid = d.t_valid;
  p.Et = s.Et; p.Xt = s.Xt; p.loss_partials = s.loss_partials; p.round_out = 1; p.err = s.d_err;
  p.LB = f.recon_LB;
  const bool loss_only_fast = s.x3 && !store_est && s.loss_fast;
  if (loss_only_fast) {
    // The reconstruction that only feeds the loss (both denominators on the Gram route) runs the hi x hi operand
    // pass alone, with the same two-level accumulation.  What it drops, W_lo (*) H_hi + W_hi (*) H_lo, are the
    // rounding residuals of the factors (relative size 2^-12, random signs); the ABI switches this on only for
    // problems where their effect on ||est - X||^2 is below 1e-6 relative (decide_loss_mode in cmf_abi.cu:
    // a million or more factor entries, K L loss^2 >= 0.2); the factors W and H never see it.
    p.sub_units = 4 * strict_sub_units();   // (a uniform scale bias of est moves the loss far less than it would move W or H)
    p.Xlo = s.Xlo;
  } else if (s.x3) {
    set_x3(s, p, f.KW, 0);
    p.Elo = s.Elo; p.Xlo = s.Xlo;
  }
  if (!need_loss) { p.t_own = 0; p.Xlo = nullptr; }
  const int grid = s.recon_grid;
  launch_recon(s, grid, recon_smem_bytes(f.recon_wrows), stream, s.tmW_k1, s.tmH_k1, p);
  CMF_TRY(launch_ok(loss_only_fast ? "tc_recon_loss_1pass" : "tc_recon"));
  if (!need_loss) return 0;
  ew::sum_doubles_kernel<<<1, 1024, 0, stream>>>(s.loss_partials, grid, s.d_sumsq);
  return launch_ok("loss_sum");
}

inline int w_terms(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  WTermsParams p{};
  p.Np = d.Np; p.L = d.L; p.n_tiles_n = (int)ceil_div_ll(d.Np, 128); p.n_lag_groups = s.n_lag_groups;
  p.n_chunks = s.n_chunks; p.h = d.h;
  p.Lv = f.Lv; p.Kp = d.Kp; p.s = f.s; p.CB = f.CB; p.brows = s.x3 ? wterms8_brows(f.s) : wterms_brows(f.s);
  p.n_src = (s.gram & 2) ? 1 : 2;
  p.n_items = (long long)p.n_tiles_n * p.n_lag_groups * f.CB * p.n_src * p.n_chunks;
  p.stages_total = ceil_div_ll(d.Tloc, 32);
  p.part = (s.n_chunks == 1) ? s.numden : s.wpart;
  p.per_src = s.wcount; p.err = s.d_err;
  p.x3 = s.x3; p.lo_off = f.KW; p.sub_units = s.x3 ? strict_sub_units() : 0;
  launch_wterms(s, s.wterms_grid, stream, s.tmX_k2, s.tmE_k2, s.tmXlo_k2, s.tmElo_k2, p);
  CMF_TRY(launch_ok("tc_wterms"));
  if (s.n_chunks > 1) {
    const long long n4 = p.n_src * s.wcount / 4;
    ew::sum_splits_kernel<<<ew_blocks(s, n4), 256, 0, stream>>>((float4*)s.numden, (const float4*)s.wpart, n4, n4, s.n_chunks);
    CMF_TRY(launch_ok("w_terms_sum"));
  }
  if (s.gram & 2) CMF_TRY(den_w_gram(s, stream));
  return 0;
}

// ---- Gram route, H step: den_H = R (*) H - tail -------------------------------
inline int den_h_gram(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  // (a) Wt = round(W)^T
  {
    dim3 grid((unsigned)ceil_div_ll(d.Np, 32), (unsigned)ceil_div_ll(d.Kp, 32), (unsigned)d.L);
    transpose_round_w_kernel<<<grid, 256, 0, stream>>>(s.W, s.Wt, d.Np, d.Kp, (long long)(s.x3 ? 2 : 1) * s.NpA,
                                                       s.x3 ? s.NpA : 0);
    CMF_TRY(launch_ok("transpose_round_w"));
  }
  // (b) G = Wt Wt^T : a plain GEMM on the recon kernel (one "lag", reduction blocks = 32-feature chunks)
  {
    ReconParams p{};
    p.Np = 0; p.L = 1; p.n_tiles_n = (int)ceil_div_ll(s.LK, 128); p.wrows = 256;
    p.s = 0; p.CB = (int)ceil_div_ll(d.Np, 32); p.h_shift = 0; p.cb_cols = p.CB;
    p.n_rows = s.LK; p.ld_out = s.LK; p.store_mode = 0; p.w_kp = d.Kp; p.w_np = d.Np;
    p.n_tiles = (long long)p.n_tiles_n * (s.g_rows / 256);
    p.t_own = 0; p.t_valid = s.g_rows;
    p.Et = s.G; p.Xt = nullptr; p.loss_partials = s.loss_partials + d.num_sms; p.round_out = 0; p.err = s.d_err;
    const int grid = (int)(p.n_tiles < d.num_sms ? p.n_tiles : d.num_sms);
    set_x3(s, p, s.NpA, s.NpA);
    launch_recon(s, grid, recon_smem_bytes(256), stream, s.tmWt_a, s.tmWt_b, p);
    CMF_TRY(launch_ok("gram_G"));
  }
  // (c) R = lag-diagonal sums of G as a W-like operand Rt[l][k'][k] = R[l - (L-1)][k][k'] (rounded / folded like W)
  diag_sum_kernel<<<ew_blocks(s, (long long)s.Lr * d.Kp * d.Kp), 256, 0, stream>>>(s.G, s.LK, s.Rw, d.L, d.Kp);
  CMF_TRY(launch_ok("diag_sum"));
  if (s.x3) {
    fold_w_x3_kernel<<<ew_blocks(s, (long long)s.Lrv * d.Kp * f.KW), 256, 0, stream>>>(s.Rwv, s.Rw, s.Lr, s.Lrv, d.Kp, d.Kp, f.s, f.KW);
  } else if (f.s == 1) {
    const long long n4 = (long long)s.Lr * d.Kp * d.Kp / 4;
    ew::round_copy_kernel<<<ew_blocks(s, n4), 256, 0, stream>>>((float4*)s.Rwv, (const float4*)s.Rw, n4);
  } else {
    fold_w_kernel<<<ew_blocks(s, (long long)s.Lrv * d.Kp * 32), 256, 0, stream>>>(s.Rwv, s.Rw, s.Lr, s.Lrv, d.Kp, d.Kp, f.s);
  }
  CMF_TRY(launch_ok("fold_R"));
  // (d) den_H^T[t][k] = sum_l sum_k' Rt[l][k'][k] H^T[t + l - (L-1)][k']: tensor_transconv of R with H on the
  //     H-terms kernel (features := components; the four lag groups fill the 128 MMA rows, where the reconstruction
  //     kernel used for this until round 2 left 96 of them idle)
  float* den = s.hterms + d.TO * d.Kp;
  {
    const Fold& r = s.fr;
    HTermsParams p{};
    p.Np = d.Kp; p.J = r.J; p.n_chunks_n = (int)ceil_div_ll(d.Kp, 32); p.wrows = r.hterms_wrows;
    p.s = f.s; p.CB = f.CB; p.Kp = d.Kp;
    p.n_src = 1;
    p.n_time_tiles = d.TO / 256 + 1;
    p.n_items = p.n_time_tiles;
    p.TO = d.TO; p.out = den; p.carry = s.hcarry_r; p.hd = r.h_hd;
    p.sub_units = s.x3 ? strict_sub_units() : 0; p.n_stages = r.h_stages; p.staged = r.h_staged;
    p.x3 = s.x3; p.lo_off = f.KW; p.err = s.d_err;
    const bool direct = f.n_glag == 1 && f.s == 1;
    const int grid = (int)(p.n_items < d.num_sms ? p.n_items : d.num_sms);
    tc_hterms_kernel<<<grid, kSThreads, hterms_smem_bytes(r.h_stages, r.hterms_wrows, d.Kp, r.h_hd, direct, r.h_staged != 0), stream>>>(
        s.tmRw_k3, s.tmHs_k3, s.tmHs_k3, s.tmHslo_k3, s.tmHslo_k3, p);
    CMF_TRY(launch_ok("gram_den_h"));
    if (r.h_hd > 0) {
      const long long wmax = r.h_hd < 256 ? r.h_hd : 256;
      const long long total = (p.n_time_tiles - 1) * wmax * (d.Kp / 4);
      if (total > 0) {
        hterms_carry_kernel<<<ew_blocks(s, total), 256, 0, stream>>>(den, s.hcarry_r, d.TO, p.n_time_tiles, r.h_hd, d.Kp, 1);
        CMF_TRY(launch_ok("gram_den_h_carry"));
      }
    }
  }
  // (e) remove the terms of est that lie past the end of the data (only the shard that sees the end)
  if (s.ntail > 0) {
    const long long t0 = d.t_valid;
    CMF_TRY(tail_est(s, stream));
    {   // den_H[tau] -= sum_{l : tau + l >= t0} W[l]^T est_ext[tau + l],  tau in [t0 - h, Tloc)
      const long long m_off = t0 - d.h;
      const long long rows = d.Tloc - m_off;
      if (rows > 0) {
        // one split per lag (the reduction (l, n) is long and the output tiny); partials live in G's buffer
        simt::TailHA a{s.Etail, d.Np, m_off, t0, s.ntail};
        simt::HTermsB b{s.W, d.Kp};
        // the reduction (l, n) is long and the output tiny: split it into about two blocks per SM (this runs on
        // ONE rank while the others wait at the next exchange); partials live in G's buffer
        const long long per_split = round_up_ll(rows, 128) * d.Kp;
        const long long R = (long long)d.L * d.Np;
        const long long m_tiles = ceil_div_ll(rows, 128);
        long long r_chunk = round_up_ll(ceil_div_ll(R * m_tiles, 2ll * d.num_sms), 16);
        if (r_chunk < 64) r_chunk = 64;
        long long nsplit = ceil_div_ll(R, r_chunk);
        if (per_split * nsplit > s.g_rows * s.LK) { nsplit = (s.g_rows * s.LK) / per_split; r_chunk = round_up_ll(ceil_div_ll(R, nsplit), 16); nsplit = ceil_div_ll(R, r_chunk); }
        CMF_CHECK(nsplit >= 1 && per_split * nsplit <= s.g_rows * s.LK, "tail scratch too small");
        simt::PartEpi e{s.G, d.Kp, rows, d.Kp, per_split};
        if (d.Kp <= 64) {
          dim3 g2((unsigned)m_tiles, 1, (unsigned)nsplit);
          simt::shift_gemm_kernel<128, 64, 16, 8, 4><<<g2, 256, 0, stream>>>(a, b, e, R, r_chunk, 1);
        } else {
          dim3 grid((unsigned)m_tiles, (unsigned)ceil_div_ll(d.Kp, 128), (unsigned)nsplit);
          simt::shift_gemm_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, stream>>>(a, b, e, R, r_chunk, 1);
        }
        CMF_TRY(launch_ok("tail_den_h_parts"));
        simt::tail_sub_kernel<<<ew_blocks(s, rows * d.Kp * 8), 256, 0, stream>>>(den, s.G, (int)nsplit, per_split, rows, d.Kp, m_off, d.Tloc);
        CMF_TRY(launch_ok("tail_den_h"));
      }
    }
  }
  return 0;
}

inline int h_terms(TcState& s, cudaStream_t stream) {
  const Dims& d = s.d;
  const Fold& f = s.f;
  HTermsParams p{};
  p.Np = d.Np; p.J = f.J; p.n_chunks_n = (int)ceil_div_ll(d.Np, 32); p.wrows = f.hterms_wrows;
  p.s = f.s; p.CB = f.CB; p.Kp = d.Kp;
  p.n_src = (s.gram & 1) ? 1 : 2;                 // the Gram route contracts X only
  p.n_time_tiles = d.TO / 256 + 1;
  p.n_split = s.h_split;
  p.n_items = p.n_time_tiles * p.n_src * p.n_split;
  p.TO = d.TO; p.out = s.h_split > 1 ? s.hparts : s.hterms; p.carry = s.hcarry; p.hd = f.h_hd;
  p.sub_units = s.h_sub; p.n_stages = f.h_stages; p.staged = f.h_staged;
  p.x3 = s.x3; p.lo_off = f.KW; p.err = s.d_err;
  const bool direct = f.n_glag == 1 && f.s == 1;
  tc_hterms_kernel<<<s.hterms_grid, kSThreads, hterms_smem_bytes(f.h_stages, f.hterms_wrows, d.Kp, f.h_hd, direct, f.h_staged != 0), stream>>>(
      s.tmW_k3, s.tmX_k3, s.tmE_k3, s.tmXlo_k3, s.tmElo_k3, p);
  CMF_TRY(launch_ok("tc_hterms"));
  if (s.h_split > 1) {
    // out[src] = sum of the partial outputs of its items (fixed order)
    const long long n4 = d.TO * d.Kp / 4;
    for (int src = 0; src < p.n_src; ++src) {
      ew::sum_splits_kernel<<<ew_blocks(s, n4), 256, 0, stream>>>((float4*)(s.hterms + (size_t)src * d.TO * d.Kp),
                                                                 (const float4*)(s.hparts + (size_t)src * s.h_split * d.TO * d.Kp),
                                                                 n4, n4, s.h_split);
      CMF_TRY(launch_ok("h_terms_sum"));
    }
  }
  if (f.h_hd > 0) {
    const long long wmax = f.h_hd < 256 ? f.h_hd : 256;
    const long long per_src = (p.n_time_tiles - 1) * wmax * (d.Kp / 4);
    if (per_src > 0) {
      if (s.h_split == 1) {
        hterms_carry_kernel<<<ew_blocks(s, per_src * p.n_src), 256, 0, stream>>>(s.hterms, s.hcarry, d.TO, p.n_time_tiles, f.h_hd, d.Kp, p.n_src);
        CMF_TRY(launch_ok("hterms_carry"));
      } else {
        const size_t per_carry = (size_t)p.n_time_tiles * f.h_hd * d.Kp;
        for (int src = 0; src < p.n_src; ++src)
          for (int sp = 0; sp < s.h_split; ++sp) {
            hterms_carry_kernel<<<ew_blocks(s, per_src), 256, 0, stream>>>(s.hterms + (size_t)src * d.TO * d.Kp,
                                                                          s.hcarry + (size_t)(src * s.h_split + sp) * per_carry,
                                                                          d.TO, p.n_time_tiles, f.h_hd, d.Kp, 1);
            CMF_TRY(launch_ok("hterms_carry"));
          }
      }
    }
  }
  if (s.gram & 1) CMF_TRY(den_h_gram(s, stream));
  return 0;
}

// Device-side pipeline errors (bounded waits that expired) surface here.
inline int check(TcState& s, cudaStream_t stream) {
  if (!s.ready) return 0;
  int e = 0;
  CMF_CUDA(cudaMemcpyAsync(&e, s.d_err, 4, cudaMemcpyDeviceToHost, stream));
  CMF_CUDA(cudaStreamSynchronize(stream));
  if (e != 0) {         // a runtime failure, not an argument error
    set_error("tensor-core kernel pipeline error %d (a barrier wait timed out)", e);
    return 1;
  }
  return 0;
}

}  // namespace tc
}  // namespace cmf
