// C ABI of the cmf_b200 library (see include/cmf_b200.h for the contract and the
// reference interfaces each entry point replaces).
//
// Device data layout (all fp32, all "time-major": the time index is the slow
// axis, so a lag is a whole-row offset and every shifted operand window of the
// three contractions is 16-byte aligned):
//   Xt, Et : RT x Np          row tau = local time, tau in [0, RT)
//   Ht     : (h + RT) x Kp    row tau + h, h = L-1 zero/halo rows in front
//   W      : L x Np x Kp      (the reference's own order, padded)
// Np = N rounded up to 4, Kp = K rounded up to 8.  Rows past the valid data
// are zeros: the reference's "drop terms that fall off the matrix"
// (common.py:24-27, :83-84) becomes "read a zero".
#include "../../include/cmf_b200.h"

#include <cstdarg>
#include <cstdlib>
#include <cmath>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "dev_cache.cuh"
#include "ew_kernels.cuh"
#include "simt_gemm.cuh"
#include "tc_path.cuh"
#include "peer_kernels.cuh"
#include "gd_kernels.cuh"
#include "hals_kernels.cuh"
#include "dataset_kernels.cuh"

namespace cmf {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
}

}  // namespace cmf

using namespace cmf;

struct cmf_mu_s {
  cmf_mu_params p{};
  int N = 0, K = 0, L = 0, Np = 0, Kp = 0, h = 0;
  long long Tloc = 0, TO = 0, RT = 0, RH = 0, t_valid = 0;
  int dev = 0, num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int precision = CMF_PREC_FP32;
  int round_ops = 0;                 // store operands pre-rounded to TF32
  bool use_tc = false;
  bool x3 = false;                   // CMF_PREC_TF32X3: Xt / Et hold TF32 hi halves, Xlo / Elo the lo halves
  float *Xlo = nullptr, *Elo = nullptr;

  float *Xt = nullptr, *Et = nullptr, *Ht = nullptr, *W = nullptr;   // Ht, W: fp32 masters (the TF32 operand
                                                                     // copies of the tcgen05 path live in tcs)
  float *numden = nullptr, *wpart = nullptr, *hterms = nullptr;
  long long wcount = 0;              // L * Np * Kp
  int wsplits = 1;
  long long wchunk = 0;

  double *loss_partials = nullptr, *d_sumsq = nullptr, *d_ring = nullptr, *d_xpart = nullptr;
  long long n_loss_partials = 0;
  int ring_cap = 0;
  int* d_neg = nullptr;

  double sumsq_x = 0.0, norm_x = 0.0;
  int has_neg = 0;
  bool have_data = false, have_factors = false, est_valid = false, wterms_valid = false;
  bool sumsq_current = true;         // d_sumsq is the residual of the factors est_valid refers to (do_recon)
  bool est_stored = false;           // the est buffer holds the reconstruction (est_valid alone: only its loss is current)

  long long launches = 0;
  int profiling = 0;
  int loss_mode = 0;                 // 0: auto (the cheapest loss evaluation whose error bound allows it), 1: always the
                                     // full-precision residual, 2: from the W terms whenever the route is full Gram
  bool loss_gram = false;            // this batch takes the loss from the W terms (decide_loss_mode)
  double last_loss = -1.0;           // the most recent loss the host has seen (< 0: none yet)
  // one MU iteration captured as a CUDA graph (launch-bound small problems; replayed by cmf_mu_step)
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_dirty = true, graph_ok = true;
  long long graph_launches = 0;
  cudaGraphExec_t sgraph_exec = nullptr;           // the same for one iteration of cmf_mu_step_sharded
  bool sgraph_dirty = true, sgraph_ok = true;
  long long sgraph_launches = 0;
  int* d_counter = nullptr;
  float kernel_ms[4] = {0, 0, 0, 0};
  std::vector<cudaEvent_t> ev_pool;
  tc::LaunchLog llog;                              // profiling == 2: an event after every launch
  std::vector<std::pair<std::string, std::pair<long long, double>>> launch_table;   // label -> (launches, ms)

  tc::TcState tcs;

  // staging for host <-> device conversions (grown on demand, kept: cudaMalloc / cudaFree in the read-back
  // path cost up to 0.7 s per call on a busy driver)
  void* stage_buf = nullptr;
  size_t stage_bytes = 0;

  // ---- gradient solvers (gd_kernels.cuh) ----
  struct GdState {
    bool ready = false, cached = false;  // cached: numden / hterms hold the terms of the CURRENT W, H
    float *P = nullptr, *Ppart = nullptr, *Pt = nullptr, *v = nullptr, *y = nullptr, *d_inv = nullptr;
    double* d_lam = nullptr;
    gd::PowerState* st = nullptr;
    int batch_iter = 64, max_batches = 16;   // power iteration: batches of launches, `done` read back in between
    int iters = 0;                           // iterations issued by the last lipschitz_W
    bool settled = true;                     // ... and whether its Rayleigh quotient settled (2 x rel. change <= 1e-7)
  } gdst;

  // ---- HALS (hals_kernels.cuh) ----
  struct HalsState {
    bool ready = false, open = false;    // open: between cmf_hals_begin and cmf_hals_end (the residual is current)
    float *Rt = nullptr, *part = nullptr, *part_h = nullptr, *delta = nullptr, *Wk = nullptr, *prev = nullptr;
    double *w2 = nullptr, *d_diff = nullptr;
    int n_chunks = 1, rows_per_chunk = 1;
  } halsst;

  // ---- peer-memory collectives of the sharded iteration (peer_kernels.cuh) ----
  struct PeerState {
    bool attached = false;
    void* shared = nullptr;              // own allocation: Control | halo staging (2 x h x Kp) | ring
    size_t shared_bytes = 0;
    void* opened[peer::kMaxPeers][3] = {};   // IPC mappings of the peers' allocations (bases, for close)
    peer::Peers P{};
    uint32_t bar_epoch = 0;                 // (exchange / halo epochs live on the device: peer::Control)
    int ring_cap = 1024;
  } peer;
};

// what a rank publishes to its peers (cmf_mu_peer_export); plain bytes, CMF_PEER_BLOB_BYTES at most
struct PeerBlob {
  uint32_t magic;
  int dev, h, Kp, ring_cap;
  long long wcount;
  cudaIpcMemHandle_t handle[3];          // allocations holding numden, W, the shared block
  unsigned long long offset[3];          // of those pointers inside their allocations
};
static_assert(sizeof(PeerBlob) <= CMF_PEER_BLOB_BYTES, "peer blob too large");
constexpr uint32_t kPeerMagic = 0x434d4650u;

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define CMF_ENTER(hh)                                              \
  CMF_CHECK((hh) != nullptr, "null solver handle");                \
  DeviceGuard _guard((hh)->dev);                                   \
  CMF_CHECK(_guard.ok, "cannot select CUDA device %d", (hh)->dev)

inline int ew_grid(const cmf_mu_s* h, long long n_items) {
  long long blocks = ceil_div_ll(n_items, 256);
  long long cap = (long long)h->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <class T> int dmalloc(T** p, long long count) {
  return cached_malloc((void**)p, (size_t)(count > 0 ? count : 1) * sizeof(T));
}
// (W and the W-term buffer: their CUDA-IPC handles go to the peer processes of a sharded solve)
template <class T> int dmalloc_plain(T** p, long long count) {
  CMF_CUDA(cudaMalloc((void**)p, (size_t)(count > 0 ? count : 1) * sizeof(T)));
  return 0;
}

// the handle's staging buffer, at least `bytes` large (the previous contents are not kept)
int stage_get(cmf_mu_s* h, size_t bytes, void** out) {
  if (h->stage_bytes < bytes) {
    CMF_CUDA(cudaStreamSynchronize(h->stream));
    cached_free(h->stage_buf);
    h->stage_buf = nullptr;
    h->stage_bytes = 0;
    CMF_CUDA(cudaMalloc(&h->stage_buf, bytes));
    h->stage_bytes = bytes;
  }
  *out = h->stage_buf;
  return 0;
}

int launch_check(cmf_mu_s* h, const char* what) {
  h->launches++;
  tc::log_launch(what);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---- contraction launches (fp32 FFMA path) ------------------------------
int simt_recon(cmf_mu_s* h) {
  simt::ReconA a{h->Ht, h->Kp, h->h};
  simt::ReconB b{h->W, h->Np, h->Kp};
  simt::ReconEpi e{h->Et, h->Xt, h->loss_partials, h->Np, h->Tloc, h->t_valid, h->round_ops};
  const long long R = (long long)h->L * h->Kp;
  long long nblocks;
  if (h->Np > 64) {
    dim3 grid((unsigned)(h->RT / 128), (unsigned)ceil_div_ll(h->Np, 128), 1);
    nblocks = (long long)grid.x * grid.y;
    simt::shift_gemm_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, h->stream>>>(a, b, e, R, R, 1);
  } else {
    dim3 grid((unsigned)(h->RT / 128), 1, 1);
    nblocks = grid.x;
    simt::shift_gemm_kernel<128, 64, 16, 8, 4><<<grid, 256, 0, h->stream>>>(a, b, e, R, R, 1);
  }
  CMF_TRY(launch_check(h, "recon"));
  ew::sum_doubles_kernel<<<1, 1024, 0, h->stream>>>(h->loss_partials, nblocks, h->d_sumsq);
  return launch_check(h, "loss_sum");
}

int simt_w_terms(cmf_mu_s* h) {
  const int LKp = h->L * h->Kp;
  simt::WTermsA a{h->Xt, h->Et, h->Np};
  simt::WTermsB b{h->Ht, h->Kp, h->h, LKp};
  float* part = (h->wsplits == 1) ? h->numden : h->wpart;
  simt::WTermsEpi e{part, h->Np, h->Kp, LKp, h->wcount};
  dim3 grid((unsigned)ceil_div_ll(h->Np, 128), (unsigned)ceil_div_ll(LKp, 128), (unsigned)(h->wsplits * 2));
  simt::shift_gemm_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, h->stream>>>(a, b, e, h->Tloc, h->wchunk, 2);
  CMF_TRY(launch_check(h, "w_terms"));
  if (h->wsplits > 1) {
    const long long n4 = 2 * h->wcount / 4;
    ew::sum_splits_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>(
        (float4*)h->numden, (const float4*)h->wpart, n4, n4, h->wsplits);
    CMF_TRY(launch_check(h, "w_terms_sum"));
  }
  return 0;
}

int simt_h_terms(cmf_mu_s* h) {
  simt::HTermsA a{h->Xt, h->Et, h->Np};
  simt::HTermsB b{h->W, h->Kp};
  simt::HTermsEpi e{h->hterms, h->Kp, h->TO * h->Kp};
  const long long R = (long long)h->L * h->Np;
  if (h->Kp <= 16) {
    dim3 grid((unsigned)(h->TO / 256), 1, 2);
    simt::shift_gemm_kernel<256, 16, 16, 4, 4><<<grid, 256, 0, h->stream>>>(a, b, e, R, R, 2);
  } else if (h->Kp <= 32) {
    dim3 grid((unsigned)(h->TO / 256), 1, 2);
    simt::shift_gemm_kernel<256, 32, 16, 8, 4><<<grid, 256, 0, h->stream>>>(a, b, e, R, R, 2);
  } else if (h->Kp <= 64) {
    dim3 grid((unsigned)(h->TO / 128), 1, 2);
    simt::shift_gemm_kernel<128, 64, 16, 8, 4><<<grid, 256, 0, h->stream>>>(a, b, e, R, R, 2);
  } else {
    dim3 grid((unsigned)(h->TO / 128), (unsigned)ceil_div_ll(h->Kp, 128), 2);
    simt::shift_gemm_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, h->stream>>>(a, b, e, R, R, 2);
  }
  return launch_check(h, "h_terms");
}

// ---- phase dispatch -----------------------------------------------------
bool gram_w(const cmf_mu_s* h) { return h->use_tc && (h->tcs.mask & 2) && (h->tcs.gram & 2); }
bool gram_h(const cmf_mu_s* h) { return h->use_tc && (h->tcs.mask & 4) && (h->tcs.gram & 1); }

int ensure_est_buffer(cmf_mu_s* h) {
  if (h->Et) return 0;
  CMF_TRY(big_alloc((void**)&h->Et, ((size_t)h->RT * h->Np + 128) * 4, h->dev));   // + slack: see make_map_k2src
  CMF_CUDA(cudaMemsetAsync(h->Et, 0, (size_t)h->RT * h->Np * 4, h->stream));
  if (h->x3) {
    CMF_TRY(big_alloc((void**)&h->Elo, ((size_t)h->RT * h->Np + 128) * 4, h->dev));
    CMF_CUDA(cudaMemsetAsync(h->Elo, 0, (size_t)h->RT * h->Np * 4, h->stream));
  }
  if (h->use_tc) CMF_TRY(tc::attach_est(h->tcs, h->Et, h->Elo));
  h->graph_dirty = h->sgraph_dirty = true;             // kernel arguments baked into a captured iteration changed
  return 0;
}

// store_est = false: only the loss is wanted (legal when neither MU step reads est)
// need_loss = false (the reconstruction between the W and the H step of one iteration): the residual is not formed
// (tc::recon); d_sumsq then belongs to older factors until the next full reconstruction (sumsq_current).
int do_recon(cmf_mu_s* h, bool store_est = true, bool need_loss = true) {
  CMF_CHECK(h->have_data && h->have_factors, "recon before data/factors were set");
  if (store_est || !(h->use_tc && (h->tcs.mask & 1))) CMF_TRY(ensure_est_buffer(h));
  if (h->use_tc && (h->tcs.mask & 1)) {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::recon(h->tcs, h->stream, store_est, need_loss));
    h->launches += tc::launch_counter() - n0;
    h->est_stored = store_est;
    h->sumsq_current = need_loss;
  } else {
    CMF_TRY(simt_recon(h));
    h->est_stored = true;
    h->sumsq_current = true;
  }
  h->est_valid = true;
  return 0;
}
int ensure_est_stored(cmf_mu_s* h) {
  if (h->est_valid && h->est_stored) return 0;
  return do_recon(h, true);
}

int do_w_terms(cmf_mu_s* h) {
  CMF_CHECK(h->have_data && h->have_factors, "w_terms before data/factors were set");
  CMF_CHECK(h->est_valid || gram_w(h), "w_terms needs a current reconstruction (call cmf_mu_recon)");
  if (!gram_w(h)) CMF_TRY(ensure_est_stored(h));
  if (h->use_tc && (h->tcs.mask & 2)) {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::w_terms(h->tcs, h->stream));
    h->launches += tc::launch_counter() - n0;
  } else {
    CMF_TRY(simt_w_terms(h));
  }
  h->wterms_valid = true;
  return 0;
}
int do_w_apply(cmf_mu_s* h) {
  CMF_CHECK(h->wterms_valid, "w_apply before w_terms");
  const long long n4 = h->wcount / 4;
  ew::mu_update_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>(
      (float4*)h->W, (const float4*)h->numden, (const float4*)(h->numden + h->wcount), n4,
      (float4*)tc::fused_w_op(h->tcs));
  CMF_TRY(launch_check(h, "w_update"));
  {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::refresh_w(h->tcs, h->stream, true));
    h->launches += tc::launch_counter() - n0;
  }
  h->wterms_valid = false;
  h->est_valid = false;
  return 0;
}
int do_h_terms(cmf_mu_s* h) {
  CMF_CHECK(h->have_data && h->have_factors, "h terms before data/factors were set");
  CMF_CHECK(h->est_valid || gram_h(h), "h terms need a current reconstruction (call cmf_mu_recon)");
  if (!gram_h(h)) CMF_TRY(ensure_est_stored(h));
  if (h->use_tc && (h->tcs.mask & 4)) {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::h_terms(h->tcs, h->stream));
    h->launches += tc::launch_counter() - n0;
    return 0;
  }
  return simt_h_terms(h);
}
int do_h_apply(cmf_mu_s* h) {
  const long long n4 = h->Tloc * h->Kp / 4;
  ew::mu_update_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>(
      (float4*)(h->Ht + (long long)h->h * h->Kp), (const float4*)h->hterms,
      (const float4*)(h->hterms + h->TO * h->Kp), n4,
      tc::fused_h_op(h->tcs) ? (float4*)(tc::fused_h_op(h->tcs) + (long long)h->h * h->Kp) : nullptr);
  CMF_TRY(launch_check(h, "h_update"));
  {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::refresh_h(h->tcs, h->stream, h->h, h->Tloc, true));
    h->launches += tc::launch_counter() - n0;
  }
  h->est_valid = false;
  return 0;
}

// refresh the TF32 operand copies from the fp32 masters (no-op on the fp32 path)
int sync_ops_W(cmf_mu_s* h) {
  const long long n0 = tc::launch_counter();
  const int rc = tc::refresh_w(h->tcs, h->stream);
  h->launches += tc::launch_counter() - n0;
  return rc;
}
int sync_ops_H(cmf_mu_s* h, long long row0, long long nrows) {
  const long long n0 = tc::launch_counter();
  const int rc = tc::refresh_h(h->tcs, h->stream, row0, nrows);
  h->launches += tc::launch_counter() - n0;
  return rc;
}

// profiling == 2: install / harvest the per-launch log around a batch of iterations (the caller has synchronised)
void launch_log_begin(cmf_mu_s* h) {
  if (h->profiling < 2) return;
  h->llog.stream = h->stream;
  h->llog.labels.clear();
  tc::launch_log() = &h->llog;
  tc::log_launch("<start>");
}
void launch_log_end(cmf_mu_s* h) {
  if (tc::launch_log() != &h->llog) return;
  tc::launch_log() = nullptr;
  auto& L = h->llog;
  for (size_t i = 1; i < L.labels.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, L.pool[i - 1], L.pool[i]) != cudaSuccess) { cudaGetLastError(); continue; }
    bool found = false;
    for (auto& row : h->launch_table)
      if (row.first == L.labels[i]) { row.second.first++; row.second.second += ms; found = true; break; }
    if (!found) h->launch_table.push_back({L.labels[i], {1, (double)ms}});
  }
  L.labels.clear();
}

cudaEvent_t get_event(cmf_mu_s* h, size_t i) {
  while (h->ev_pool.size() <= i) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[i];
}

// staging copy of a host/device 2-D block (rows x cols elements of es bytes)
// into a dense device buffer
int stage_block(const void* src, int mem, long long ld, long long rows, long long cols,
                size_t es, void* dst, cudaStream_t s) {
  CMF_CUDA(cudaMemcpy2DAsync(dst, (size_t)cols * es, src, (size_t)ld * es, (size_t)cols * es,
                             (size_t)rows, mem == CMF_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
  return 0;
}

template <class TI>
int load_transposed(cmf_mu_s* h, const TI* src, int mem, long long ld, long long rows, long long cols,
                    float* dst, long long ldd, int round_in) {
  // dst[c][r] = src[r][c]; host sources go through a bounded device staging buffer
  if (rows == 0 || cols == 0) return 0;
  if (mem == CMF_DEVICE) {
    dim3 grid((unsigned)ceil_div_ll(cols, 32), (unsigned)ceil_div_ll(rows, 32));
    CMF_CHECK(grid.y <= 65535, "too many rows for the transpose grid");
    ew::transpose_convert_kernel<TI, float><<<grid, 256, 0, h->stream>>>(src, ld, dst, ldd, rows, cols, round_in);
    return launch_check(h, "transpose_in");
  }
  const long long budget = 64ll << 20;
  long long ch = budget / (long long)(rows * sizeof(TI));
  ch = (ch / 32) * 32;
  if (ch < 32) ch = 32;
  if (ch > cols) ch = cols;
  TI* stg = nullptr;
  CMF_TRY(stage_get(h, (size_t)rows * ch * sizeof(TI), (void**)&stg));
  int rc = 0;
  for (long long c0 = 0; c0 < cols && rc == 0; c0 += ch) {
    const long long w = (cols - c0 < ch) ? cols - c0 : ch;
    rc = stage_block(src + c0, CMF_HOST, ld, rows, w, sizeof(TI), stg, h->stream);
    if (rc) break;
    dim3 grid((unsigned)ceil_div_ll(w, 32), (unsigned)ceil_div_ll(rows, 32));
    ew::transpose_convert_kernel<TI, float><<<grid, 256, 0, h->stream>>>(stg, w, dst + c0 * ldd, ldd, rows, w, round_in);
    rc = launch_check(h, "transpose_in");
    // the staging buffer is reused by the next chunk
    if (rc == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("sync failed in load_transposed"); rc = 1; }
  }
  return rc;
}

template <class TO>
int store_transposed(cmf_mu_s* h, const float* src, long long lds, long long rows_out, long long cols_out,
                     TO* dst, int mem, long long ldd, const float* src_lo = nullptr) {
  // src_lo (optional): a second array of the same layout that is added on the way out (3xTF32 hi + lo)
  // dst[r][c] = src[c][r], dst is rows_out x cols_out (ld ldd); src is cols_out x lds
  if (rows_out == 0 || cols_out == 0) return 0;
  if (mem == CMF_DEVICE) {
    dim3 grid((unsigned)ceil_div_ll(rows_out, 32), (unsigned)ceil_div_ll(cols_out, 32));
    CMF_CHECK(grid.y <= 65535 * 32ll, "too many columns");
    // source viewed as (cols_out x rows_out); grid.x walks its columns (= rows_out)
    if (grid.y > 65535) {
      // walk in slabs of 65535*32 source rows
      const long long slab = 65535ll * 32;
      for (long long c0 = 0; c0 < cols_out; c0 += slab) {
        const long long w = (cols_out - c0 < slab) ? cols_out - c0 : slab;
        dim3 g((unsigned)ceil_div_ll(rows_out, 32), (unsigned)ceil_div_ll(w, 32));
        ew::transpose_convert_kernel<float, TO><<<g, 256, 0, h->stream>>>(src + c0 * lds, lds, dst + c0, ldd, w, rows_out, 0,
                                                                          src_lo ? src_lo + c0 * lds : nullptr);
        CMF_TRY(launch_check(h, "transpose_out"));
      }
      return 0;
    }
    ew::transpose_convert_kernel<float, TO><<<grid, 256, 0, h->stream>>>(src, lds, dst, ldd, cols_out, rows_out, 0, src_lo);
    return launch_check(h, "transpose_out");
  }
  const long long budget = 64ll << 20;
  long long ch = budget / (long long)(rows_out * sizeof(TO));
  ch = (ch / 32) * 32;
  if (ch < 32) ch = 32;
  if (ch > cols_out) ch = cols_out;
  TO* stg = nullptr;
  CMF_TRY(stage_get(h, (size_t)rows_out * ch * sizeof(TO), (void**)&stg));
  int rc = 0;
  for (long long c0 = 0; c0 < cols_out && rc == 0; c0 += ch) {
    const long long w = (cols_out - c0 < ch) ? cols_out - c0 : ch;
    dim3 grid((unsigned)ceil_div_ll(rows_out, 32), (unsigned)ceil_div_ll(w, 32));
    ew::transpose_convert_kernel<float, TO><<<grid, 256, 0, h->stream>>>(src + c0 * lds, lds, stg, w, w, rows_out, 0,
                                                                         src_lo ? src_lo + c0 * lds : nullptr);
    rc = launch_check(h, "transpose_out");
    if (rc) break;
    if (cudaMemcpy2DAsync(dst + c0, (size_t)ldd * sizeof(TO), stg, (size_t)w * sizeof(TO), (size_t)w * sizeof(TO),
                          (size_t)rows_out, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
      set_error("device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
    }
  }
  return rc;
}

int peer_detach(cmf_mu_s* h) {
  auto& ps = h->peer;
  if (ps.attached) cudaStreamSynchronize(h->stream);
  for (int p = 0; p < peer::kMaxPeers; ++p)
    for (int k = 0; k < 3; ++k)
      if (ps.opened[p][k]) { cudaIpcCloseMemHandle(ps.opened[p][k]); ps.opened[p][k] = nullptr; }
  ps.attached = false;
  return 0;
}

void free_all(cmf_mu_s* h) {
  // nothing of this solver may still run when its blocks go back to the cache (cudaFree would have waited; the
  // cache does not)
  if (h->stream) cudaStreamSynchronize(h->stream);
  peer_detach(h);
  cached_free(h->peer.shared);
  cached_free(h->stage_buf);
  cached_free(h->gdst.P); cached_free(h->gdst.Ppart); cached_free(h->gdst.Pt); cached_free(h->gdst.v); cached_free(h->gdst.y);
  cached_free(h->gdst.d_inv); cached_free(h->gdst.d_lam); cached_free(h->gdst.st);
  cached_free(h->halsst.Rt); cached_free(h->halsst.part); cached_free(h->halsst.part_h); cached_free(h->halsst.delta);
  cached_free(h->halsst.Wk); cached_free(h->halsst.prev); cached_free(h->halsst.w2); cached_free(h->halsst.d_diff);
  tc::destroy(h->tcs);
  {
    const size_t nt_bytes = ((size_t)h->RT * h->Np + 128) * 4;
    big_free(h->Xt, nt_bytes, h->dev); big_free(h->Et, nt_bytes, h->dev);
    big_free(h->Xlo, nt_bytes, h->dev); big_free(h->Elo, nt_bytes, h->dev);
  }
  cached_free(h->Ht); cached_free(h->W);
  cached_free(h->numden); cached_free(h->wpart); cached_free(h->hterms);
  cached_free(h->loss_partials); cached_free(h->d_sumsq); cached_free(h->d_ring); cached_free(h->d_xpart);
  cached_free(h->d_neg);
  cached_free(h->d_counter);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->sgraph_exec) cudaGraphExecDestroy(h->sgraph_exec);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  for (auto e : h->llog.pool) cudaEventDestroy(e);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
}

}  // namespace

// ==========================================================================
extern "C" {

int cmf_abi_version(void) { return CMF_B200_ABI_VERSION; }

int cmf_release_cached_memory(void) {
  release_cached_blocks();
  return 0;
}

const char* cmf_last_error(void) { return last_error().c_str(); }

int cmf_device_count(int* count) {
  CMF_CHECK(count != nullptr, "null argument");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  return 0;
}

int cmf_precision_supported(int precision, int n_features, int n_components, int maxlag) {
  if (precision == CMF_PREC_FP32) return 1;
  if (precision == CMF_PREC_TF32 || precision == CMF_PREC_TF32X3)
    return tc::shape_supported(n_features, n_components, maxlag) ? 1 : 0;
  return 0;
}

int cmf_mu_create(cmf_mu_t** out, const cmf_mu_params* p) {
  CMF_CHECK(out != nullptr && p != nullptr, "null argument");
  *out = nullptr;
  CMF_CHECK(p->n_features >= 1 && p->n_components >= 1 && p->maxlag >= 1, "dimensions must be positive");
  CMF_CHECK(p->t_local >= 1 && p->t_global >= p->t_local && p->t_offset >= 0 &&
                p->t_offset + p->t_local <= p->t_global,
            "inconsistent time range: t_local=%lld t_offset=%lld t_global=%lld", p->t_local, p->t_offset, p->t_global);
  CMF_CHECK(p->precision == CMF_PREC_FP32 || p->precision == CMF_PREC_TF32 || p->precision == CMF_PREC_TF32X3,
            "unknown precision %d", p->precision);
  CMF_CHECK(p->denominators == CMF_DEN_DIRECT || p->denominators == CMF_DEN_GRAM || p->denominators == CMF_DEN_AUTO,
            "unknown denominators mode %d", p->denominators);
  CMF_CHECK(p->t_local == p->t_global || p->t_local >= p->maxlag - 1,
            "a time shard must hold at least L-1 columns (t_local=%lld, L=%d)", p->t_local, p->maxlag);
  int ndev = 0;
  CMF_CUDA(cudaGetDeviceCount(&ndev));
  CMF_CHECK(p->device >= 0 && p->device < ndev, "device %d out of range (%d visible)", p->device, ndev);

  cmf_mu_s* h = new cmf_mu_s();
  h->p = *p;
  h->dev = p->device;
  DeviceGuard guard(h->dev);
  if (!guard.ok) { delete h; set_error("cannot select CUDA device %d", p->device); return 1; }

  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->dev) != cudaSuccess) { delete h; set_error("cudaGetDeviceProperties failed"); return 1; }
  if (prop.major != 10) {
    delete h;
    set_error("cmf_b200 is built for sm_100a only; device %d is sm_%d%d", p->device, prop.major, prop.minor);
    return 3;
  }
  h->num_sms = prop.multiProcessorCount;

  h->N = p->n_features; h->K = p->n_components; h->L = p->maxlag;
  h->use_tc = (p->precision != CMF_PREC_FP32);
  h->x3 = (p->precision == CMF_PREC_TF32X3);
  if (h->use_tc && !tc::shape_supported(h->N, h->K, h->L)) {
    delete h;
    set_error("precision tf32 / tf32x3 has no tensor-core kernel for N=%d K=%d L=%d; use fp32", p->n_features, p->n_components, p->maxlag);
    return 2;
  }
  h->Np = round_up(h->N, 4); h->h = h->L - 1;
  h->Kp = h->use_tc ? tc::padded_k(h->K) : round_up(h->K, 8);
  h->Tloc = p->t_local;
  h->TO = round_up_ll(h->Tloc, 256);
  h->RT = round_up_ll(h->TO + h->h, 256);
  h->RH = h->h + h->RT;
  {
    long long remaining = p->t_global - p->t_offset;       // columns that exist from t_offset on
    long long want = h->Tloc + h->h;
    h->t_valid = remaining < want ? remaining : want;
  }
  h->precision = p->precision;
  h->round_ops = h->use_tc ? 1 : 0;

  if (p->stream) {
    h->stream = (cudaStream_t)p->stream;
  } else {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete h; set_error("cudaStreamCreate failed"); return 1;
    }
    h->own_stream = true;
  }

  h->wcount = (long long)h->L * h->Np * h->Kp;
  // split the W-term reduction over time so that the grid fills the machine
  {
    long long tiles = ceil_div_ll(h->Np, 128) * ceil_div_ll((long long)h->L * h->Kp, 128) * 2;
    long long want = ceil_div_ll(2ll * h->num_sms, tiles);
    long long max_by_len = h->Tloc / 512; if (max_by_len < 1) max_by_len = 1;
    long long max_by_mem = (1ll << 30) / (2 * h->wcount * 4); if (max_by_mem < 1) max_by_mem = 1;
    long long s = want;
    if (s > max_by_len) s = max_by_len;
    if (s > max_by_mem) s = max_by_mem;
    if (s > 4096) s = 4096;
    if (s < 1) s = 1;
    h->wchunk = round_up_ll(ceil_div_ll(h->Tloc, s), 16);
    h->wsplits = (int)ceil_div_ll(h->Tloc, h->wchunk);
  }

  int rc = 0;
  auto A = [&](int r) { if (rc == 0) rc = r; };
  A(big_alloc((void**)&h->Xt, ((size_t)h->RT * h->Np + 128) * 4, h->dev));   // + slack: the K2 box of a ragged feature count reads on
  if (h->x3) A(big_alloc((void**)&h->Xlo, ((size_t)h->RT * h->Np + 128) * 4, h->dev));
  A(dmalloc(&h->Ht, h->RH * h->Kp));
  A(dmalloc_plain(&h->W, h->wcount));
  A(dmalloc_plain(&h->numden, 2 * h->wcount));
  if (h->wsplits > 1) A(dmalloc(&h->wpart, 2 * h->wcount * h->wsplits));
  A(dmalloc(&h->hterms, 2 * h->TO * h->Kp));
  h->n_loss_partials = (h->RT / 128) * ceil_div_ll(h->Np, 64) + 1024;
  A(dmalloc(&h->loss_partials, h->n_loss_partials));
  A(dmalloc(&h->d_sumsq, 1));
  h->ring_cap = 1024;
  A(dmalloc(&h->d_ring, h->ring_cap));
  A(dmalloc(&h->d_xpart, (long long)h->num_sms * 8));
  A(dmalloc(&h->d_neg, 1));
  A(dmalloc(&h->d_counter, 1));
  if (rc == 0) {
    cudaError_t e = cudaSuccess;
    auto Z = [&](void* ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMemsetAsync(ptr, 0, bytes, h->stream); };
    Z(h->Xt, (size_t)h->RT * h->Np * 4);
    if (h->x3) Z(h->Xlo, (size_t)h->RT * h->Np * 4);
    Z(h->Ht, (size_t)h->RH * h->Kp * 4);
    Z(h->W, (size_t)h->wcount * 4);
    Z(h->numden, (size_t)2 * h->wcount * 4);
    Z(h->hterms, (size_t)2 * h->TO * h->Kp * 4);
    Z(h->d_sumsq, 8);
    Z(h->d_neg, 4);
    if (e != cudaSuccess) { set_error("cudaMemset failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  if (rc == 0 && h->use_tc) {
    tc::Dims d{h->N, h->K, h->L, h->Np, h->Kp, h->h, h->Tloc, h->TO, h->RT, h->RH, h->t_valid, h->num_sms};
    {
      const double contraction_flops = 2.0 * h->N * h->K * (double)h->L * (double)h->Tloc;
      // auto: the Gram operators cost about K / N (W step) and 2 K / N (H step) of the direct contractions and
      // add T-independent small kernels: worth it for large shards with N >= 4 K
      const bool want = p->denominators == CMF_DEN_GRAM ||
                        (p->denominators == CMF_DEN_AUTO && contraction_flops >= 2e11 && h->N >= 4 * h->K);
      h->tcs.gram_request = want ? 3 : 0;
    }
    rc = tc::init(h->tcs, d, h->Xt, nullptr, h->Ht, h->W, h->numden, h->hterms, h->loss_partials,
                  h->n_loss_partials, h->d_sumsq, h->stream, h->Xlo);
  }
  // est^T: needed up front unless both denominators come from the Gram route (then on first demand)
  if (rc == 0 && !(h->use_tc && h->tcs.mask == 7 && h->tcs.gram == 3)) rc = ensure_est_buffer(h);
  if (rc != 0) {
    free_all(h);
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int cmf_mu_destroy(cmf_mu_t* h) {
  if (!h) return 0;
  DeviceGuard guard(h->dev);
  cudaStreamSynchronize(h->stream);
  free_all(h);
  delete h;
  return 0;
}

// X^T is in place at full precision (x3: unsplit): local ||X||^2 over owned columns, the negativity flag, and the
// hi/lo split of the 3xTF32 mode
static int finish_data(cmf_mu_s* h) {
  const long long n4 = h->Tloc * h->Np / 4;
  const int grid = ew_grid(h, n4);
  CMF_CUDA(cudaMemsetAsync(h->d_neg, 0, 4, h->stream));
  ew::sumsq_neg_kernel<<<grid, 256, 0, h->stream>>>((const float4*)h->Xt, n4, h->d_xpart, h->d_neg);
  CMF_TRY(launch_check(h, "sumsq_x"));
  ew::sum_doubles_kernel<<<1, 1024, 0, h->stream>>>(h->d_xpart, grid, h->d_sumsq);
  CMF_TRY(launch_check(h, "sumsq_x_final"));
  if (h->x3) {
    const long long m4 = h->RT * h->Np / 4;
    tc::split_inplace_kernel<<<ew_grid(h, m4), 256, 0, h->stream>>>((float4*)h->Xt, (float4*)h->Xlo, m4);
    CMF_TRY(launch_check(h, "split_x"));
  }
  CMF_CUDA(cudaMemcpyAsync(&h->sumsq_x, h->d_sumsq, 8, cudaMemcpyDeviceToHost, h->stream));
  CMF_CUDA(cudaMemcpyAsync(&h->has_neg, h->d_neg, 4, cudaMemcpyDeviceToHost, h->stream));
  CMF_CUDA(cudaStreamSynchronize(h->stream));
  h->norm_x = std::sqrt(h->sumsq_x);
  h->graph_dirty = h->sgraph_dirty = true;
  h->last_loss = -1.0;
  h->have_data = true;
  h->est_valid = false;
  h->wterms_valid = false;
  return 0;
}

// Per-feature sums over the owned columns: s1[n] = sum_t X[n,t], s2[n] = sum_t X[n,t]^2, sabs[n] = sum_t |X[n,t]|
// (each N doubles, HOST; any may be NULL).  What the reference's dataset normalisations reduce
// (songbird.py:18-19, maze.py:71-72, vox_celeb.py:100-102); a sharded driver all-reduces them before scaling.
int cmf_mu_row_stats(cmf_mu_t* h, double* s1, double* s2, double* sabs) {
  CMF_ENTER(h);
  CMF_CHECK(h->have_data, "no data set");
  const int rows_per_chunk = 1024;
  const int nchunks = (int)ceil_div_ll(h->Tloc, rows_per_chunk);
  double *part = nullptr, *out = nullptr;
  CMF_TRY(dmalloc(&part, (long long)nchunks * 3 * h->Np));
  int rc = dmalloc(&out, 3ll * h->Np);
  std::vector<double> host((size_t)3 * h->Np);
  if (rc == 0) {
    dim3 grid((unsigned)ceil_div_ll(h->Np, 256), (unsigned)nchunks);
    ew::row_stats_kernel<<<grid, 256, 0, h->stream>>>(h->Xt, h->Xlo, h->Tloc, h->Np, rows_per_chunk, part);
    rc = launch_check(h, "row_stats");
  }
  if (rc == 0) {
    ew::row_stats_sum_kernel<<<(unsigned)ceil_div_ll(3ll * h->Np, 256), 256, 0, h->stream>>>(part, nchunks, h->Np, out);
    rc = launch_check(h, "row_stats_sum");
  }
  if (rc == 0 && (cudaMemcpyAsync(host.data(), out, host.size() * 8, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                  cudaStreamSynchronize(h->stream) != cudaSuccess)) { set_error("row stats read-back failed"); rc = 1; }
  cached_free(part); cached_free(out);
  if (rc) return rc;
  for (int n = 0; n < h->N; ++n) {
    if (s1) s1[n] = host[n];
    if (s2) s2[n] = host[(size_t)h->Np + n];
    if (sabs) sabs[n] = host[(size_t)2 * h->Np + n];
  }
  return 0;
}

// X[n, :] *= scale[n] on the device (owned columns and the static right halo), then the same bookkeeping as
// cmf_mu_set_data (local ||X||^2, negativity flag, hi/lo split).  scale: N doubles, HOST.
int cmf_mu_scale_rows(cmf_mu_t* h, const double* scale) {
  CMF_ENTER(h);
  CMF_CHECK(scale != nullptr, "null argument");
  CMF_CHECK(h->have_data, "no data set");
  std::vector<float> sc((size_t)h->Np, 0.f);
  for (int n = 0; n < h->N; ++n) sc[n] = (float)scale[n];
  float* d_sc = nullptr;
  CMF_TRY(dmalloc(&d_sc, h->Np));
  int rc = 0;
  if (cudaMemcpyAsync(d_sc, sc.data(), (size_t)h->Np * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { set_error("H2D copy failed"); rc = 1; }
  if (rc == 0) {
    const long long total = h->RT * h->Np;
    ew::scale_rows_kernel<<<ew_grid(h, total), 256, 0, h->stream>>>(h->Xt, h->Xlo, h->RT, h->Np, d_sc, h->x3 ? 0 : h->round_ops);
    rc = launch_check(h, "scale_rows");
  }
  if (rc == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("sync failed in scale_rows"); rc = 1; }
  cached_free(d_sc);
  if (rc) return rc;
  return finish_data(h);
}

int cmf_mu_set_data(cmf_mu_t* h, const void* X, int dtype, int mem, long long ld, long long ncols) {
  CMF_ENTER(h);
  CMF_CHECK(X != nullptr, "null data pointer");
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(mem == CMF_HOST || mem == CMF_DEVICE, "unknown memory space %d", mem);
  CMF_CHECK(ncols >= h->Tloc && ncols <= h->Tloc + h->h, "ncols=%lld must lie in [t_local, t_local+L-1] = [%lld, %lld]",
            ncols, h->Tloc, h->Tloc + h->h);
  CMF_CHECK(ld >= ncols, "leading dimension %lld < ncols %lld", ld, ncols);
  if (ncols > h->t_valid) ncols = h->t_valid;    // nothing exists past the global end
  CMF_CUDA(cudaMemsetAsync(h->Xt, 0, (size_t)h->RT * h->Np * 4, h->stream));
  const int round_in = h->x3 ? 0 : h->round_ops;
  if (dtype == CMF_F32) CMF_TRY(load_transposed<float>(h, (const float*)X, mem, ld, h->N, ncols, h->Xt, h->Np, round_in));
  else CMF_TRY(load_transposed<double>(h, (const double*)X, mem, ld, h->N, ncols, h->Xt, h->Np, round_in));
  return finish_data(h);
}

int cmf_mu_data_stats(cmf_mu_t* h, double* sumsq, int* has_negative) {
  CMF_ENTER(h);
  CMF_CHECK(h->have_data, "no data set");
  if (sumsq) *sumsq = h->sumsq_x;
  if (has_negative) *has_negative = h->has_neg;
  return 0;
}

int cmf_mu_set_norm_x(cmf_mu_t* h, double norm_x) {
  CMF_ENTER(h);
  CMF_CHECK(norm_x >= 0.0, "norm_x must be non-negative");
  h->norm_x = norm_x;
  h->graph_dirty = h->sgraph_dirty = true;
  return 0;
}

int cmf_mu_set_factors(cmf_mu_t* h, const void* W0, const void* H0, int dtype, int mem, long long ldh) {
  CMF_ENTER(h);
  h->last_loss = -1.0;            // new factors: no loss of theirs has been seen yet
  CMF_CHECK(W0 != nullptr && H0 != nullptr, "W or H not initalized.");   // base.py:59-60
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(mem == CMF_HOST || mem == CMF_DEVICE, "unknown memory space %d", mem);
  CMF_CHECK(ldh >= h->Tloc, "leading dimension of H %lld < t_local %lld", ldh, h->Tloc);
  const long long wn = (long long)h->L * h->N * h->K;
  const size_t es = dtype == CMF_F32 ? 4 : 8;
  const void* wsrc = W0;
  void* wstage = nullptr;
  if (mem == CMF_HOST) {
    CMF_TRY(stage_get(h, (size_t)wn * es, &wstage));
    if (cudaMemcpyAsync(wstage, W0, (size_t)wn * es, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
      set_error("H2D copy of W failed"); return 1;
    }
    wsrc = wstage;
  }
  const int grid = ew_grid(h, h->wcount);
  if (dtype == CMF_F32)
    ew::w_pad_in_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)wsrc, h->W, h->L, h->N, h->K, h->Np, h->Kp, 0);
  else
    ew::w_pad_in_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)wsrc, h->W, h->L, h->N, h->K, h->Np, h->Kp, 0);
  int rc = launch_check(h, "w_pad_in");
  if (rc == 0 && cudaMemsetAsync(h->Ht, 0, (size_t)h->RH * h->Kp * 4, h->stream) != cudaSuccess) { set_error("memset failed"); rc = 1; }
  if (rc == 0) {
    float* dst = h->Ht + (long long)h->h * h->Kp;
    if (dtype == CMF_F32) rc = load_transposed<float>(h, (const float*)H0, mem, ldh, h->K, h->Tloc, dst, h->Kp, 0);
    else rc = load_transposed<double>(h, (const double*)H0, mem, ldh, h->K, h->Tloc, dst, h->Kp, 0);
  }
  if (rc == 0) rc = sync_ops_W(h);
  if (rc == 0) rc = sync_ops_H(h, 0, h->RH);
  if (rc == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("sync failed in set_factors"); rc = 1; }
  if (rc) return rc;
  h->have_factors = true;
  h->est_valid = false;
  h->wterms_valid = false;
  return 0;
}

int cmf_mu_init_stats(cmf_mu_t* h, double* x_dot_est, double* est_sumsq) {
  CMF_ENTER(h);
  CMF_CHECK(x_dot_est != nullptr && est_sumsq != nullptr, "null argument");
  CMF_CHECK(h->have_data && h->have_factors, "init stats before data/factors were set");
  CMF_TRY(ensure_est_stored(h));
  const long long n4 = h->Tloc * h->Np / 4;
  int grid = ew_grid(h, n4);
  if (grid > h->num_sms * 4) grid = h->num_sms * 4;      // d_xpart holds 8*num_sms doubles
  ew::dot_sumsq_kernel<<<grid, 256, 0, h->stream>>>((const float4*)h->Xt, (const float4*)h->Et, n4, h->d_xpart,
                                                    (const float4*)h->Xlo, (const float4*)h->Elo);
  CMF_TRY(launch_check(h, "dot_sumsq"));
  std::vector<double> host((size_t)grid * 2);
  CMF_CUDA(cudaMemcpyAsync(host.data(), h->d_xpart, host.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  CMF_CUDA(cudaStreamSynchronize(h->stream));
  double a = 0.0, b = 0.0;
  for (int i = 0; i < grid; ++i) { a += host[2 * i]; b += host[2 * i + 1]; }
  *x_dot_est = a;
  *est_sumsq = b;
  return 0;
}

int cmf_mu_scale_factors(cmf_mu_t* h, double scale_w, double scale_h) {
  CMF_ENTER(h);
  h->last_loss = -1.0;            // new factors: no loss of theirs has been seen yet
  CMF_CHECK(h->have_factors, "W or H not initalized.");
  ew::scale_kernel<<<ew_grid(h, h->wcount / 4), 256, 0, h->stream>>>((float4*)h->W, h->wcount / 4, (float)scale_w, 0);
  CMF_TRY(launch_check(h, "scale_w"));
  const long long n4 = h->RH * h->Kp / 4;
  ew::scale_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>((float4*)h->Ht, n4, (float)scale_h, 0);
  CMF_TRY(launch_check(h, "scale_h"));
  CMF_TRY(sync_ops_W(h));
  CMF_TRY(sync_ops_H(h, 0, h->RH));
  h->est_valid = false;
  h->wterms_valid = false;
  return 0;
}

int cmf_mu_halo_width(cmf_mu_t* h, int* n_cols, int* halo_ld) {
  CMF_ENTER(h);
  if (n_cols) *n_cols = h->h;
  if (halo_ld) *halo_ld = h->Kp;
  return 0;
}

int cmf_mu_halo_export(cmf_mu_t* h, float* left_edge, float* right_edge) {
  CMF_ENTER(h);
  CMF_CHECK(h->have_factors, "no factors set");
  const size_t bytes = (size_t)h->h * h->Kp * 4;
  if (bytes == 0) return 0;
  CMF_CHECK(h->Tloc >= h->h, "shard shorter than the halo");
  if (left_edge)
    CMF_CUDA(cudaMemcpyAsync(left_edge, h->Ht + (long long)h->h * h->Kp, bytes, cudaMemcpyDeviceToDevice, h->stream));
  if (right_edge)
    CMF_CUDA(cudaMemcpyAsync(right_edge, h->Ht + (long long)(h->h + h->Tloc - h->h) * h->Kp, bytes,
                             cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

int cmf_mu_halo_import(cmf_mu_t* h, const float* left_halo, const float* right_halo) {
  CMF_ENTER(h);
  const size_t bytes = (size_t)h->h * h->Kp * 4;
  if (bytes == 0) return 0;
  if (left_halo) CMF_CUDA(cudaMemcpyAsync(h->Ht, left_halo, bytes, cudaMemcpyDeviceToDevice, h->stream));
  else CMF_CUDA(cudaMemsetAsync(h->Ht, 0, bytes, h->stream));
  float* r = h->Ht + (long long)(h->h + h->Tloc) * h->Kp;
  if (right_halo) CMF_CUDA(cudaMemcpyAsync(r, right_halo, bytes, cudaMemcpyDeviceToDevice, h->stream));
  else CMF_CUDA(cudaMemsetAsync(r, 0, bytes, h->stream));
  CMF_TRY(sync_ops_H(h, 0, h->h));
  CMF_TRY(sync_ops_H(h, h->h + h->Tloc, h->h));
  h->est_valid = false;
  return 0;
}

int cmf_mu_recon(cmf_mu_t* h) { CMF_ENTER(h); return do_recon(h); }
int cmf_mu_recon_loss(cmf_mu_t* h) { CMF_ENTER(h); return do_recon(h, !(gram_w(h) && gram_h(h))); }
int cmf_mu_w_terms(cmf_mu_t* h) { CMF_ENTER(h); return do_w_terms(h); }

int cmf_mu_w_terms_buffer(cmf_mu_t* h, float** dev_ptr, long long* count) {
  CMF_ENTER(h);
  if (dev_ptr) *dev_ptr = h->numden;
  if (count) *count = h->wcount;
  return 0;
}

int cmf_mu_w_apply(cmf_mu_t* h) { CMF_ENTER(h); return do_w_apply(h); }

int cmf_mu_h_step(cmf_mu_t* h) {
  CMF_ENTER(h);
  CMF_TRY(do_h_terms(h));
  return do_h_apply(h);
}

int cmf_mu_needs_mid_recon(cmf_mu_t* h, int* needed) {
  CMF_CHECK(h != nullptr && needed != nullptr, "null argument");
  *needed = gram_h(h) ? 0 : 1;
  return 0;
}

int cmf_mu_resid_sumsq(cmf_mu_t* h, double* sumsq) {
  CMF_ENTER(h);
  CMF_CHECK(sumsq != nullptr, "null argument");
  CMF_CHECK(h->est_valid, "Residuals not initialized.");                // base.py:95-96
  if (!h->sumsq_current) CMF_TRY(do_recon(h, h->est_stored));           // (only after an iteration that failed half-way)
  CMF_CUDA(cudaMemcpyAsync(sumsq, h->d_sumsq, 8, cudaMemcpyDeviceToHost, h->stream));
  CMF_CUDA(cudaStreamSynchronize(h->stream));
  if (h->use_tc) CMF_TRY(tc::check(h->tcs, h->stream));
  return 0;
}

int cmf_mu_resid_sumsq_buffer(cmf_mu_t* h, double** dev_ptr) {
  CMF_CHECK(h != nullptr && dev_ptr != nullptr, "null argument");
  *dev_ptr = h->d_sumsq;
  return 0;
}

int cmf_mu_loss(cmf_mu_t* h, double* loss) {
  double s = 0.0;
  CMF_TRY(cmf_mu_resid_sumsq(h, &s));
  *loss = std::sqrt(s > 0.0 ? s : 0.0) / h->norm_x;
  h->last_loss = *loss;
  return 0;
}

int cmf_mu_set_loss_mode(cmf_mu_t* h, int mode) {
  CMF_CHECK(h != nullptr, "null solver handle");
  CMF_CHECK(mode >= 0 && mode <= 2, "loss mode must be 0 (auto), 1 (full precision) or 2 (from the W terms)");
  if (mode != h->loss_mode) h->graph_dirty = h->sgraph_dirty = true;
  h->loss_mode = mode;
  return 0;
}

// Full Gram route (neither MU step reads est): how the loss of the coming steps is formed.  The decision needs a
// loss the host has already seen; the step functions take it anew every kLossBatch iterations.
//
// (a) from the W terms, no reconstruction at all (ew::wterms_dot_kernel): ||X - est||^2 = ||X||^2 - 2 <W, num_W> +
//     <W, den_W> with the W terms of the UPDATED factors, which the next iteration's W step needs anyway - the
//     iteration becomes  W update (cached terms) -> H terms -> H update -> W terms -> loss.  The identity is exact;
//     what it costs is cancellation: the three terms carry the relative error e of the contractions (3xTF32 with
//     two-level accumulation: a truncation bias below 1e-6, measured), so the loss moves by about e / (2 loss^2)
//     relative.  It is used while loss^2 >= 0.1, i.e. e-fold amplification of at most 5 (an error below 1e-5, a tenth
//     of the parity bar), and re-decided every kLossBatch steps; W and H never see it.
// (b) 3xTF32 only, otherwise: may the loss-only reconstruction run one operand pass?  (tc::recon explains the bound)
constexpr int kLossBatch = 8;
static void decide_loss_mode(cmf_mu_s* h) {
  const bool full_gram = gram_w(h) && gram_h(h) && h->loss_mode == 0 && h->last_loss > 0.0;
  static const bool identity_enabled = [] { const char* e = getenv("CMF_LOSS_IDENTITY"); return !e || atoi(e) != 0; }();
  // auto: the fp32-grade mode only.  Plain TF32 contractions carry the rounding of their operands and a truncation
  // bias of ~6e-8 per MMA step of a chain (2e-4 on the W terms at config C), which the identity would amplify to
  // 3e-4 of the loss there and to several 1e-3 on small problems; mode 2 asks for it explicitly.
  const bool ident = gram_w(h) && gram_h(h) &&
                     (h->loss_mode == 2 || (full_gram && h->x3 && identity_enabled && h->last_loss * h->last_loss >= 0.1));
  const double kl = (double)h->K * (double)h->L;
  // The omitted cross passes perturb the loss through <resid, W_lo (*) H_hi + W_hi (*) H_lo> = <W_lo, gW> + <H_lo, gH>:
  // sums of L N K and K T rounding residuals with random signs against the current gradients, i.e. a relative
  // perturbation of order 2^-12 / sqrt(min(L N K, K T)) times |gradient| / loss^2 - plus their own energy,
  // ~1.9e-8 / (K L loss^2).  Both are far below 1e-6 for a million or more factor entries and K L loss^2 >= 0.2
  // (tests/test_parity_gpu.py::test_one_pass_loss measures it); small problems keep all three passes.
  const double n_w = (double)h->L * h->N * h->K, n_h = (double)h->K * (double)h->p.t_global;
  const bool fast = h->x3 && full_gram && n_w >= 1048576.0 && n_h >= 1048576.0 &&
                    kl * h->last_loss * h->last_loss >= 0.2;
  if (fast != (h->tcs.loss_fast != 0) || ident != h->loss_gram) {
    h->tcs.loss_fast = fast ? 1 : 0;
    h->loss_gram = ident;
    h->graph_dirty = h->sgraph_dirty = true;
  }
}
static bool loss_batched(const cmf_mu_s* h) { return gram_w(h) && gram_h(h) && h->x3 && h->loss_mode == 0; }

// the loss of the current factors from their W terms: d_sumsq = ||X_local||^2 + sum W (den_W - 2 num_W)
static int do_loss_identity(cmf_mu_s* h) {
  CMF_CHECK(h->wterms_valid, "loss identity before w_terms");
  const long long n4 = h->wcount / 4;
  const int grid = ew_grid(h, n4);
  ew::wterms_dot_kernel<<<grid, 256, 0, h->stream>>>((const float4*)h->W, (const float4*)h->numden,
                                                     (const float4*)(h->numden + h->wcount), n4, h->d_xpart);
  CMF_TRY(launch_check(h, "loss_identity"));
  ew::sum_doubles_offset_kernel<<<1, 1024, 0, h->stream>>>(h->d_xpart, grid, h->d_sumsq, h->sumsq_x);
  CMF_TRY(launch_check(h, "loss_sum"));
  h->est_valid = true;        // the loss of the current factors is known; est itself is not stored
  h->est_stored = false;
  h->sumsq_current = true;
  return 0;
}

// one MU iteration (reference MultUpdate.update, mult.py:15-25) issued on the solver's stream
static int issue_iteration(cmf_mu_s* h, bool prof, size_t& ne, int slot) {
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (!h->loss_gram) CMF_TRY(do_w_terms(h));         // (loss identity: cached by the previous iteration)
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  CMF_TRY(do_w_apply(h));
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (!gram_h(h)) CMF_TRY(do_recon(h, true, false));       // the Gram H step does not read est
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  CMF_TRY(do_h_terms(h));
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  CMF_TRY(do_h_apply(h));
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (h->loss_gram) {
    CMF_TRY(do_w_terms(h));                          // the W terms of the updated factors: loss now, W step next
    CMF_TRY(do_loss_identity(h));
  } else {
    CMF_TRY(do_recon(h, !(gram_w(h) && gram_h(h))));   // full Gram route: est is only needed for the loss
  }
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (slot >= 0) ew::loss_from_sumsq_kernel<<<1, 1, 0, h->stream>>>(h->d_sumsq, h->norm_x, h->d_ring, slot);
  else ew::loss_from_sumsq_counter_kernel<<<1, 1, 0, h->stream>>>(h->d_sumsq, h->norm_x, h->d_ring, h->d_counter);
  return launch_check(h, "loss");
}

// capture one iteration into h->graph_exec; on any failure fall back to plain launches for good
static void capture_iteration(cmf_mu_s* h) {
  if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
  h->graph_dirty = true;
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    h->graph_ok = false;
    return;
  }
  const long long l0 = h->launches;
  // issue_iteration walks the host-side state flags although nothing executes during capture: put them back, so
  // that a capture that fails half-way leaves the plain-launch fallback a consistent state
  const bool est_valid = h->est_valid, est_stored = h->est_stored, wterms_valid = h->wterms_valid;
  const bool sumsq_current = h->sumsq_current;
  size_t ne = 0;
  const int rc = issue_iteration(h, false, ne, -1);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  h->est_valid = est_valid; h->est_stored = est_stored; h->wterms_valid = wterms_valid;
  h->sumsq_current = sumsq_current;
  h->graph_launches = h->launches - l0;
  h->launches = l0;                       // nothing ran yet; replays add graph_launches each
  if (rc != 0 || e != cudaSuccess || graph == nullptr ||
      cudaGraphInstantiate(&h->graph_exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    h->graph_exec = nullptr;
    h->graph_ok = false;
  } else {
    h->graph_dirty = false;
  }
  if (graph) cudaGraphDestroy(graph);
}

int cmf_mu_step(cmf_mu_t* h, int n_steps, double* loss_out, float* ms_out) {
  CMF_ENTER(h);
  CMF_CHECK(n_steps >= 0, "n_steps must be >= 0");
  CMF_CHECK(h->have_data && h->have_factors, "step before data/factors were set");
  if (n_steps == 0) return 0;
  for (int k = 0; k < 4; ++k) h->kernel_ms[k] = 0.f;
  const bool prof = h->profiling != 0;
  static const bool graphs_enabled = [] { const char* e = getenv("CMF_GRAPH"); return !e || atoi(e) != 0; }();
  const int batch = loss_batched(h) && kLossBatch < h->ring_cap ? kLossBatch : h->ring_cap;
  int done = 0;
  while (done < n_steps) {
    const int chunk = (n_steps - done < batch) ? n_steps - done : batch;
    decide_loss_mode(h);
    const bool ident = h->loss_gram;
    if (ident) { if (!h->wterms_valid) CMF_TRY(do_w_terms(h)); }
    else if (!h->est_valid) CMF_TRY(do_recon(h));
    bool use_graph = graphs_enabled && !prof && h->graph_ok;
    if (use_graph && (h->graph_dirty || !h->graph_exec)) {
      capture_iteration(h);
      use_graph = h->graph_ok && h->graph_exec != nullptr;
      last_error().clear();
    }
    size_t ne = 0;
    if (use_graph) CMF_CUDA(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
    launch_log_begin(h);
    for (int i = 0; i < chunk; ++i) {
      if (ms_out) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
      if (use_graph) {
        CMF_CUDA(cudaGraphLaunch(h->graph_exec, h->stream));
        h->launches += h->graph_launches;
      } else {
        CMF_TRY(issue_iteration(h, prof, ne, i));
      }
    }
    if (use_graph) { h->est_valid = true; h->est_stored = !(gram_w(h) && gram_h(h)); h->wterms_valid = ident; h->sumsq_current = true; }
    if (ms_out) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
    if (loss_out)
      CMF_CUDA(cudaMemcpyAsync(loss_out + done, h->d_ring, (size_t)chunk * 8, cudaMemcpyDeviceToHost, h->stream));
    CMF_CUDA(cudaMemcpyAsync(&h->last_loss, h->d_ring + (chunk - 1), 8, cudaMemcpyDeviceToHost, h->stream));
    CMF_CUDA(cudaStreamSynchronize(h->stream));
    launch_log_end(h);
    if (h->use_tc) CMF_TRY(tc::check(h->tcs, h->stream));
    // unpack event timings
    const int per = (ms_out ? 1 : 0) + (prof ? 7 : 0);
    for (int i = 0; i < chunk; ++i) {
      const size_t b = (size_t)i * per;
      if (ms_out) {
        const size_t nxt = (i + 1 < chunk) ? (size_t)(i + 1) * per : ne - 1;
        CMF_CUDA(cudaEventElapsedTime(ms_out + done + i, h->ev_pool[b], h->ev_pool[nxt]));
      }
      if (prof) {
        const size_t e0 = b + (ms_out ? 1 : 0);
        float t[6];
        for (int k = 0; k < 6; ++k) CMF_CUDA(cudaEventElapsedTime(&t[k], h->ev_pool[e0 + k], h->ev_pool[e0 + k + 1]));
        h->kernel_ms[1] += ident ? t[5] : t[0];   // w_terms (loss identity: at the end of the iteration, + the loss)
        h->kernel_ms[3] += t[1] + t[4];           // elementwise updates
        h->kernel_ms[0] += ident ? 0.f : t[2] + t[5];   // reconstructions (+ loss)
        h->kernel_ms[2] += t[3];                  // h_terms
      }
    }
    done += chunk;
  }
  return 0;
}

int cmf_mu_get_W(cmf_mu_t* h, void* W_out, int dtype, int mem) {
  CMF_ENTER(h);
  CMF_CHECK(W_out != nullptr, "null argument");
  CMF_CHECK(h->have_factors, "W or H not initalized.");
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  const long long wn = (long long)h->L * h->N * h->K;
  const size_t es = dtype == CMF_F32 ? 4 : 8;
  void* dst = W_out;
  void* stage = nullptr;
  if (mem == CMF_HOST) { CMF_TRY(stage_get(h, (size_t)wn * es, &stage)); dst = stage; }
  const int grid = ew_grid(h, wn);
  if (dtype == CMF_F32) ew::w_pad_out_kernel<float><<<grid, 256, 0, h->stream>>>(h->W, (float*)dst, h->L, h->N, h->K, h->Np, h->Kp);
  else ew::w_pad_out_kernel<double><<<grid, 256, 0, h->stream>>>(h->W, (double*)dst, h->L, h->N, h->K, h->Np, h->Kp);
  int rc = launch_check(h, "w_pad_out");
  if (rc == 0 && mem == CMF_HOST &&
      cudaMemcpyAsync(W_out, stage, (size_t)wn * es, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { set_error("D2H copy of W failed"); rc = 1; }
  if (rc == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("sync failed: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; }
  return rc;
}

int cmf_mu_get_H(cmf_mu_t* h, void* H_out, int dtype, int mem, long long ldh) {
  CMF_ENTER(h);
  CMF_CHECK(H_out != nullptr, "null argument");
  CMF_CHECK(h->have_factors, "W or H not initalized.");
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(ldh >= h->Tloc, "leading dimension too small");
  const float* src = h->Ht + (long long)h->h * h->Kp;
  int rc;
  if (dtype == CMF_F32) rc = store_transposed<float>(h, src, h->Kp, h->K, h->Tloc, (float*)H_out, mem, ldh);
  else rc = store_transposed<double>(h, src, h->Kp, h->K, h->Tloc, (double*)H_out, mem, ldh);
  if (rc == 0) CMF_CUDA(cudaStreamSynchronize(h->stream));
  return rc;
}

int cmf_mu_get_est(cmf_mu_t* h, void* est_out, int dtype, int mem, long long ld) {
  CMF_ENTER(h);
  CMF_CHECK(est_out != nullptr, "null argument");
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(ld >= h->Tloc, "leading dimension too small");
  CMF_TRY(ensure_est_stored(h));
  int rc;
  if (dtype == CMF_F32) rc = store_transposed<float>(h, h->Et, h->Np, h->N, h->Tloc, (float*)est_out, mem, ld, h->Elo);
  else rc = store_transposed<double>(h, h->Et, h->Np, h->N, h->Tloc, (double*)est_out, mem, ld, h->Elo);
  if (rc == 0) CMF_CUDA(cudaStreamSynchronize(h->stream));
  return rc;
}

int cmf_mu_h_terms(cmf_mu_t* h, void* num_out, void* den_out, int dtype) {
  CMF_ENTER(h);
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  if (!h->est_valid && !gram_h(h)) CMF_TRY(do_recon(h));
  CMF_TRY(do_h_terms(h));
  for (int s = 0; s < 2; ++s) {
    void* out = s ? den_out : num_out;
    if (!out) continue;
    const float* src = h->hterms + (long long)s * h->TO * h->Kp;
    if (dtype == CMF_F32) CMF_TRY(store_transposed<float>(h, src, h->Kp, h->K, h->Tloc, (float*)out, CMF_HOST, h->Tloc));
    else CMF_TRY(store_transposed<double>(h, src, h->Kp, h->K, h->Tloc, (double*)out, CMF_HOST, h->Tloc));
  }
  CMF_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int cmf_mu_get_w_terms(cmf_mu_t* h, void* num_out, void* den_out, int dtype) {
  CMF_ENTER(h);
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(h->wterms_valid, "no W terms computed (call cmf_mu_w_terms)");
  const long long wn = (long long)h->L * h->N * h->K;
  const size_t es = dtype == CMF_F32 ? 4 : 8;
  void* stage = nullptr;
  CMF_TRY(stage_get(h, (size_t)wn * es, &stage));
  int rc = 0;
  for (int s = 0; s < 2 && rc == 0; ++s) {
    void* out = s ? den_out : num_out;
    if (!out) continue;
    const float* src = h->numden + (long long)s * h->wcount;
    const int grid = ew_grid(h, wn);
    if (dtype == CMF_F32) ew::w_pad_out_kernel<float><<<grid, 256, 0, h->stream>>>(src, (float*)stage, h->L, h->N, h->K, h->Np, h->Kp);
    else ew::w_pad_out_kernel<double><<<grid, 256, 0, h->stream>>>(src, (double*)stage, h->L, h->N, h->K, h->Np, h->Kp);
    rc = launch_check(h, "w_pad_out");
    if (rc == 0 && (cudaMemcpyAsync(out, stage, (size_t)wn * es, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                    cudaStreamSynchronize(h->stream) != cudaSuccess)) { set_error("D2H copy failed"); rc = 1; }
  }
  return rc;
}

int cmf_mu_launch_count(cmf_mu_t* h, long long* count) {
  CMF_CHECK(h != nullptr && count != nullptr, "null argument");
  *count = h->launches;
  return 0;
}

const char* cmf_mu_path_name(cmf_mu_t* h) {
  if (!h) return "none";
  if (!h->use_tc) return "ffma-fp32";
  if (h->x3) return h->tcs.gram == 3 ? "tcgen05-tf32x3+gram" : (h->tcs.gram ? "tcgen05-tf32x3+gram(partial)" : "tcgen05-tf32x3");
  if (h->tcs.mask != 7) return "tcgen05-tf32(partial)";
  if (h->tcs.gram == 3) return "tcgen05-tf32+gram";
  return h->tcs.gram ? "tcgen05-tf32+gram(partial)" : "tcgen05-tf32";
}

int cmf_mu_kernel_ms(cmf_mu_t* h, float out[4]) {
  CMF_CHECK(h != nullptr && out != nullptr, "null argument");
  for (int k = 0; k < 4; ++k) out[k] = h->kernel_ms[k];
  return 0;
}

int cmf_mu_set_profiling(cmf_mu_t* h, int on) {
  CMF_CHECK(h != nullptr, "null solver handle");
  h->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
  if (on >= 2) h->launch_table.clear();
  return 0;
}

// profiling level 2: one line "label launches total_ms" per kernel label seen since profiling was switched on
int cmf_mu_launch_table(cmf_mu_t* h, char* buf, long long cap) {
  CMF_CHECK(h != nullptr && buf != nullptr && cap > 0, "null argument");
  std::string out;
  char line[256];
  for (auto& row : h->launch_table) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", row.first.c_str(), row.second.first, row.second.second);
    out += line;
  }
  CMF_CHECK((long long)out.size() + 1 <= cap, "buffer too small for the launch table (%zu bytes needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

// ---- peer-memory collectives --------------------------------------------------
namespace {
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
int allocation_base(const void* ptr, unsigned long long* offset) {
  static GetAddressRangeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess) fn = (GetAddressRangeFn)f;
  }
  CMF_CHECK(fn != nullptr, "cuMemGetAddressRange is not available from this driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  CMF_CHECK(fn(&base, &size, (CUdeviceptr)ptr) == CUDA_SUCCESS, "cuMemGetAddressRange failed");
  *offset = (unsigned long long)((CUdeviceptr)ptr - base);
  return 0;
}
size_t peer_halo_bytes(const cmf_mu_s* h) { return (size_t)round_up_ll(2ll * h->h * h->Kp * 4 + 16, 256); }
}  // namespace

int cmf_mu_peer_export(cmf_mu_t* h, void* blob) {
  CMF_ENTER(h);
  CMF_CHECK(blob != nullptr, "null argument");
  auto& ps = h->peer;
  if (!ps.shared) {
    ps.shared_bytes = 256 + peer_halo_bytes(h) + (size_t)ps.ring_cap * peer::kMaxPeers * 8;
    CMF_CUDA(cudaMalloc(&ps.shared, ps.shared_bytes));
    CMF_CUDA(cudaMemset(ps.shared, 0, ps.shared_bytes));
  }
  static_assert(sizeof(peer::Control) <= 256, "control block larger than its slot");
  PeerBlob b{};
  b.magic = kPeerMagic; b.dev = h->dev; b.h = h->h; b.Kp = h->Kp; b.ring_cap = ps.ring_cap; b.wcount = h->wcount;
  const void* ptrs[3] = {h->numden, h->W, ps.shared};
  for (int k = 0; k < 3; ++k) {
    CMF_CUDA(cudaIpcGetMemHandle(&b.handle[k], const_cast<void*>(ptrs[k])));
    CMF_TRY(allocation_base(ptrs[k], &b.offset[k]));
  }
  memset(blob, 0, CMF_PEER_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int cmf_mu_peer_attach(cmf_mu_t* h, int rank, int world, const void* blobs) {
  CMF_ENTER(h);
  CMF_CHECK(blobs != nullptr, "null argument");
  CMF_CHECK(world >= 1 && world <= peer::kMaxPeers && rank >= 0 && rank < world, "rank %d / world %d out of range (at most %d peers)",
            rank, world, peer::kMaxPeers);
  auto& ps = h->peer;
  CMF_CHECK(ps.shared != nullptr, "cmf_mu_peer_export must be called first");
  CMF_CHECK(!ps.attached, "peers already attached");
  CMF_CHECK(h->wcount % 4 == 0, "W element count must be a multiple of 4");
  peer::Peers P{};
  P.rank = rank; P.world = world;
  for (int p = 0; p < world; ++p) {
    PeerBlob b;
    memcpy(&b, (const char*)blobs + (size_t)p * CMF_PEER_BLOB_BYTES, sizeof(b));
    CMF_CHECK(b.magic == kPeerMagic, "peer %d: not a cmf peer blob", p);
    CMF_CHECK(b.h == h->h && b.Kp == h->Kp && b.wcount == h->wcount && b.ring_cap == ps.ring_cap,
              "peer %d was created with different dimensions", p);
    char* base[3];
    if (p == rank) {
      base[0] = (char*)h->numden; base[1] = (char*)h->W; base[2] = (char*)ps.shared;
    } else {
      for (int k = 0; k < 3; ++k) {
        void* m = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&m, b.handle[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          set_error("cudaIpcOpenMemHandle failed for peer %d (device %d): %s", p, b.dev, cudaGetErrorString(e));
          cudaGetLastError();
          peer_detach(h);
          return 1;
        }
        ps.opened[p][k] = m;
        base[k] = (char*)m + b.offset[k];
      }
    }
    P.numden[p] = (float*)base[0];
    P.W[p] = (float*)base[1];
    P.ctl[p] = (peer::Control*)base[2];
    P.halo_in[p] = (float*)(base[2] + 256);
    P.ring[p] = (double*)(base[2] + 256 + peer_halo_bytes(h));
  }
  ps.P = P;
  ps.attached = true;
  return 0;
}

int cmf_mu_peer_detach(cmf_mu_t* h) {
  CMF_ENTER(h);
  return peer_detach(h);
}

namespace {
int peer_ensure_shared(cmf_mu_s* h) {
  auto& ps = h->peer;
  if (ps.shared) return 0;
  DeviceGuard guard(h->dev);
  CMF_CHECK(guard.ok, "cannot select CUDA device %d", h->dev);
  ps.shared_bytes = 256 + peer_halo_bytes(h) + (size_t)ps.ring_cap * peer::kMaxPeers * 8;
  CMF_CUDA(cudaMalloc(&ps.shared, ps.shared_bytes));
  CMF_CUDA(cudaMemset(ps.shared, 0, ps.shared_bytes));
  return 0;
}
}  // namespace

int cmf_mu_peer_attach_local(cmf_mu_t* h, int rank, int world, cmf_mu_t* const* handles) {
  CMF_ENTER(h);
  CMF_CHECK(handles != nullptr, "null argument");
  CMF_CHECK(world >= 1 && world <= peer::kMaxPeers && rank >= 0 && rank < world, "rank %d / world %d out of range (at most %d peers)",
            rank, world, peer::kMaxPeers);
  CMF_CHECK(handles[rank] == h, "handles[rank] must be this solver");
  auto& ps = h->peer;
  CMF_CHECK(!ps.attached, "peers already attached");
  CMF_CHECK(h->wcount % 4 == 0, "W element count must be a multiple of 4");
  peer::Peers P{};
  P.rank = rank; P.world = world;
  for (int p = 0; p < world; ++p) {
    cmf_mu_s* o = handles[p];
    CMF_CHECK(o != nullptr, "null peer handle %d", p);
    CMF_CHECK(o->h == h->h && o->Kp == h->Kp && o->wcount == h->wcount && o->peer.ring_cap == ps.ring_cap,
              "peer %d was created with different dimensions", p);
    CMF_TRY(peer_ensure_shared(o));
    if (p != rank) {
      CMF_CHECK(o->dev != h->dev, "peer %d shares device %d with rank %d: one GPU per shard", p, o->dev, rank);
      int can = 0;
      CMF_CUDA(cudaDeviceCanAccessPeer(&can, h->dev, o->dev));
      CMF_CHECK(can, "device %d cannot access device %d over NVLink / PCIe peer-to-peer", h->dev, o->dev);
      cudaError_t e = cudaDeviceEnablePeerAccess(o->dev, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      CMF_CHECK(e == cudaSuccess, "cudaDeviceEnablePeerAccess(%d) failed: %s", o->dev, cudaGetErrorString(e));
    }
    char* sh = (char*)o->peer.shared;
    P.numden[p] = o->numden;
    P.W[p] = o->W;
    P.ctl[p] = (peer::Control*)sh;
    P.halo_in[p] = (float*)(sh + 256);
    P.ring[p] = (double*)(sh + 256 + peer_halo_bytes(o));
  }
  ps.P = P;
  ps.attached = true;
  return 0;
}

int cmf_mu_halo_exchange_peer(cmf_mu_t* h) {
  CMF_ENTER(h);
  auto& ps = h->peer;
  CMF_CHECK(ps.attached, "cmf_mu_halo_exchange_peer needs attached peers");
  CMF_CHECK(h->have_factors, "W or H not initalized.");
  if (h->h > 0) {
    peer::tick_kernel<<<1, 1, 0, h->stream>>>(ps.P.ctl[ps.P.rank], peer::kTickHalo);
    CMF_TRY(launch_check(h, "peer_tick"));
    peer::halo_exchange_kernel<<<2, 256, 0, h->stream>>>(ps.P, h->Ht, h->h, h->Kp, h->Tloc);
    CMF_TRY(launch_check(h, "halo_exchange"));
    CMF_TRY(sync_ops_H(h, 0, h->h));
    CMF_TRY(sync_ops_H(h, h->h + h->Tloc, h->h));
  }
  h->est_valid = false;
  int perr = 0;
  CMF_CUDA(cudaMemcpyAsync(&perr, &ps.P.ctl[ps.P.rank]->err, 4, cudaMemcpyDeviceToHost, h->stream));
  CMF_CUDA(cudaStreamSynchronize(h->stream));
  if (perr != 0) { set_error("peer halo exchange timed out (error %d): a rank did not reach the exchange", perr); return 1; }
  return 0;
}

// one sharded MU iteration issued on the solver's stream (constant kernel arguments: epochs and the loss slot are
// device counters, see peer::Control)
static int issue_sharded_iteration(cmf_mu_s* h, bool prof, size_t& ne, int ex_grid) {
  auto& ps = h->peer;
  const long long n4 = h->wcount / 4;
  peer::Control* me = ps.P.ctl[ps.P.rank];
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (!h->loss_gram) CMF_TRY(do_w_terms(h));         // (loss identity: cached by the previous iteration)
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  // reduce-scatter of the partial W terms + W update + all-gather of W: one kernel
  CMF_CHECK(h->wterms_valid, "w exchange before w_terms");
  peer::tick_kernel<<<1, 1, 0, h->stream>>>(me, peer::kTickExchange | (h->h > 0 ? peer::kTickHalo : 0));
  CMF_TRY(launch_check(h, "peer_tick"));
  peer::wstep_exchange_kernel<<<ex_grid, 256, 0, h->stream>>>(ps.P, n4);
  CMF_TRY(launch_check(h, "wstep_exchange"));
  h->wterms_valid = false;
  h->est_valid = false;
  CMF_TRY(sync_ops_W(h));
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (!gram_h(h)) CMF_TRY(do_recon(h, true, false));
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  CMF_TRY(do_h_terms(h));
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  CMF_TRY(do_h_apply(h));
  if (h->h > 0) {
    peer::halo_exchange_kernel<<<2, 256, 0, h->stream>>>(ps.P, h->Ht, h->h, h->Kp, h->Tloc);
    CMF_TRY(launch_check(h, "halo_exchange"));
    CMF_TRY(sync_ops_H(h, 0, h->h));
    CMF_TRY(sync_ops_H(h, h->h + h->Tloc, h->h));
  }
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  if (h->loss_gram) {
    // the local W-term partials of the updated factors (with the fresh halos): the loss now (the local values sum
    // to the global squared residual), the W exchange of the next iteration after it
    CMF_TRY(do_w_terms(h));
    CMF_TRY(do_loss_identity(h));
  } else {
    CMF_TRY(do_recon(h, !(gram_w(h) && gram_h(h))));
  }
  if (prof) CMF_CUDA(cudaEventRecord(get_event(h, ne++), h->stream));
  peer::sumsq_push_kernel<<<1, 32, 0, h->stream>>>(ps.P, h->d_sumsq);
  return launch_check(h, "sumsq_push");
}

// capture one sharded iteration into h->sgraph_exec; on any failure fall back to plain launches for good
static void capture_sharded_iteration(cmf_mu_s* h, int ex_grid) {
  if (h->sgraph_exec) { cudaGraphExecDestroy(h->sgraph_exec); h->sgraph_exec = nullptr; }
  h->sgraph_dirty = true;
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    h->sgraph_ok = false;
    return;
  }
  const long long l0 = h->launches;
  const bool est_valid = h->est_valid, est_stored = h->est_stored, wterms_valid = h->wterms_valid;
  const bool sumsq_current = h->sumsq_current;
  size_t ne = 0;
  const int rc = issue_sharded_iteration(h, false, ne, ex_grid);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  h->est_valid = est_valid; h->est_stored = est_stored; h->wterms_valid = wterms_valid;
  h->sumsq_current = sumsq_current;
  h->sgraph_launches = h->launches - l0;
  h->launches = l0;
  if (rc != 0 || e != cudaSuccess || graph == nullptr ||
      cudaGraphInstantiate(&h->sgraph_exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    h->sgraph_exec = nullptr;
    h->sgraph_ok = false;
  } else {
    h->sgraph_dirty = false;
  }
  if (graph) cudaGraphDestroy(graph);
}

// n_steps x MultUpdate.update() on a time shard, collectives over peer memory; every rank calls it
// with the same n_steps.  loss_out[i] = GLOBAL loss after step i (identical on all ranks).
int cmf_mu_step_sharded(cmf_mu_t* h, int n_steps, double* loss_out) {
  CMF_ENTER(h);
  CMF_CHECK(n_steps >= 0, "n_steps must be >= 0");
  CMF_CHECK(h->have_data && h->have_factors, "step before data/factors were set");
  auto& ps = h->peer;
  CMF_CHECK(ps.attached, "cmf_mu_step_sharded needs cmf_mu_peer_attach");
  if (n_steps == 0) return 0;
  for (int k = 0; k < 4; ++k) h->kernel_ms[k] = 0.f;
  const bool prof = h->profiling != 0;
  const long long n4 = h->wcount / 4;
  const int G = ps.P.world;
  long long slice = ceil_div_ll(n4, G);
  int ex_grid = (int)ceil_div_ll(slice, 256);
  if (ex_grid > h->num_sms) ex_grid = h->num_sms;       // every block spins: all must be resident
  if (ex_grid < 1) ex_grid = 1;
  static const bool graphs_enabled = [] { const char* e = getenv("CMF_GRAPH"); return !e || atoi(e) != 0; }();
  const int batch = loss_batched(h) && kLossBatch < ps.ring_cap ? kLossBatch : ps.ring_cap;
  std::vector<double> ring((size_t)ps.ring_cap * peer::kMaxPeers);
  int done = 0;
  while (done < n_steps) {
    const int chunk = (n_steps - done < batch) ? n_steps - done : batch;
    decide_loss_mode(h);               // (last_loss is the GLOBAL loss: every rank decides alike)
    const bool ident = h->loss_gram;
    if (ident) { if (!h->wterms_valid) CMF_TRY(do_w_terms(h)); }
    else if (!h->est_valid) CMF_TRY(do_recon(h));
    bool use_graph = graphs_enabled && !prof && h->sgraph_ok;
    if (use_graph && (h->sgraph_dirty || !h->sgraph_exec)) {
      capture_sharded_iteration(h, ex_grid);
      use_graph = h->sgraph_ok && h->sgraph_exec != nullptr;
      last_error().clear();
    }
    size_t ne = 0;
    launch_log_begin(h);
    peer::tick_kernel<<<1, 1, 0, h->stream>>>(ps.P.ctl[ps.P.rank], peer::kTickSlotReset);
    CMF_TRY(launch_check(h, "peer_tick"));
    for (int i = 0; i < chunk; ++i) {
      if (use_graph) {
        CMF_CUDA(cudaGraphLaunch(h->sgraph_exec, h->stream));
        h->launches += h->sgraph_launches;
      } else {
        CMF_TRY(issue_sharded_iteration(h, prof, ne, ex_grid));
      }
    }
    if (use_graph) { h->est_valid = true; h->est_stored = !(gram_w(h) && gram_h(h)); h->wterms_valid = ident; h->sumsq_current = true; }
    peer::barrier_kernel<<<1, 32, 0, h->stream>>>(ps.P, ++ps.bar_epoch);
    CMF_TRY(launch_check(h, "peer_barrier"));
    CMF_CUDA(cudaMemcpyAsync(ring.data(), ps.P.ring[ps.P.rank], (size_t)chunk * peer::kMaxPeers * 8, cudaMemcpyDeviceToHost,
                             h->stream));
    int perr = 0;
    CMF_CUDA(cudaMemcpyAsync(&perr, &ps.P.ctl[ps.P.rank]->err, 4, cudaMemcpyDeviceToHost, h->stream));
    CMF_CUDA(cudaStreamSynchronize(h->stream));
    launch_log_end(h);
    if (h->use_tc) CMF_TRY(tc::check(h->tcs, h->stream));
    if (perr != 0) {      // a runtime failure, not an argument error
      set_error("peer exchange timed out (error %d): a rank did not reach the collective", perr);
      return 1;
    }
    for (int i = 0; i < chunk; ++i) {
      double ssum = 0.0;
      for (int p = 0; p < G; ++p) ssum += ring[(size_t)i * peer::kMaxPeers + p];
      h->last_loss = std::sqrt(ssum > 0.0 ? ssum : 0.0) / h->norm_x;
      if (loss_out) loss_out[done + i] = h->last_loss;
    }
    if (prof) {
      for (int i = 0; i < chunk; ++i) {
        const size_t e0 = (size_t)i * 7;
        float t[6];
        for (int k = 0; k < 6; ++k) CMF_CUDA(cudaEventElapsedTime(&t[k], h->ev_pool[e0 + k], h->ev_pool[e0 + k + 1]));
        h->kernel_ms[1] += ident ? t[5] : t[0];         // w_terms (loss identity: at the end, + the loss)
        h->kernel_ms[3] += t[1] + t[4];                 // exchange + updates, halo exchange
        h->kernel_ms[0] += ident ? 0.f : t[2] + t[5];   // reconstructions (+ loss)
        h->kernel_ms[2] += t[3];                        // h_terms
      }
    }
    done += chunk;
  }
  return 0;
}

// ---- gradient descent / block coordinate descent (reference algs/gradient_descent.py) ----------
namespace {
int gd_ensure(cmf_mu_s* h) {
  auto& g = h->gdst;
  if (g.ready) return 0;
  CMF_CHECK(!gram_w(h) && !gram_h(h), "the gradient solvers contract the residual directly: create the solver with CMF_DEN_DIRECT");
  const long long pcount = (long long)h->L * h->Kp * h->Kp;
  const long long n = (long long)h->L * h->Kp;
  if (!h->use_tc) {
    CMF_TRY(dmalloc(&g.P, pcount));
    if (h->wsplits > 1) CMF_TRY(dmalloc(&g.Ppart, pcount * h->wsplits));
  }
  CMF_TRY(dmalloc(&g.Pt, pcount));
  CMF_TRY(dmalloc(&g.v, n));
  CMF_TRY(dmalloc(&g.y, n));
  CMF_TRY(dmalloc(&g.d_inv, 1));
  CMF_TRY(dmalloc(&g.d_lam, 1));
  CMF_TRY(dmalloc(&g.st, 1));
  CMF_CUDA(cudaMemsetAsync(g.v, 0, (size_t)n * 4, h->stream));
  g.ready = true;
  return 0;
}

// lambda_max of the block-Toeplitz autocorrelation matrix of H (lipschitz_W, gradient_descent.py:54-69);
// lambda and 1 / lambda stay on the device (d_lam, d_inv)
int gd_lipschitz(cmf_mu_s* h) {
  CMF_TRY(gd_ensure(h));
  auto& g = h->gdst;
  const float* P = nullptr;
  if (h->use_tc) {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::autocorr(h->tcs, h->stream));
    h->launches += tc::launch_counter() - n0;
    P = h->tcs.P;
  } else {
    const int LKp = h->L * h->Kp;
    const long long pcount = (long long)h->L * h->Kp * h->Kp;
    gd::AutoA a{h->Ht + (long long)h->h * h->Kp, h->Kp};
    simt::WTermsB b{h->Ht, h->Kp, h->h, LKp};
    gd::AutoEpi e{h->wsplits == 1 ? g.P : g.Ppart, h->Kp, LKp, pcount};
    dim3 grid((unsigned)ceil_div_ll(h->Kp, 128), (unsigned)ceil_div_ll(LKp, 128), (unsigned)h->wsplits);
    simt::shift_gemm_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, h->stream>>>(a, b, e, h->Tloc, h->wchunk, 1);
    CMF_TRY(launch_check(h, "autocorr_H"));
    if (h->wsplits > 1) {
      ew::sum_splits_kernel<<<ew_grid(h, pcount / 4), 256, 0, h->stream>>>((float4*)g.P, (const float4*)g.Ppart, pcount / 4,
                                                                          pcount / 4, h->wsplits);
      CMF_TRY(launch_check(h, "autocorr_H_sum"));
    }
    P = g.P;
  }
  const int n = h->L * h->Kp;
  gd::transpose_blocks_kernel<<<ew_grid(h, (long long)n * h->Kp), 256, 0, h->stream>>>(P, g.Pt, h->L, h->Kp);
  CMF_TRY(launch_check(h, "autocorr_transpose"));
  gd::power_reset_kernel<<<1, 1024, 0, h->stream>>>(g.st, g.v, n);
  CMF_TRY(launch_check(h, "power_reset"));
  const int mv_blocks = (int)ceil_div_ll((long long)n * 32, 256);
  // A Rayleigh quotient is a LOWER bound of lambda_max (the W step 1 / lambda would be too long), so the iteration is
  // not cut off at a fixed count: after every batch the `done` flag comes back (8 bytes; the solver step is
  // synchronous anyway) and unsettled iterations go on, warm-started, for up to max_batches batches.  What is left
  // unsettled after that is reported through cmf_gd_lipschitz_state.
  g.iters = 0;
  g.settled = false;
  for (int b = 0; b < g.max_batches && !g.settled; ++b) {
    for (int it = 0; it < g.batch_iter; ++it) {
      gd::toeplitz_matvec_kernel<<<mv_blocks, 256, 0, h->stream>>>(P, g.Pt, g.v, g.y, h->L, h->Kp, g.st);
      gd::power_normalize_kernel<<<1, 1024, 0, h->stream>>>(g.v, g.y, n, g.st, 1e-7, g.d_lam, g.d_inv);
      h->launches += 2;
    }
    g.iters += g.batch_iter;
    cudaError_t e = cudaGetLastError();
    CMF_CHECK(e == cudaSuccess, "power iteration launch failed: %s", cudaGetErrorString(e));
    double done = 0.0;
    CMF_CUDA(cudaMemcpyAsync(&done, &g.st->done, 8, cudaMemcpyDeviceToHost, h->stream));
    CMF_CUDA(cudaStreamSynchronize(h->stream));
    g.settled = done != 0.0;
  }
  return 0;
}

int gd_projected_step_W(cmf_mu_s* h) {
  CMF_CHECK(h->wterms_valid, "W step before its gradient was cached");
  const long long n4 = h->wcount / 4;
  gd::projected_step_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>(
      (float4*)h->W, (const float4*)h->numden, (const float4*)(h->numden + h->wcount), n4, h->gdst.d_inv, 0.f,
      (float4*)tc::fused_w_op(h->tcs));
  CMF_TRY(launch_check(h, "gd_w_step"));
  {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::refresh_w(h->tcs, h->stream, true));
    h->launches += tc::launch_counter() - n0;
  }
  h->wterms_valid = false;
  h->est_valid = false;
  return 0;
}

int gd_projected_step_H(cmf_mu_s* h, double step) {
  const long long n4 = h->Tloc * h->Kp / 4;
  gd::projected_step_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>(
      (float4*)(h->Ht + (long long)h->h * h->Kp), (const float4*)h->hterms, (const float4*)(h->hterms + h->TO * h->Kp), n4,
      nullptr, (float)step,
      tc::fused_h_op(h->tcs) ? (float4*)(tc::fused_h_op(h->tcs) + (long long)h->h * h->Kp) : nullptr);
  CMF_TRY(launch_check(h, "gd_h_step"));
  {
    const long long n0 = tc::launch_counter();
    CMF_TRY(tc::refresh_h(h->tcs, h->stream, h->h, h->Tloc, true));
    h->launches += tc::launch_counter() - n0;
  }
  h->est_valid = false;
  return 0;
}
}  // namespace

// cache_resids + cache_gW + cache_gH (GradDescent.__init__, gradient_descent.py:36-38): afterwards the W-term and
// H-term buffers hold num / den of the CURRENT factors, i.e. gW = den_W - num_W and gH = den_H - num_H.
int cmf_gd_cache(cmf_mu_t* h) {
  CMF_ENTER(h);
  CMF_CHECK(h->have_data && h->have_factors, "gradient cache before data/factors were set");
  CMF_CHECK(h->t_valid == h->Tloc && h->p.t_local == h->p.t_global, "the gradient solvers run on one GPU (no time sharding)");
  CMF_TRY(gd_ensure(h));
  CMF_TRY(do_recon(h));
  CMF_TRY(do_w_terms(h));
  CMF_TRY(do_h_terms(h));
  h->gdst.cached = true;
  return 0;
}

// lipschitz_W (gradient_descent.py:54-69)
int cmf_gd_lipschitz_w(cmf_mu_t* h, double* lam) {
  CMF_ENTER(h);
  CMF_CHECK(lam != nullptr, "null argument");
  CMF_CHECK(h->have_factors, "W or H not initalized.");
  CMF_TRY(gd_lipschitz(h));
  CMF_CUDA(cudaMemcpyAsync(lam, h->gdst.d_lam, 8, cudaMemcpyDeviceToHost, h->stream));
  CMF_CUDA(cudaStreamSynchronize(h->stream));
  if (h->use_tc) CMF_TRY(tc::check(h->tcs, h->stream));
  return 0;
}

int cmf_gd_lipschitz_state(cmf_mu_t* h, int* settled, int* iterations) {
  CMF_ENTER(h);
  CMF_CHECK(settled != nullptr && iterations != nullptr, "null argument");
  *settled = h->gdst.settled ? 1 : 0;
  *iterations = h->gdst.iters;
  return 0;
}

// GradDescent.update (gradient_descent.py:81-92) when block_descent == 0, BlockDescent.update (:132-147) otherwise.
// step_size_h is the caller's current H step (the reference adapts it in converged(), :94-113).
int cmf_gd_step(cmf_mu_t* h, int block_descent, double step_size_h, double* loss_out) {
  CMF_ENTER(h);
  CMF_CHECK(h->gdst.cached, "cmf_gd_step before cmf_gd_cache");
  CMF_TRY(gd_lipschitz(h));
  CMF_TRY(gd_projected_step_W(h));
  if (!block_descent) {
    CMF_TRY(gd_projected_step_H(h, step_size_h));       // both steps use the gradients of the OLD factors
    CMF_TRY(do_recon(h));
    CMF_TRY(do_w_terms(h));
    CMF_TRY(do_h_terms(h));
  } else {
    CMF_TRY(do_recon(h, true, false));                   // est with the new W (its loss is not read)
    CMF_TRY(do_h_terms(h));
    CMF_TRY(gd_projected_step_H(h, step_size_h));
    CMF_TRY(do_recon(h));
    CMF_TRY(do_w_terms(h));                              // gW for the next iteration; gH is recomputed there
  }
  if (loss_out) CMF_TRY(cmf_mu_loss(h, loss_out));
  else CMF_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- HALS (reference algs/hals.py + algs/accelerated.py) -------------------------------------------
namespace {
int hals_ensure(cmf_mu_s* h) {
  auto& s = h->halsst;
  if (s.ready) return 0;
  CMF_CHECK(h->p.t_local == h->p.t_global, "HALS runs on one GPU (no time sharding)");
  CMF_CHECK(h->Tloc > h->L, "HALS needs more time points than lags (T=%lld, L=%d)", h->Tloc, h->L);
  const long long bx = ceil_div_ll(h->Np, 256);
  long long nc = ceil_div_ll(4ll * h->num_sms, bx);
  const long long max_nc = ceil_div_ll(h->Tloc, 64);
  if (nc > max_nc) nc = max_nc;
  if (nc < 1) nc = 1;
  s.rows_per_chunk = (int)ceil_div_ll(h->Tloc, nc);
  s.n_chunks = (int)ceil_div_ll(h->Tloc, s.rows_per_chunk);
  const long long big = (h->wcount > h->RH * h->Kp) ? h->wcount : h->RH * h->Kp;
  CMF_TRY(dmalloc(&s.Rt, h->RT * h->Np));
  CMF_TRY(dmalloc(&s.part, (long long)s.n_chunks * h->Np));
  CMF_TRY(dmalloc(&s.part_h, s.n_chunks));
  CMF_TRY(dmalloc(&s.delta, h->Np));
  CMF_TRY(dmalloc(&s.Wk, (long long)h->L * h->Np));
  CMF_TRY(dmalloc(&s.prev, big));
  CMF_TRY(dmalloc(&s.w2, h->L));
  CMF_TRY(dmalloc(&s.d_diff, 1));
  s.ready = true;
  return 0;
}
}  // namespace

// setup: the residual est - X of the current factors (cache_resids, base.py:57-62), on the device
int cmf_hals_begin(cmf_mu_t* h) {
  CMF_ENTER(h);
  CMF_CHECK(h->have_data && h->have_factors, "HALS before data/factors were set");
  CMF_TRY(hals_ensure(h));
  CMF_TRY(ensure_est_stored(h));
  const long long n4 = h->RT * h->Np / 4;
  hals::resid_kernel<<<ew_grid(h, n4), 256, 0, h->stream>>>((float4*)h->halsst.Rt, (const float4*)h->Et, (const float4*)h->Xt,
                                                            (const float4*)h->Elo, (const float4*)h->Xlo, n4);
  CMF_TRY(launch_check(h, "hals_resid"));
  h->halsst.open = true;
  return 0;
}

// one sweep over all (k, l) blocks of W (update_W, hals.py:47-48, 78-105); diff_norm = ||W_new - W_old||_F
// (the quantity the inner-iteration stop rule compares, accelerated.py:57-69; may be NULL)
int cmf_hals_sweep_w(cmf_mu_t* h, double* diff_norm) {
  CMF_ENTER(h);
  auto& s = h->halsst;
  CMF_CHECK(s.open, "cmf_hals_sweep_w outside cmf_hals_begin / cmf_hals_end");
  if (diff_norm) CMF_CUDA(cudaMemcpyAsync(s.prev, h->W, (size_t)h->wcount * 4, cudaMemcpyDeviceToDevice, h->stream));
  const dim3 grid((unsigned)ceil_div_ll(h->Np, 256), (unsigned)s.n_chunks);
  const float* hprev = nullptr;
  for (int k = 0; k < h->K; ++k)
    for (int l = 0; l < h->L; ++l) {
      const float* hcur = h->Ht + (long long)(h->h - l) * h->Kp + k;      // h(t) = H^T[h + t - l][k]
      hals::w_pass_kernel<<<grid, 256, 0, h->stream>>>(s.Rt, h->Np, h->Tloc, s.rows_per_chunk, hprev, s.delta, hcur, h->Kp,
                                                       s.part, s.part_h);
      hals::w_solve_kernel<<<grid.x, 256, 0, h->stream>>>(h->W + (long long)l * h->Np * h->Kp + k, h->Np, h->Kp, s.part,
                                                          s.part_h, s.n_chunks, s.delta);
      h->launches += 2;
      hprev = hcur;
    }
  hals::w_pass_kernel<<<grid, 256, 0, h->stream>>>(s.Rt, h->Np, h->Tloc, s.rows_per_chunk, hprev, s.delta, nullptr, h->Kp,
                                                   s.part, s.part_h);
  CMF_TRY(launch_check(h, "hals_w_sweep"));
  h->est_valid = false;
  h->wterms_valid = false;
  if (diff_norm) {
    hals::diff_sumsq_kernel<<<1, 1024, 0, h->stream>>>(h->W, s.prev, h->wcount, s.d_diff);
    CMF_TRY(launch_check(h, "hals_diff"));
    double ss = 0.0;
    CMF_CUDA(cudaMemcpyAsync(&ss, s.d_diff, 8, cudaMemcpyDeviceToHost, h->stream));
    CMF_CUDA(cudaStreamSynchronize(h->stream));
    *diff_norm = std::sqrt(ss);
  }
  return 0;
}

// one sweep over H (update_H, hals.py:68-70, 113-181): per component and lag the batch t = l (mod L), t < T - L,
// then the entry t = T - L + l with the motif cut by the end of the data
int cmf_hals_sweep_h(cmf_mu_t* h, double* diff_norm) {
  CMF_ENTER(h);
  auto& s = h->halsst;
  CMF_CHECK(s.open, "cmf_hals_sweep_h outside cmf_hals_begin / cmf_hals_end");
  const long long hcount = h->RH * h->Kp;
  if (diff_norm) CMF_CUDA(cudaMemcpyAsync(s.prev, h->Ht, (size_t)hcount * 4, cudaMemcpyDeviceToDevice, h->stream));
  const long long T = h->Tloc;
  const int L = h->L;
  for (int k = 0; k < h->K; ++k) {
    hals::gather_component_kernel<<<L, 256, 0, h->stream>>>(h->W, h->Np, h->Kp, k, s.Wk, s.w2);
    h->launches++;
    float* Hcol = h->Ht + (long long)h->h * h->Kp + k;
    for (int l = 0; l < L; ++l) {
      const long long span = T - L - l;                          // batch entries: range(l, T - L, L)
      const long long n_batch = span > 0 ? ceil_div_ll(span, L) : 0;
      if (n_batch > 0) {
        hals::h_entries_kernel<<<(unsigned)n_batch, 256, 0, h->stream>>>(s.Rt, h->Np, s.Wk, s.w2, L, Hcol, h->Kp, l, L);
        h->launches++;
      }
      hals::h_entries_kernel<<<1, 256, 0, h->stream>>>(s.Rt, h->Np, s.Wk, s.w2, L - l, Hcol, h->Kp, T - L + l, 0);
      h->launches++;
    }
  }
  cudaError_t e = cudaGetLastError();
  CMF_CHECK(e == cudaSuccess, "HALS H sweep launch failed: %s", cudaGetErrorString(e));
  h->est_valid = false;
  if (diff_norm) {
    hals::diff_sumsq_kernel<<<1, 1024, 0, h->stream>>>(h->Ht, s.prev, hcount, s.d_diff);
    CMF_TRY(launch_check(h, "hals_diff"));
    double ss = 0.0;
    CMF_CUDA(cudaMemcpyAsync(&ss, s.d_diff, 8, cudaMemcpyDeviceToHost, h->stream));
    CMF_CUDA(cudaStreamSynchronize(h->stream));
    *diff_norm = std::sqrt(ss);
  }
  return 0;
}

// end of update(): cache_resids from scratch and the loss (accelerated.py:83-84)
int cmf_hals_end(cmf_mu_t* h, double* loss_out) {
  CMF_ENTER(h);
  CMF_CHECK(h->halsst.open, "cmf_hals_end without cmf_hals_begin");
  h->halsst.open = false;
  CMF_TRY(sync_ops_W(h));
  CMF_TRY(sync_ops_H(h, 0, h->RH));
  CMF_TRY(do_recon(h));
  if (loss_out) CMF_TRY(cmf_mu_loss(h, loss_out));
  return 0;
}

// ---- stateless primitives ------------------------------------------------
static int make_tmp(cmf_mu_t** h, int N, long long T, int K, int L, int device, int precision) {
  cmf_mu_params p{};
  p.n_features = N; p.n_components = K; p.maxlag = L;
  p.t_local = T; p.t_global = T; p.t_offset = 0;
  p.device = device; p.precision = precision; p.stream = nullptr;
  return cmf_mu_create(h, &p);
}

int cmf_predict(const void* W, const void* H, void* est_out, int dtype, int n_features,
                long long n_timepoints, int n_components, int maxlag, int device, int precision) {
  cmf_mu_t* h = nullptr;
  CMF_TRY(make_tmp(&h, n_features, n_timepoints, n_components, maxlag, device, precision));
  h->have_data = true;   // X stays zero; only the reconstruction is wanted
  int rc = cmf_mu_set_factors(h, W, H, dtype, CMF_HOST, n_timepoints);
  if (rc == 0) rc = cmf_mu_get_est(h, est_out, dtype, CMF_HOST, n_timepoints);
  cmf_mu_destroy(h);
  return rc;
}

int cmf_tensor_transconv(const void* W, const void* X, void* out, int dtype, int n_features,
                         long long n_timepoints, int n_components, int maxlag, int device, int precision) {
  cmf_mu_t* h = nullptr;
  CMF_TRY(make_tmp(&h, n_features, n_timepoints, n_components, maxlag, device, precision));
  int rc = cmf_mu_set_data(h, X, dtype, CMF_HOST, n_timepoints, n_timepoints);
  // any H works: only the numerator (the transposed convolution of X) is read back
  std::vector<double> zeros;
  if (rc == 0) {
    zeros.assign((size_t)n_components * (size_t)n_timepoints, 0.0);
    rc = cmf_mu_set_factors(h, W, zeros.data(), dtype, CMF_HOST, n_timepoints);
  }
  if (rc == 0) rc = cmf_mu_h_terms(h, out, nullptr, dtype);
  cmf_mu_destroy(h);
  return rc;
}

// CMF.score (reference model.py:202-221): 1 - ||cmf_predict(W, H) - X||^2 / ||X||^2, formed on the device by the
// fused reconstruction + residual kernel: neither est nor the residual (N x T each) travels back to the host.
int cmf_score(const void* W, const void* H, const void* X, int dtype, int n_features, long long n_timepoints,
              int n_components, int maxlag, int device, int precision, double* r2_out) {
  CMF_CHECK(r2_out != nullptr, "null argument");
  cmf_mu_t* h = nullptr;
  CMF_TRY(make_tmp(&h, n_features, n_timepoints, n_components, maxlag, device, precision));
  int rc = cmf_mu_set_data(h, X, dtype, CMF_HOST, n_timepoints, n_timepoints);
  if (rc == 0) rc = cmf_mu_set_factors(h, W, H, dtype, CMF_HOST, n_timepoints);
  if (rc == 0) rc = cmf_mu_recon_loss(h);
  double ss = 0.0;
  if (rc == 0) rc = cmf_mu_resid_sumsq(h, &ss);
  if (rc == 0) *r2_out = 1.0 - ss / h->sumsq_x;
  cmf_mu_destroy(h);
  return rc;
}


// ==========================================================================
// The step before the solver: device matrices, the synthetic data set, the spectrogram
// ==========================================================================
}  // extern "C"

struct DevBuf {
  float* p = nullptr;
  int dev = 0;
  ~DevBuf() {
    if (p) { DeviceGuard g(dev); cached_free(p); }
  }
};
struct cmf_dmat_s {
  std::shared_ptr<DevBuf> buf;
  float* ptr = nullptr;
  long long rows = 0, cols = 0, ld = 0;
  int dev = 0;
};
struct cmf_synth_s {
  cmf_synth_params p;
  long long lead = 0, Tp = 0;          // H and the reconstruction start `lead` = min(L-1, t_offset) columns early
  int num_sms = 148;
  float *W = nullptr, *H = nullptr;    // L x N x K; K x Tp
  cmf_dmat_s data;                     // N x t_local view (ld = Tp) of the N x Tp reconstruction buffer
};

namespace {

int ds_grid(int num_sms, long long n_items) {
  long long blocks = ceil_div_ll(n_items, 256);
  const long long cap = (long long)num_sms * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

int new_dmat(cmf_dmat_s* m, long long rows, long long cols, int dev) {
  auto buf = std::make_shared<DevBuf>();
  buf->dev = dev;
  CMF_CUDA(cudaMalloc((void**)&buf->p, (size_t)(rows * cols > 0 ? rows * cols : 1) * 4));
  m->buf = buf; m->ptr = buf->p; m->rows = rows; m->cols = cols; m->ld = cols; m->dev = dev;
  return 0;
}

// rows x cols fp32 on the device (ld lds) -> host array of `dtype` (ld ldd)
int dmat_to_host(const float* src, long long lds, long long rows, long long cols, void* out, int dtype, long long ldd,
                 int num_sms) {
  CMF_CHECK(out != nullptr, "null argument");
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(ldd >= cols, "leading dimension %lld < %lld columns", ldd, cols);
  if (rows == 0 || cols == 0) return 0;
  if (dtype == CMF_F32) {
    CMF_CUDA(cudaMemcpy2D(out, (size_t)ldd * 4, src, (size_t)lds * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToHost));
    return 0;
  }
  long long rch = (32ll << 20) / cols;                     // 256 MB of doubles at a time
  if (rch < 1) rch = 1;
  if (rch > rows) rch = rows;
  double* tmp = nullptr;
  CMF_CUDA(cudaMalloc((void**)&tmp, (size_t)rch * cols * 8));
  int rc = 0;
  for (long long r0 = 0; r0 < rows && rc == 0; r0 += rch) {
    const long long nr = rows - r0 < rch ? rows - r0 : rch;
    ds::copy_convert_kernel<double><<<ds_grid(num_sms, nr * cols), 256>>>(tmp, cols, src + r0 * lds, lds, nr, cols);
    if (cudaMemcpy2D((double*)out + r0 * ldd, (size_t)ldd * 8, tmp, (size_t)cols * 8, (size_t)cols * 8, (size_t)nr,
                     cudaMemcpyDeviceToHost) != cudaSuccess) {
      set_error("device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
    }
  }
  cached_free(tmp);
  return rc;
}

int synth_noise(const cmf_synth_s* s, float* dst, const float* base, long long ld, float times) {
  const cmf_synth_params& p = s->p;
  ds::synth_noise_kernel<<<ds_grid(s->num_sms, (long long)p.n_features * p.t_local), 256>>>(
      dst, base, p.n_features, ld, p.t_local, p.t_offset, p.n_timebins, ds::stream_key(p.seed, ds::kStreamNoise),
      (float)p.noise_scale, times);
  CMF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

int cmf_dmat_info(cmf_dmat_t* m, const float** dev_ptr, long long* rows, long long* cols, long long* ld, int* device) {
  CMF_CHECK(m != nullptr, "null matrix handle");
  if (dev_ptr) *dev_ptr = m->ptr;
  if (rows) *rows = m->rows;
  if (cols) *cols = m->cols;
  if (ld) *ld = m->ld;
  if (device) *device = m->dev;
  return 0;
}

int cmf_dmat_get(cmf_dmat_t* m, void* out, int dtype, long long ld) {
  CMF_CHECK(m != nullptr, "null matrix handle");
  DeviceGuard guard(m->dev);
  CMF_CHECK(guard.ok, "cannot select CUDA device %d", m->dev);
  cudaDeviceProp prop;
  CMF_CUDA(cudaGetDeviceProperties(&prop, m->dev));
  return dmat_to_host(m->ptr, m->ld, m->rows, m->cols, out, dtype, ld, prop.multiProcessorCount);
}

int cmf_dmat_destroy(cmf_dmat_t* m) {
  delete m;
  return 0;
}

int cmf_synth_destroy(cmf_synth_t* s) {
  if (!s) return 0;
  {
    DeviceGuard guard(s->p.device);
    cached_free(s->W);
    cached_free(s->H);
  }
  delete s;
  return 0;
}

int cmf_synth_create(cmf_synth_t** out, const cmf_synth_params* p) {
  CMF_CHECK(out != nullptr && p != nullptr, "null argument");
  *out = nullptr;
  CMF_CHECK(p->n_components >= 1 && p->n_features >= 1 && p->n_lags >= 1 && p->n_timebins >= 1,
            "dimensions must be positive");
  CMF_CHECK(p->t_local >= 1 && p->t_offset >= 0 && p->t_offset + p->t_local <= p->n_timebins,
            "inconsistent time range: t_local=%lld t_offset=%lld n_timebins=%lld", p->t_local, p->t_offset, p->n_timebins);
  CMF_CHECK(p->H_sparsity >= 0.0 && p->H_sparsity <= 1.0, "H_sparsity must lie in [0, 1]");
  CMF_CHECK(cmf_precision_supported(p->precision, p->n_features, p->n_components, p->n_lags),
            "precision %d has no kernel for N=%d K=%d L=%d; use fp32", p->precision, p->n_features, p->n_components, p->n_lags);
  int ndev = 0;
  CMF_CUDA(cudaGetDeviceCount(&ndev));
  CMF_CHECK(p->device >= 0 && p->device < ndev, "device %d out of range (%d visible)", p->device, ndev);
  DeviceGuard guard(p->device);
  CMF_CHECK(guard.ok, "cannot select CUDA device %d", p->device);
  cudaDeviceProp prop;
  CMF_CUDA(cudaGetDeviceProperties(&prop, p->device));

  cmf_synth_s* s = new cmf_synth_s();
  s->p = *p;
  s->num_sms = prop.multiProcessorCount;
  const int N = p->n_features, K = p->n_components, L = p->n_lags;
  s->lead = p->t_offset < L - 1 ? p->t_offset : L - 1;
  s->Tp = p->t_local + s->lead;
  cmf_mu_t* h = nullptr;
  auto body = [&]() -> int {
    CMF_CUDA(cudaMalloc((void**)&s->W, (size_t)L * N * K * 4));
    CMF_CUDA(cudaMalloc((void**)&s->H, (size_t)K * s->Tp * 4));
    cmf_dmat_s buf;
    CMF_TRY(new_dmat(&buf, N, s->Tp, p->device));
    ds::synth_h_kernel<<<ds_grid(s->num_sms, (long long)K * s->Tp), 256>>>(
        s->H, K, s->Tp, s->Tp, p->t_offset - s->lead, p->n_timebins, ds::stream_key(p->seed, ds::kStreamH),
        (float)(1.0 - p->H_sparsity));
    CMF_CUDA(cudaGetLastError());
    ds::synth_w_kernel<<<(unsigned)ceil_div_ll(N, 128), 128>>>(s->W, L, N, K, ds::stream_key(p->seed, ds::kStreamMotif));
    CMF_CUDA(cudaGetLastError());
    CMF_CUDA(cudaDeviceSynchronize());
    // data = cmf_predict(W, H) (+ noise below): the K1 kernel of a throw-away solver over the Tp columns
    CMF_TRY(make_tmp(&h, N, s->Tp, K, L, p->device, p->precision));
    h->have_data = true;            // X stays zero; only the reconstruction is wanted
    CMF_TRY(cmf_mu_set_factors(h, s->W, s->H, CMF_F32, CMF_DEVICE, s->Tp));
    CMF_TRY(cmf_mu_get_est(h, buf.ptr, CMF_F32, CMF_DEVICE, s->Tp));
    s->data = buf;
    s->data.ptr = buf.ptr + s->lead;
    s->data.cols = p->t_local;
    CMF_TRY(synth_noise(s, s->data.ptr, s->data.ptr, s->Tp, 1.f));
    CMF_CUDA(cudaDeviceSynchronize());
    return 0;
  };
  const int rc = body();
  if (h) cmf_mu_destroy(h);
  if (rc != 0) { cmf_synth_destroy(s); return rc; }
  *out = s;
  return 0;
}

int cmf_synth_get(cmf_synth_t* s, int what, void* out, int dtype, long long ld) {
  CMF_CHECK(s != nullptr, "null generator handle");
  DeviceGuard guard(s->p.device);
  CMF_CHECK(guard.ok, "cannot select CUDA device %d", s->p.device);
  const cmf_synth_params& p = s->p;
  switch (what) {
    case CMF_SYNTH_W: {
      const long long cols = (long long)p.n_features * p.n_components;
      return dmat_to_host(s->W, cols, p.n_lags, cols, out, dtype, cols, s->num_sms);
    }
    case CMF_SYNTH_H:
      return dmat_to_host(s->H + s->lead, s->Tp, p.n_components, p.t_local, out, dtype, ld, s->num_sms);
    case CMF_SYNTH_DATA:
      return dmat_to_host(s->data.ptr, s->data.ld, p.n_features, p.t_local, out, dtype, ld, s->num_sms);
    case CMF_SYNTH_NOISE:
    case CMF_SYNTH_GENERATE: {
      cmf_dmat_s tmp;
      CMF_TRY(new_dmat(&tmp, p.n_features, p.t_local, p.device));
      if (what == CMF_SYNTH_NOISE) {
        CMF_TRY(synth_noise(s, tmp.ptr, nullptr, tmp.ld, 1.f));
      } else {
        CMF_CUDA(cudaMemcpy2D(tmp.ptr, (size_t)tmp.ld * 4, s->data.ptr, (size_t)s->data.ld * 4, (size_t)p.t_local * 4,
                              (size_t)p.n_features, cudaMemcpyDeviceToDevice));
        CMF_TRY(synth_noise(s, tmp.ptr, tmp.ptr, tmp.ld, 1.f));
      }
      return dmat_to_host(tmp.ptr, tmp.ld, p.n_features, p.t_local, out, dtype, ld, s->num_sms);
    }
    default:
      CMF_CHECK(false, "unknown item %d", what);
  }
  return 0;
}

int cmf_synth_matrix(cmf_synth_t* s, int what, cmf_dmat_t** out) {
  CMF_CHECK(s != nullptr && out != nullptr, "null argument");
  CMF_CHECK(what == CMF_SYNTH_DATA || what == CMF_SYNTH_GENERATE, "only DATA and GENERATE exist as device matrices");
  DeviceGuard guard(s->p.device);
  CMF_CHECK(guard.ok, "cannot select CUDA device %d", s->p.device);
  cmf_dmat_s* m = new cmf_dmat_s();
  if (what == CMF_SYNTH_DATA) {
    *m = s->data;                                   // shares the buffer
  } else {
    const cmf_synth_params& p = s->p;
    int rc = new_dmat(m, p.n_features, p.t_local, p.device);
    if (rc == 0 && cudaMemcpy2D(m->ptr, (size_t)m->ld * 4, s->data.ptr, (size_t)s->data.ld * 4, (size_t)p.t_local * 4,
                                (size_t)p.n_features, cudaMemcpyDeviceToDevice) != cudaSuccess) {
      set_error("device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = 1;
    }
    if (rc == 0) rc = synth_noise(s, m->ptr, m->ptr, m->ld, 1.f);
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) { set_error("generate failed"); rc = 1; }
    if (rc != 0) { delete m; return rc; }
  }
  *out = m;
  return 0;
}

int cmf_spectrogram(const void* audio, int dtype, int mem, long long n_samples, double fs, int nperseg, int noverlap,
                    const double* window, int normalize, int device, cmf_dmat_t** out) {
  CMF_CHECK(audio != nullptr && window != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  CMF_CHECK(dtype == CMF_F32 || dtype == CMF_F64, "unknown dtype %d", dtype);
  CMF_CHECK(mem == CMF_HOST || (mem == CMF_DEVICE && dtype == CMF_F32), "device audio must be float32");
  CMF_CHECK(nperseg >= 2 && noverlap >= 0 && noverlap < nperseg, "need 0 <= noverlap < nperseg, nperseg >= 2");
  CMF_CHECK(n_samples >= nperseg, "fewer samples (%lld) than one segment (%d)", n_samples, nperseg);
  CMF_CHECK(fs > 0.0, "sampling rate must be positive");
  int ndev = 0;
  CMF_CUDA(cudaGetDeviceCount(&ndev));
  CMF_CHECK(device >= 0 && device < ndev, "device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  CMF_CHECK(guard.ok, "cannot select CUDA device %d", device);
  cudaDeviceProp prop;
  CMF_CUDA(cudaGetDeviceProperties(&prop, device));
  const int hop = nperseg - noverlap, n_freq = nperseg / 2 + 1;
  const long long n_seg = (n_samples - noverlap) / hop;
  const int segld = nperseg | 1;                   // odd row stride: the 32 lanes of a warp read 32 segments
  const size_t smem = ((size_t)2 * nperseg + (size_t)ds::kSegs * segld + (size_t)n_freq * (ds::kSegs + 1)) * 4;
  CMF_CHECK(smem <= 227 * 1024, "nperseg=%d is too long for the on-chip DFT (limit about 1400)", nperseg);

  std::vector<float> w((size_t)nperseg);
  double w2 = 0.0;
  for (int i = 0; i < nperseg; ++i) { w[i] = (float)window[i]; w2 += window[i] * window[i]; }
  CMF_CHECK(w2 > 0.0, "the window is identically zero");
  float *d_audio = nullptr, *d_win = nullptr;
  double* d_part = nullptr;
  cmf_dmat_s* m = new cmf_dmat_s();
  auto body = [&]() -> int {
    const float* a = (const float*)audio;
    if (mem == CMF_HOST) {
      CMF_CUDA(cudaMalloc((void**)&d_audio, (size_t)n_samples * 4));
      if (dtype == CMF_F32) {
        CMF_CUDA(cudaMemcpy(d_audio, audio, (size_t)n_samples * 4, cudaMemcpyHostToDevice));
      } else {
        std::vector<float> tmp((size_t)n_samples);
        const double* src = (const double*)audio;
        for (long long i = 0; i < n_samples; ++i) tmp[(size_t)i] = (float)src[i];
        CMF_CUDA(cudaMemcpy(d_audio, tmp.data(), (size_t)n_samples * 4, cudaMemcpyHostToDevice));
      }
      a = d_audio;
    }
    CMF_CUDA(cudaMalloc((void**)&d_win, (size_t)nperseg * 4));
    CMF_CUDA(cudaMemcpy(d_win, w.data(), (size_t)nperseg * 4, cudaMemcpyHostToDevice));
    CMF_TRY(new_dmat(m, n_freq, n_seg, device));
    CMF_CUDA(cudaFuncSetAttribute(ds::spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ds::spectrogram_kernel<<<(unsigned)ceil_div_ll(n_seg, ds::kSegs), 256, smem>>>(
        a, n_samples, d_win, nperseg, segld, hop, n_seg, n_freq, (float)(1.0 / (fs * w2)), m->ptr, m->ld);
    CMF_CUDA(cudaGetLastError());
    if (normalize) {
      const long long chunk = 1 << 16;
      const int nchunk = (int)ceil_div_ll(n_seg, chunk);
      CMF_CUDA(cudaMalloc((void**)&d_part, (size_t)n_freq * nchunk * 2 * 8));
      ds::row_moments_kernel<<<dim3((unsigned)nchunk, (unsigned)n_freq), 256>>>(m->ptr, m->ld, n_seg, chunk, nchunk, d_part);
      CMF_CUDA(cudaGetLastError());
      long long bx = ceil_div_ll(n_seg, 256 * 8);
      if (bx > 64) bx = 64;
      ds::row_std_scale_kernel<<<dim3((unsigned)bx, (unsigned)n_freq), 256>>>(m->ptr, m->ld, n_seg, d_part, nchunk);
      CMF_CUDA(cudaGetLastError());
    }
    CMF_CUDA(cudaDeviceSynchronize());
    return 0;
  };
  const int rc = body();
  cached_free(d_audio); cached_free(d_win); cached_free(d_part);
  if (rc != 0) { delete m; return rc; }
  *out = m;
  return 0;
}

}  // extern "C"
