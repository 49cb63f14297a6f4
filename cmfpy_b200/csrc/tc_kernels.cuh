// tcgen05 / TMEM / TMA shift-GEMM kernels (CMF_PREC_TF32), sm_100a.
//
// All three contractions of the MU iteration run as warp-specialised persistent
// kernels: one TMA producer warp, one MMA-issuing warp (a single elected thread
// issues tcgen05.mma), four epilogue warps draining tensor memory.  A lag is
// never materialised: it is a row offset of a shared-memory operand window
//   * K-major no-swizzle "panels"  [k-chunk][row][16 B]  (row pitch 16 B: any
//     row shift is a legal descriptor start address), or
//   * MN-major SWIZZLE_128B_BASE32B rows of 128 B where the N-direction atom
//     stride is ONE ROW (overlapping atoms: lag a+1 is lag a moved by a row).
// Both were validated on hardware with tools/umma_probe.py (profiles/r01_umma_probe.log):
// swizzles are functions of the absolute shared-memory address, so shifted
// windows read what TMA wrote, and every layout used here sustains the full
// 128 cycles per 128x256x8 MMA.
//
// Every operand row is 128 bytes (32 floats).  The padded component count
// Kp in {8, 16, 32, 64, 128} is brought to that width by
//   * folding s = 32/Kp consecutive lags into "virtual components" when Kp < 32
//     (virtual lag l' covers real lags s*l' .. s*l'+s-1; its row shift is s*l'):
//       Wv[l'][n][(dl,k)] = W[s*l'+dl][n][k],  Hv[r][(dl,k)] = H^T[r-dl][k]
//     (fold_*_kernel below; Hv is an s-fold interleaved copy of the SMALL factor
//     only - X and est are never copied), or
//   * splitting into CB = Kp/32 column blocks when Kp > 32.
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmf {
namespace tc {

using namespace ptx;

constexpr int kKp = 32;              // operand row width in floats (virtual components per column block)
constexpr uint32_t kTimeoutCycles = 2000000000u;   // ~1 s: a pipeline bug must never hang the GPU

// error codes written to *err (0 = fine)
enum { kErrTimeout = 1 };

struct Abort {
  volatile int* flag;     // shared memory
  int* global_err;
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity) const {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (true) {
#pragma unroll 1
      for (int i = 0; i < 64; ++i)
        if (mbar_try_wait(bar, parity)) return true;
      if (*flag) return false;
      if ((unsigned long long)(clock64() - t0) > kTimeoutCycles) {
        *flag = 1;
        atomicExch(global_err, kErrTimeout);
        return false;
      }
    }
  }
};

// The same bounded wait for warps that have time (the epilogue warps of the two-level kernels wait for whole
// sub-chunks): they back off between polls so that their spinning does not take issue slots from the MMA and TMA
// threads that share their schedulers.
__device__ __forceinline__ bool wait_relaxed(const Abort& ab, uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (true) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      __nanosleep(128);
      if (mbar_try_wait(bar, parity)) return true;
    }
    if (*ab.flag) return false;
    if ((unsigned long long)(clock64() - t0) > kTimeoutCycles) {
      *ab.flag = 1;
      atomicExch(ab.global_err, kErrTimeout);
      return false;
    }
  }
}

// The MMA-issuing warp waits with this form: the outcome does not steer its control flow (after a timeout the
// error word is set, every later wait returns at once and the kernel runs to its end on garbage, which the host
// reports), so its loop counters, descriptors and tensor-memory addresses stay provably warp-uniform and the
// compiler keeps them in uniform registers - tcgen05.mma takes its operands from there.
__device__ __forceinline__ void wait_uniform(const Abort& ab, uint64_t* bar, uint32_t parity) {
  (void)ab.wait(bar, parity);
  __syncwarp();            // the lanes may leave the spin loop apart: converge before the uniform code goes on
}

// ---- the element-wise tail of the K1 epilogues: 32 consecutive est^T rows of one lane's feature column ----------
// Row tests are 32-bit compares against per-block counts (nv rows below t_valid hold values, the others zeros; the
// no rows below t_own enter the residual) and addresses are running pointers.  Written per element with 64-bit
// `tau < t_own` tests, `off0 + j * np` products and the store-mode branches inside, this tail was ~55 instructions per
// element, and on short reconstructions (config B: 96 MMAs per tile) the epilogue warps' instruction stream - not the
// MMAs, not memory - was the length of K1 (ncu: 8 % tensor-active, the epilogue warps never idle).
__device__ __forceinline__ int rows_below(long long limit, long long tau0) {
  const long long d = limit - tau0;
  return d <= 0 ? 0 : (d >= 32 ? 32 : (int)d);
}
// (volatile: the loads keep their program order, so all of them are in flight before the first use - left to itself
// the compiler interleaved loads and uses under the register pressure of the epilogues and serialised the latencies)
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// no >= 1.  Rows >= no are not read: the running pointer stops at row no - 1 (always valid memory) and their
// values are zeroed afterwards - one code path for every block, no per-row branches.
__device__ __forceinline__ void recon_load_x(float (&x)[32], int no, const float* __restrict__ xp,
                                             const float* __restrict__ xlop, size_t np) {
#pragma unroll
  for (int j = 0; j < 32; ++j) { x[j] = ld_stream(xp); xp += (j + 1 < no) ? np : 0; }
  if (xlop) {
    float xl[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { xl[j] = ld_stream(xlop); xlop += (j + 1 < no) ? np : 0; }
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] += xl[j];
  }
  if (no < 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = j < no ? x[j] : 0.f;
  }
}
__device__ __forceinline__ void recon_finish(const float* v, const float (&x)[32], int nv, int no, float* __restrict__ ep,
                                             float* __restrict__ elop, size_t np, bool store, bool round_out,
                                             float& tile_loss) {
  if (no > 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = (j < nv ? v[j] : 0.f) - x[j];
      if (j < no) tile_loss = fmaf(d, d, tile_loss);
    }
  }
  if (!store) return;
  if (elop) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float vv = j < nv ? v[j] : 0.f;
      const float hi = round_tf32(vv);
      *ep = hi;
      *elop = round_tf32(vv - hi);
      ep += np; elop += np;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float vv = j < nv ? v[j] : 0.f;
      if (round_out) vv = round_tf32(vv);
      *ep = vv;
      ep += np;
    }
  }
}

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int nstages) {
    if (++stage == nstages) { stage = 0; phase ^= 1; }
  }
};

// ==========================================================================
// K1  reconstruction + loss
//   est^T[tau][n] = sum_l sum_k W[l][n][k] * H^T[tau-l][k]
//   (reference cmf_predict, cmfpy/common.py:50-58, and cache_resids / loss,
//    cmfpy/algs/base.py:57-62, 90-97)
// UMMA view per tile: D[128 n][256 tau] += A_l[128 n][32 k] * B_l[256 tau][32 k]^T
//   A_l = W[l][n0..n0+128][:]  K-major SWIZZLE_128B, streamed through a ring
//   B_l = rows (L-1-l) .. of the H^T window: K-major SWIZZLE_128B rows of 128 B,
//         a lag is +128 B on the descriptor start address (swizzle is a function
//         of the absolute smem address, so shifted reads see what TMA wrote)
// Two TMEM accumulators (2 x 256 columns) ping-pong so the epilogue of tile i
// (TMEM -> registers -> est^T, fused loss against X^T) overlaps the MMAs of
// tile i+1.
// ==========================================================================
struct ReconParams {
  int Np, L, n_tiles_n, wrows;     // Np = A rows per lag; L = number of VIRTUAL lags
  int s, CB, h_shift;              // lag stride in rows, reduction blocks, B row of window row 0 minus tile start
  int cb_cols;                     // A column blocks per lag: block cb reads lag (l + cb / cb_cols), columns (cb % cb_cols)*32
  int n_rows, ld_out;              // valid output rows n, leading dimension of the output
  int store_mode;                  // 0: out[tau][n] (est^T; loss, tail mask, rounding)  2: W layout out[(l*n_rows_w+n)*Kp+k], tau = l*Kp+k
  int w_kp, w_np;                  // store_mode 2: Kp and Np of the W-layout output
  int skip_store;                  // store_mode 0: accumulate the loss only, do not write est (nothing reads it
                                   // when both denominators come from the Gram route)
  long long n_tiles;               // n_tiles_n * (RT / 256)
  long long t_own, t_valid;
  float* Et;
  const float* Xt;
  double* loss_partials;           // one per CTA
  int round_out;
  // 3xTF32 (x3 = 1): CB counts 3 * cbx reduction blocks, combo = block / cbx selects the operand halves
  // (0: A lo, B hi;  1: A hi, B lo;  2: A hi, B hi - the small cross terms first, while the accumulator is
  // small and its truncation costs nothing); the lo halves sit lo_off columns to the right in both operand arrays.
  int x3, cbx, lo_off;
  int lo_off_b;                    // lo offset of the B operand when it differs from A's (0: same as lo_off)
  int LB;                          // lags per window (0: all L).  Long lag ranges are walked in blocks of LB lags, each
                                   // with its own window of 256 + s (LB - 1) rows, into the same accumulator
  float* Elo;                      // x3: est^T = Et (hi) + Elo; null: the result is stored unsplit
  const float* Xlo;                // x3: X^T = Xt (hi) + Xlo
  int sub_units;                   // tc_recon_x3_kernel: (lag, reduction block) units per tensor-memory sub-chunk
  int* err;
};

// (reduction block) -> (column block within a combo, column offsets of the A and B halves)
struct X3Sel { int cbr, a_off, b_off; };
__device__ __forceinline__ X3Sel x3_select(int x3, int cbx, int lo_off, int lo_off_b, int vcb) {
  if (!x3) return X3Sel{vcb, 0, 0};
  const int combo = vcb / cbx;
  return X3Sel{vcb - combo * cbx, combo == 0 ? lo_off : 0, combo == 1 ? (lo_off_b ? lo_off_b : lo_off) : 0};
}

constexpr int kReconLagsPerStage = 2;              // lags per pipeline stage: 8 MMAs per barrier round trip
constexpr int kReconStages = 3;
constexpr int kReconThreads = 192;
constexpr int kReconABytes = 128 * kKp * 4;       // 16 KB per lag
constexpr int kReconStageBytes = kReconLagsPerStage * kReconABytes;

__host__ __device__ inline size_t recon_smem_bytes(int wrows) {
  return 1024 + (size_t)kReconStages * kReconStageBytes + 2 * (size_t)wrows * kKp * 4 + 256;
}

// Plain TF32 (one operand pass, one tensor-memory chain per tile); the 3xTF32 mode runs tc_recon_x3_kernel
// (tc_strict_kernels.cuh) on the same tiles and operands.
__global__ void __launch_bounds__(kReconThreads, 1)
tc_recon_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                const ReconParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = smem;                                           // [stages][2 lags x 16 KB]
  uint8_t* Hs = As + kReconStages * kReconStageBytes;           // [2][wrows * 128]
  const uint32_t hbytes = (uint32_t)p.wrows * kKp * 4;
  uint64_t* bars = (uint64_t*)(Hs + 2 * hbytes);
  uint64_t* full = bars;                                        // [stages]
  uint64_t* empty = bars + kReconStages;                        // [stages]
  uint64_t* hfull = bars + 2 * kReconStages;                    // [2]
  uint64_t* hempty = hfull + 2;                                 // [2]
  uint64_t* tfull = hempty + 2;                                 // [2]
  uint64_t* tempty = tfull + 2;                                 // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  volatile int* abort_flag = (volatile int*)(tmem_slot + 1);
  __shared__ double red[4];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler, too
  if (tid == 0) {
    for (int i = 0; i < kReconStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4);
    }
    *abort_flag = 0;
    fence_mbar_init();
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmH);
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const Abort ab{abort_flag, p.err};
  const int L = p.L, wrows = p.wrows;
  const int LB = (p.LB > 0 && p.LB < L) ? p.LB : L;             // lags per window
  const int n_lb = (L + LB - 1) / LB;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    {
      // all 32 lanes run the loop in uniform control flow; one elected lane issues the TMA instructions
      // (see mma_tf32_ss in sm100_ptx.cuh for why)
      PipeState ps;
      struct Chunk { long long tile; int cb, lb; bool valid; };
      auto next_chunk = [&](Chunk c) {
        if (++c.lb >= n_lb) {
          c.lb = 0;
          if (++c.cb >= p.CB) { c.cb = 0; c.tile += gridDim.x; c.valid = c.tile < p.n_tiles; }
        }
        return c;
      };
      // window of chunk number wc (its buffer is known to be free)
      auto issue_window = [&](const Chunk& c, long long wc) {
        const int hb = (int)(wc & 1);
        const long long tt = c.tile / p.n_tiles_n;
        uint8_t* hdst = Hs + (size_t)hb * hbytes;
        const X3Sel sel = x3_select(0, p.cbx, p.lo_off, p.lo_off_b, c.cb);
        const int l1 = min(L, (c.lb + 1) * LB);                 // window row 0 holds lag l1 - 1 of this block
        if (elect_one()) {
          mbar_arrive_expect_tx(&hfull[hb], hbytes);
          for (int rb = 0; rb < wrows / 64; ++rb)
            tma_load_2d(hdst + (size_t)rb * 64 * 128, &tmH, &hfull[hb], sel.cbr * 32 + sel.b_off,
                        (int)(tt * 256 + p.h_shift + p.s * (L - l1) + rb * 64));
        }
      };
      Chunk cur{(long long)blockIdx.x, 0, 0, (long long)blockIdx.x < p.n_tiles};
      long long wc = 0;
      if (cur.valid) issue_window(cur, 0);
      while (cur.valid) {
        const Chunk nxt = next_chunk(cur);
        bool prefetched = !nxt.valid;
        const int nhb = (int)((wc + 1) & 1);
        const uint32_t npar = (uint32_t)(((wc + 1) >> 1) & 1) ^ 1;
        const int nt = (int)(cur.tile % p.n_tiles_n);
        const X3Sel sel = x3_select(0, p.cbx, p.lo_off, p.lo_off_b, cur.cb);
        const int l0 = cur.lb * LB, l1 = min(L, l0 + LB);
        for (int l = l0; l < l1; l += kReconLagsPerStage) {
          wait_uniform(ab, &empty[ps.stage], ps.phase ^ 1);
          const int nl = min(kReconLagsPerStage, l1 - l);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[ps.stage], nl * kReconABytes);
            for (int u = 0; u < nl; ++u)
              tma_load_2d(As + (size_t)ps.stage * kReconStageBytes + u * kReconABytes, &tmW, &full[ps.stage],
                          (sel.cbr % p.cb_cols) * 32 + sel.a_off, (l + u + sel.cbr / p.cb_cols) * p.Np + nt * 128);
          }
          ps.advance(kReconStages);
          // The next window goes out as soon as its buffer is free.  (Blocking on it here, as round 1 did after the
          // second stage of a chunk, parks the producer until the MMAs of the PREVIOUS chunk retire while the W ring
          // runs dry.)
          if (!prefetched && __shfl_sync(0xffffffffu, (int)mbar_test_wait(&hempty[nhb], npar), 0)) {
            issue_window(nxt, wc + 1);
            prefetched = true;
          }
        }
        if (!prefetched) {
          wait_uniform(ab, &hempty[nhb], npar);
          issue_window(nxt, wc + 1);
        }
        cur = nxt;
        ++wc;
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    {
      const uint32_t idesc = make_idesc_tf32(128, 256, 0, 0);
      // ONE thread issues every MMA of the CTA: its instruction stream is kept to a few adds per MMA (descriptors
      // differ only in their 14-bit start-address field, bytes >> 4), or the tensor pipe ends up waiting for it
      const uint64_t adesc0 = make_smem_desc(smem_u32(As), 16, 1024, kSwz128);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(Hs), 16, 1024, kSwz128);
      PipeState ps;
      uint32_t hb = 0, hph = 0;
      int it = 0;
      for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        wait_uniform(ab, &tempty[b], (uint32_t)((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dtm = tmem + (uint32_t)b * 256;
        uint32_t acc = 0;
        for (int chunk = 0; chunk < p.CB * n_lb; ++chunk) {
          const int lb = chunk % n_lb;
          const int l0 = lb * LB, l1 = min(L, l0 + LB);
          wait_uniform(ab, &hfull[hb], hph);
          tc_fence_after();
          uint64_t bd = bdesc0 + (uint64_t)((hb * hbytes) >> 4) + (uint64_t)((uint32_t)(p.s * (l1 - 1 - l0)) * 8);
          for (int l = l0; l < l1; l += kReconLagsPerStage) {
            wait_uniform(ab, &full[ps.stage], ps.phase);
            tc_fence_after();
            const int nl = min(kReconLagsPerStage, l1 - l);
            uint64_t ad = adesc0 + (uint64_t)(ps.stage * (kReconStageBytes >> 4));
            for (int u = 0; u < nl; ++u) {
              if (elect_one()) {
                mma_tf32_ss(dtm, ad, bd, idesc, acc);
                mma_tf32_ss(dtm, ad + 2, bd + 2, idesc, 1u);
                mma_tf32_ss(dtm, ad + 4, bd + 4, idesc, 1u);
                mma_tf32_ss(dtm, ad + 6, bd + 6, idesc, 1u);
              }
              acc = 1u;
              ad += kReconABytes >> 4;
              bd -= (uint64_t)((uint32_t)p.s * 8);
            }
            if (elect_one()) mma_commit(&empty[ps.stage]);
            ps.advance(kReconStages);
          }
          if (elect_one()) mma_commit(&hempty[hb]);
          hb ^= 1;
          hph ^= (hb == 0);
        }
        if (elect_one()) mma_commit(&tfull[b]);
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> est^T, fused loss ----------------
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    double loss_acc = 0.0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int nt = (int)(tile % p.n_tiles_n);
      const long long tt = tile / p.n_tiles_n;
      const int b = it & 1;
      if (!ab.wait(&tfull[b], (it >> 1) & 1)) break;
      tc_fence_after();
      const int n = nt * 128 + q * 32 + lane;
      const bool n_ok = n < p.n_rows;
      float tile_loss = 0.f;
      const float* __restrict__ Xt = p.Xt;
      float* __restrict__ Et = p.Et;
      const size_t np = (size_t)p.ld_out;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        float x[32];
        tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + c * 32), r);
        const long long tau0 = tt * 256 + c * 32;
        const size_t off0 = (size_t)tau0 * np + n;
        const int nv = rows_below(p.t_valid, tau0), no = rows_below(p.t_own, tau0);
        // all 32 X loads are issued before anything depends on them
        if (n_ok && no > 0) recon_load_x(x, no, Xt + off0, nullptr, np);
        tmem_ld_wait();
        if (p.store_mode == 2) {
          // tau = l*Kp + k: 32 consecutive tau are whole groups of 4 components of one lag
          if (n_ok) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const long long tau = tau0 + 4 * j4;
              const long long l = tau / p.w_kp;
              const int k = (int)(tau % p.w_kp);
              if (tau < p.t_valid)
                *reinterpret_cast<float4*>(Et + ((size_t)l * p.w_np + n) * p.w_kp + k) =
                    make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]),
                                __uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3]));
            }
          }
        } else if (n_ok) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          recon_finish(v, x, nv, no, Et + off0, nullptr, np, !p.skip_store, p.round_out != 0, tile_loss);
        }
      }
      loss_acc += (double)tile_loss;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
    }
    // block partial of the loss (epilogue warps only)
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) red[q] = loss_acc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 2 && lane == 0) p.loss_partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ==========================================================================
// K2  W terms
//   out[src][l][n][k] = sum_tau S^T[tau][n] * H^T[tau-l][k],  S = X | est
//   (reference _compute_mult_W, cmfpy/algs/mult.py:27-40)
// UMMA view per work item (128 n, 16 lags, one source, a chunk of time):
//   D[128 n][2 x (8 lags x 32 k)] += A[128 n][8 tau]^T-major * B[(lag,k)][8 tau]
//   A = S^T rows, MN-major SWIZZLE_128B_BASE32B (4 regions of 32 n)
//   B = H^T rows, MN-major SWIZZLE_128B_BASE32B, N-atom stride = one row:
//       atom a of MMA g is lag  l0 + 8g + 7 - a
// Partial sums per time chunk go to a scratch buffer and are reduced in a fixed
// order by sum_splits_kernel (deterministic; no float atomics).
// ==========================================================================
struct WTermsParams {
  int Np, L, n_tiles_n, n_lag_groups, n_chunks, h;   // L = REAL lags; lag groups hold 16 virtual lags
  int Lv, Kp, s, CB, brows;       // virtual lags, real padded K, lag stride, column blocks, H rows per stage
  int n_src;                      // 2: X and est; 1: the first source only
  long long n_items;              // n_tiles_n * n_lag_groups * CB * 2 * n_chunks
  long long stages_total;         // ceil(t_own / 32)
  float* part;                    // [chunk][src][L][Np][Kp]
  long long per_src;              // L * Np * Kp
  int x3, lo_off;                 // tc_wterms_x3_kernel: three passes over the item's time range - (S lo, H hi),
                                  // (S hi, H lo), (S hi, H hi); H lo sits lo_off columns to the right
  int sub_units;                  // tc_wterms_x3_kernel: 32-row time stages per tensor-memory sub-chunk
  // quad mode (data with at most 32 features: the lag autocorrelation of H^T): the four 32-row quarters of the
  // accumulator hold the SAME features at four time shifts - quarter r reads S^T rows tau + r * LPR * s, which
  // makes its columns lags r * LPR .. - so one item covers 4 x LPR virtual lags and no MMA row is idle.
  // The time range starts stage0 (<= 0) stages early so that every quarter sees all of its rows.
  int quad;
  long long stage0;
  int* err;
};

constexpr int kWtStages = 6;
constexpr int kWtThreads = 192;
constexpr int kWtABytes = 4 * 32 * 128;           // 128 n x 32 tau
// H rows per stage: 32 tau + 15 virtual lags of stride s, padded to 8
__host__ __device__ inline int wterms_brows(int s) { return ((32 + 15 * s + 7) / 8) * 8; }
__host__ __device__ inline size_t wterms_stage_bytes(int s) { return (size_t)kWtABytes + (size_t)wterms_brows(s) * 128; }
__host__ __device__ inline size_t wterms_smem_bytes(int s) { return 1024 + (size_t)kWtStages * wterms_stage_bytes(s) + 256; }

__global__ void __launch_bounds__(kWtThreads, 1)
tc_wterms_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmE,
                 const __grid_constant__ CUtensorMap tmH, const WTermsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* St = smem;                                            // [stages][A 16 KB | B brows x 128 B]
  const uint32_t kWtStageBytes = (uint32_t)wterms_stage_bytes(p.s);
  uint64_t* bars = (uint64_t*)(St + (size_t)kWtStages * kWtStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWtStages;
  uint64_t* tfull = bars + 2 * kWtStages;                        // [1]
  uint64_t* tempty = tfull + 1;                                  // [1]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 1);
  volatile int* abort_flag = (volatile int*)(tmem_slot + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler, too
  if (tid == 0) {
    for (int i = 0; i < kWtStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    *abort_flag = 0;
    fence_mbar_init();
    prefetch_tmap(&tmX); prefetch_tmap(&tmE); prefetch_tmap(&tmH);
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const Abort ab{abort_flag, p.err};

  // item -> (chunk, n tile, src, lag group); lag group fastest so that the CTAs
  // that stream the same S^T rows run at the same time (L2 reuse)
  auto decode = [&](long long item, int& lg, int& cb, int& src, int& nt, int& ch) {
    lg = (int)(item % p.n_lag_groups); item /= p.n_lag_groups;
    cb = (int)(item % p.CB); item /= p.CB;
    src = (int)(item % p.n_src); item /= p.n_src;
    nt = (int)(item % p.n_tiles_n); item /= p.n_tiles_n;
    ch = (int)item;
  };
  auto chunk_range = [&](int ch, long long& s0, long long& s1) {
    const long long total = p.stages_total - p.stage0;
    const long long base = total / p.n_chunks, rem = total % p.n_chunks;
    s0 = p.stage0 + ch * base + (ch < rem ? ch : rem);
    s1 = s0 + base + (ch < rem ? 1 : 0);
  };
  const int lpi = p.quad ? 64 : 16;               // virtual lags per item

  if (warp == 0) {
    {
      PipeState ps;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int lg, cb, src, nt, ch;
        decode(item, lg, cb, src, nt, ch);
        long long s0, s1;
        chunk_range(ch, s0, s1);
        const CUtensorMap* tmS = src ? &tmE : &tmX;
        const int hcol = cb * 32;
        for (long long s = s0; s < s1; ++s) {
          wait_uniform(ab, &empty[ps.stage], ps.phase ^ 1);
          uint8_t* dst = St + (size_t)ps.stage * kWtStageBytes;
          const int tau0 = (int)(s * 32);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[ps.stage], kWtStageBytes);
            if (!p.quad) {
              tma_load_3d(dst, tmS, &full[ps.stage], 0, tau0, nt * 4);    // four 32-feature regions in one box
            } else {
#pragma unroll
              for (int r = 0; r < 4; ++r)                                 // the same features, 16 r lags later
                tma_load_2d(dst + r * 4096, tmS, &full[ps.stage], 0, tau0 + r * 16 * p.s);
            }
            // Hv rows tau0 - s*(l0+15) .. tau0 + 32; row index in Hv is tau + h
            tma_load_2d(dst + kWtABytes, &tmH, &full[ps.stage], hcol, tau0 - p.s * (lg * lpi + 15) + p.h);
          }
          ps.advance(kWtStages);
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = make_idesc_tf32(128, 256, 1, 1);
      const uint64_t adesc0 = make_smem_desc(smem_u32(St), 4096, 512, 1 /*SW128_BASE32B*/);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(St) + kWtABytes, (uint32_t)p.s * 128, 512, 1);
      const uint32_t stage16 = kWtStageBytes >> 4;
      PipeState ps;
      int it = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        int lg, cb, src, nt, ch;
        decode(item, lg, cb, src, nt, ch);
        long long s0, s1;
        chunk_range(ch, s0, s1);
        wait_uniform(ab, tempty, (it & 1) ^ 1);
        tc_fence_after();
        for (long long s = s0; s < s1; ++s) {
          wait_uniform(ab, &full[ps.stage], ps.phase);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(ps.stage * stage16);
          const uint64_t bd1 = bdesc0 + (uint64_t)(ps.stage * stage16);      // lags 8..15 of the group: rows 0..
          const uint64_t bd0 = bd1 + (uint64_t)((uint32_t)p.s * 64);         // lags 0..7: 8 s rows further down
          const uint32_t acc = s > s0 ? 1u : 0u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {                                   // k-step: A +1 KB, B +8 rows
            if (elect_one()) {
              mma_tf32_ss(tmem, ad + 64 * ks, bd0 + 64 * ks, idesc, ks ? 1u : acc);
              mma_tf32_ss(tmem + 256, ad + 64 * ks, bd1 + 64 * ks, idesc, ks ? 1u : acc);
            }
          }
          if (elect_one()) mma_commit(&empty[ps.stage]);
          ps.advance(kWtStages);
        }
        if (elect_one()) mma_commit(tfull);
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      int lg, cb, src, nt, ch;
      decode(item, lg, cb, src, nt, ch);
      if (!ab.wait(tfull, it & 1)) break;
      tc_fence_after();
      const int n = p.quad ? lane : nt * 128 + q * 32 + lane;
      float* obase = p.part + ((long long)ch * p.n_src + src) * p.per_src;
#pragma unroll 1
      for (int c = 0; c < 16; ++c) {            // 16 column blocks of 32 = (g, a): one virtual lag each
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        const int g = c >> 3, a = c & 7;
        const int lv = lg * lpi + (p.quad ? 16 * q : 0) + 8 * g + 7 - a;
        if (n < p.Np && lv < p.Lv) {
          if (p.s == 1) {                       // 32 components of column block cb, real lag lv
            float4* o = reinterpret_cast<float4*>(obase + ((long long)lv * p.Np + n) * p.Kp + cb * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          } else if (p.s == 2) {                // Kp = 16: two real lags of 16 components
#pragma unroll
            for (int dl = 0; dl < 2; ++dl) {
              const int l = 2 * lv + dl;
              if (l < p.L) {
                float4* o = reinterpret_cast<float4*>(obase + ((long long)l * p.Np + n) * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  o[j] = make_float4(__uint_as_float(r[16 * dl + 4 * j]), __uint_as_float(r[16 * dl + 4 * j + 1]),
                                     __uint_as_float(r[16 * dl + 4 * j + 2]), __uint_as_float(r[16 * dl + 4 * j + 3]));
              }
            }
          } else {                              // s == 4, Kp = 8: four real lags of 8 components
#pragma unroll
            for (int dl = 0; dl < 4; ++dl) {
              const int l = 4 * lv + dl;
              if (l < p.L) {
                float4* o = reinterpret_cast<float4*>(obase + ((long long)l * p.Np + n) * 8);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  o[j] = make_float4(__uint_as_float(r[8 * dl + 4 * j]), __uint_as_float(r[8 * dl + 4 * j + 1]),
                                     __uint_as_float(r[8 * dl + 4 * j + 2]), __uint_as_float(r[8 * dl + 4 * j + 3]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// Hv[r][(dl,k)] = round_tf32(H^T[r - dl][k])     (s > 1; rows before 0 read as zero)
__global__ void __launch_bounds__(256)
fold_h_kernel(float* __restrict__ Hv, const float* __restrict__ Ht, long long row0, long long nrows, int Kp) {
  const long long total = nrows * 32;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = row0 + i / 32;
    const int v = (int)(i % 32), dl = v / Kp, k = v % Kp;
    const float x = (r - dl >= 0) ? Ht[(r - dl) * Kp + k] : 0.f;
    Hv[r * 32 + v] = round_tf32(x);
  }
}

// Wv[l'][n][(dl,k)] = round_tf32(W[s*l' + dl][n][k])   (s > 1; lags >= L read as zero)
__global__ void __launch_bounds__(256)
fold_w_kernel(float* __restrict__ Wv, const float* __restrict__ W, int L, int Lv, int Np, int Kp, int s) {
  const long long total = (long long)Lv * Np * 32;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % 32), dl = v / Kp, k = v % Kp;
    const int n = (int)((i / 32) % Np);
    const int lv = (int)(i / (32ll * Np));
    const int l = s * lv + dl;
    Wv[i] = (l < L) ? round_tf32(W[((long long)l * Np + n) * Kp + k]) : 0.f;
  }
}

// ---- 3xTF32 operand pairs: hi = RN_tf32(x), lo = RN_tf32(x - hi)  (x - hi is exact in fp32) ----
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}

// in place: hi_inout <- hi, lo_out <- lo   (X^T after it was loaded at full precision)
__global__ void __launch_bounds__(256)
split_inplace_kernel(float4* __restrict__ hi_inout, float4* __restrict__ lo_out, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = hi_inout[i];
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    hi_inout[i] = h;
    lo_out[i] = l;
  }
}

// Hv[r][c] = hi, Hv[r][KW + c] = lo of the (folded) H^T entry that column c of row r holds:
//   s > 1: c = (dl, k), entry H^T[r - dl][k];   s == 1: entry H^T[r][c]  (KW == Kp)
__global__ void __launch_bounds__(256)
fold_h_x3_kernel(float* __restrict__ Hv, const float* __restrict__ Ht, long long row0, long long nrows, int Kp, int s,
                 int KW) {
  const long long total = nrows * KW;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = row0 + i / KW;
    const int c = (int)(i % KW);
    const int dl = s > 1 ? c / Kp : 0, k = s > 1 ? c % Kp : c;
    const float x = (r - dl >= 0) ? Ht[(r - dl) * Kp + k] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    Hv[r * 2 * KW + c] = hi;
    Hv[r * 2 * KW + KW + c] = lo;
  }
}

// Wv[l'][n][c] = hi, Wv[l'][n][KW + c] = lo of W[s*l' + dl][n][k]  (lags >= L read as zero)
__global__ void __launch_bounds__(256)
fold_w_x3_kernel(float* __restrict__ Wv, const float* __restrict__ W, int L, int Lv, int Np, int Kp, int s, int KW) {
  const long long total = (long long)Lv * Np * KW;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % KW);
    const int dl = s > 1 ? c / Kp : 0, k = s > 1 ? c % Kp : c;
    const long long row = i / KW;                    // l' * Np + n
    const int n = (int)(row % Np);
    const int lv = (int)(row / Np);
    const int l = s * lv + dl;
    const float x = (l < L) ? W[((long long)l * Np + n) * Kp + k] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    Wv[row * 2 * KW + c] = hi;
    Wv[row * 2 * KW + KW + c] = lo;
  }
}

// ---------------------------------------------------------------------------
// Gram route for the denominators (exact identities, ~K/N of the direct cost):
//   den_H[k][t] = sum_d sum_k' R[d][k][k'] H[k'][t+d]  -  (terms of est past the end of the data)
//   R[d][k][k'] = sum_{l-l'=d} sum_n W[l][n][k] W[l'][n][k']
// (substitute est = sum_l' W[l'] shift(H,l') into tensor_transconv(W, est), reference common.py:61-86).
// R comes from G = Wt Wt^T (a plain GEMM on the recon kernel) summed along its lag diagonals.
// ---------------------------------------------------------------------------

// W step:  den_W[l][n][k] = sum_{l',k'} W[l'][n][k'] A[l-l'][k'][k]  -  (terms of est past the end of the data)
//   A[d][k'][k] = sum_u H[k'][u+d] H[k][u]  (lag autocorrelation of H) = P[d][k'][k] for d >= 0, P[-d][k][k'] for d < 0,
//   P[d][k'][k] = sum_t H[k'][t] H[k][t-d]: the W-terms kernel run on H^T itself.
// Mt[(l,k)][(l'v, c)] = round_tf32(A[l - l'][k'][k]) with (l', k') the real lag / component that virtual lag l'v,
// column c of Wv holds - the K-major B operand of a plain GEMM with Wv.
// x3: every row is [hi (Lv KW) | lo (Lv KW)].
__global__ void __launch_bounds__(256)
toeplitz_kernel(const float* __restrict__ P, float* __restrict__ Mt, int L, int Lv, int Kp, int s, int KW,
                long long rows_alloc, int x3) {
  const long long ld = (long long)Lv * KW;
  const long long total = rows_alloc * ld;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long tau = i / ld;
    const int col = (int)(i % ld);
    const int lv = col / KW, c = col % KW;
    const int dl = (s > 1) ? c / Kp : 0;
    const int kq = (s > 1) ? c % Kp : c;
    const int lp = s * lv + dl;
    const int l = (int)(tau / Kp), k = (int)(tau % Kp);
    float v = 0.f;
    if (l < L && lp < L && kq < Kp) {
      const int d = l - lp;
      v = (d >= 0) ? P[((size_t)d * Kp + kq) * Kp + k] : P[((size_t)(-d) * Kp + k) * Kp + kq];
    }
    if (x3) {
      const float hi = round_tf32(v);
      Mt[tau * 2 * ld + col] = hi;
      Mt[tau * 2 * ld + ld + col] = round_tf32(v - hi);
    } else {
      Mt[i] = round_tf32(v);
    }
  }
}

// The same operator as a SHIFT operand (round 2): den_W is itself a convolution over lags,
//   den_W[(l,k)][n] = sum_{l'v} sum_c Wv[l'v][n][c] * Pw[(l,k) + s_rows (Lv-1-l'v)][c],
//   Pw[rho][(dl,k')] = A[rho / Kp - s (Lv-1) - dl][k'][rho % Kp]   (zero outside |d| <= L-1),
// i.e. exactly the reconstruction est = sum_l W_l shift(H, l) with the lag autocorrelation in the place of H^T, (l,k)
// in the place of time and a lag stride of s_rows = max(32, Kp) rows: K1 reads every lag as a shifted window of ONE
// small operand ((L + s Lv) Kp rows x KW columns) instead of streaming a 2048 x 2048 block-Toeplitz matrix of which
// each 32 KB window fed 4 MMAs (0.245 -> see DESIGN.md 4).  x3: rows are [hi (KW) | lo (KW)].
__global__ void __launch_bounds__(256)
autocorr_shift_operand_kernel(const float* __restrict__ P, float* __restrict__ Pw, int L, int Lv, int Kp, int s, int KW,
                              long long rows_alloc, int x3) {
  const long long total = rows_alloc * KW;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int ld = (x3 ? 2 : 1) * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long rho = i / KW;
    const int c = (int)(i % KW);
    const int dl = (s > 1) ? c / Kp : 0;
    const int kq = (s > 1) ? c % Kp : c;
    const int k = (int)(rho % Kp);
    const long long d = rho / Kp - (long long)s * (Lv - 1) - dl;
    float v = 0.f;
    if (kq < Kp && d > -(long long)L && d < (long long)L)
      v = (d >= 0) ? P[((size_t)d * Kp + kq) * Kp + k] : P[((size_t)(-d) * Kp + k) * Kp + kq];
    if (x3) {
      const float hi = round_tf32(v);
      Pw[rho * ld + c] = hi;
      Pw[rho * ld + KW + c] = round_tf32(v - hi);
    } else {
      Pw[rho * ld + c] = round_tf32(v);
    }
  }
}

// Wt[(l*Kp + k)][n] = round_tf32(W[l][n][k])      (32x32 tiles through smem)
// x3: rows are [hi (lo_off columns) | lo]
__global__ void __launch_bounds__(256)
transpose_round_w_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Np, int Kp, long long ldt, int lo_off) {
  __shared__ float tile[32][33];
  const int l = blockIdx.z;
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int n = n0 + ty + j, k = k0 + tx;
    tile[ty + j][tx] = (n < Np && k < Kp) ? W[((size_t)l * Np + n) * Kp + k] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int k = k0 + ty + j, n = n0 + tx;
    if (k < Kp && n < Np) {
      const float v = tile[tx][ty + j], hi = round_tf32(v);
      Wt[((size_t)l * Kp + k) * ldt + n] = hi;
      if (lo_off) Wt[((size_t)l * Kp + k) * ldt + lo_off + n] = round_tf32(v - hi);
    }
  }
}

// Rt[l][k'][k] = R[l-(L-1)][k][k'] = sum_{l'} G[(l'+d)*Kp + k][l'*Kp + k'],  d = l - (L-1), l in [0, 2L-1):
// R as a W-like operand (lag, "feature" k', component k) for the H-terms kernel, which then computes
// den_H^T[t][k] = sum_l sum_k' Rt[l][k'][k] H^T[t + d][k']
__global__ void __launch_bounds__(256)
diag_sum_kernel(const float* __restrict__ G, long long ldg, float* __restrict__ Rt, int L, int Kp) {
  const long long total = (long long)(2 * L - 1) * Kp * Kp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % Kp);
    const int kq = (int)((i / Kp) % Kp);
    const int l = (int)(i / ((long long)Kp * Kp));
    const int d = l - (L - 1);
    const int lo = d < 0 ? -d : 0, hi = d > 0 ? L - d : L;     // l' range with 0 <= l'+d < L
    float acc = 0.f;
    for (int lp = lo; lp < hi; ++lp) acc += G[((size_t)(lp + d) * Kp + k) * ldg + (size_t)lp * Kp + kq];
    Rt[i] = acc;
  }
}

}  // namespace tc
}  // namespace cmf
