// tcgen05 shift-GEMM kernels with TWO-LEVEL accumulation (the fp32-grade 3xTF32 mode, and K3 in every mode).
//
// Tensor memory adds with round-toward-zero: a chain of n MMA steps of non-negative terms carries a relative bias
// of about -5.7e-8 n (tools/acc_bias_probe.py), which no operand splitting removes.  These kernels therefore never
// let a chain grow: the MMA warp accumulates SUB-CHUNKS of `sub_units` units (a unit = 4 MMA k-steps = one lag or one
// 32-row time stage) into one of two 128 x 256 tensor-memory buffers S0 / S1, alternating; eight epilogue warps
// (TMEM lane quarter x column half) fold every finished sub-chunk into a MASTER accumulator of 128 fp32 registers
// per thread with round-to-nearest adds while the MMAs of the next sub-chunk run into the other buffer.  The bias
// of a result is then that of ONE sub-chunk (8 units = 32 steps: ~7e-7, independent of L, K, N and T), the fold is
// hidden behind the MMAs, and nothing but the final result leaves the SM.
//
//   tc_recon_x3_kernel   K1  reconstruction + loss           (reference common.py:50-58, base.py:57-62, 90-97)
//   tc_wterms_x3_kernel  K2  W terms, 8 lags per work item   (reference mult.py:35-38)
//   tc_hterms_kernel     K3  H terms, every precision mode   (reference common.py:61-86)
//
// Layouts, folding (s, CB) and the shifted-window descriptors are those of tc_kernels.cuh.
#pragma once
#include "tc_kernels.cuh"

namespace cmf {
namespace tc {

constexpr int kSThreads = 384;        // warpgroup 0: TMA producer (warp 0), MMA issuer (warp 1), TMEM allocator (warp 2);
                                      // warpgroups 1-2: epilogue.  setmaxnreg moves registers from the first to the others
constexpr int kSEpiThreads = 256;
#ifndef CMF_CTL_REGS
#define CMF_CTL_REGS 40
#define CMF_EPI_REGS 232
#endif
constexpr int kSCtlRegs = CMF_CTL_REGS, kSEpiRegs = CMF_EPI_REGS;   // 128 * 40 + 256 * 232 = 64512 <= 65536
constexpr int kStrictSubUnits = 8;    // units per sub-chunk in the 3xTF32 mode (32 MMA k-steps)

// Orders the 32 adds of one block before the next tensor-memory load: without it the compiler issues all four
// loads first and needs 128 staging registers next to the 128 of the master.
__device__ __forceinline__ void pin32(float* v) {
  asm volatile(""
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                 "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]),
                 "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                 "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
               :
               : "memory");
}

// master (+)= S[32 lanes of this warp][128 columns at taddr]
template <bool kFirst>
__device__ __forceinline__ void fold_sub(float (&m)[128], uint32_t taddr) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + (uint32_t)(i * 32), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      m[i * 32 + j] = kFirst ? __uint_as_float(r[j]) : m[i * 32 + j] + __uint_as_float(r[j]);
    pin32(&m[i * 32]);
  }
}

// m[0..95] <- m[32..127]: lets a loop that is NOT unrolled walk the master accumulator 32 columns at a time with
// static register indices.  (Unrolling the epilogues four-fold made these kernels ~250 KB of code; the eight
// epilogue warps then evicted the MMA and TMA loops from the instruction caches and the tensor pipe starved:
// ncu showed 2.6 no-instruction stalls per issue against 0.12 for the single-chain kernel.)
__device__ __forceinline__ void rotate32(float (&m)[128]) {
#pragma unroll
  for (int j = 0; j < 96; ++j) m[j] = m[j + 32];
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// The MMA side of the sub-chunk protocol: a sub-chunk goes to buffer (sc & 1); its first unit waits until the
// epilogue has folded the buffer's previous contents, its last one commits it.  The first n_cross units of a work
// item - the two cross-term passes of 3xTF32, a_lo b_hi and a_hi b_lo, 2^-11 of the result - form ONE sub-chunk
// (their truncation is relative to their own small partial sum); the a_hi b_hi units after them are cut every
// sub_units units.
struct SubPlan {
  int n_units, n_cross, sub_units;
  __device__ __forceinline__ SubPlan(long long units, long long cross, int sub) {
    n_units = (int)units;
    sub_units = sub > 0 ? sub : (int)units;
    n_cross = sub > 0 ? (int)cross : 0;
  }
  __device__ __forceinline__ int n_sub() const {
    return (n_cross > 0 ? 1 : 0) + (n_units - n_cross + sub_units - 1) / sub_units;
  }
};
// (kept to a handful of 32-bit instructions per unit: ONE thread issues every MMA of the CTA, and at four MMAs -
// 512 tensor-pipe cycles - per unit its own instruction stream is what the tensor pipe ends up waiting for)
struct SubIssue {
  uint32_t b = 0, ph = 0;   // buffer of the current sub-chunk; parity of that buffer's use count
  int us = 0;               // units already in the current sub-chunk
  __device__ __forceinline__ uint32_t begin_unit(const Abort& ab, uint64_t* sempty, uint32_t tmem) {
    if (us == 0) {
      wait_uniform(ab, &sempty[b], ph ^ 1);
      tc_fence_after();
    }
    return tmem + b * 256;
  }
  // `done` = units of the item issued so far, this one included
  __device__ __forceinline__ void end_unit(uint64_t* sfull, const SubPlan& pl, int done) {
    ++us;
    if (done == pl.n_units || done == pl.n_cross || (done > pl.n_cross && us == pl.sub_units)) {
      if (elect_one()) mma_commit(&sfull[b]);
      us = 0;
      b ^= 1;
      ph ^= (b == 0);
    }
  }
};
// the epilogue side of the same sequence
struct SubDrain {
  uint32_t b = 0, ph = 0;
  __device__ __forceinline__ void next() { b ^= 1; ph ^= (b == 0); }
};

// ==========================================================================
// K1, two-level accumulation.  Same tiles, operands and stores as tc_recon_kernel; a unit is one lag of one
// reduction block (4 MMAs of 128 x 256 x 8).
// ==========================================================================
__global__ void __launch_bounds__(kSThreads, 1)
tc_recon_x3_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
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
  uint64_t* sfull = hempty + 2;                                 // [2]
  uint64_t* sempty = sfull + 2;                                 // [2]
  uint32_t* tmem_slot = (uint32_t*)(sempty + 2);
  volatile int* abort_flag = (volatile int*)(tmem_slot + 1);
  __shared__ double red[8];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler, too
  if (tid == 0) {
    for (int i = 0; i < kReconStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 1);
      mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 8);
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
  const SubPlan plan((long long)p.CB * L, p.x3 ? 2ll * p.cbx * L : 0, p.sub_units);     // units per tile

  if (warp == 0) {
    reg_dec<kSCtlRegs>();
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
        const X3Sel sel = x3_select(p.x3, p.cbx, p.lo_off, p.lo_off_b, c.cb);
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
        const X3Sel sel = x3_select(p.x3, p.cbx, p.lo_off, p.lo_off_b, cur.cb);
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
    reg_dec<kSCtlRegs>();
    // ---------------- MMA issuer ----------------
    {
      const uint32_t idesc = make_idesc_tf32(128, 256, 0, 0);
      // descriptors differ only in their 14-bit start-address field (bytes >> 4): one add each per MMA
      const uint64_t adesc0 = make_smem_desc(smem_u32(As), 16, 1024, kSwz128);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(Hs), 16, 1024, kSwz128);
      PipeState ps;
      SubIssue si;
      uint32_t hb = 0, hph = 0;
      for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int unit = 0;
        for (int chunk = 0; chunk < p.CB * n_lb; ++chunk) {
          const int lb = chunk % n_lb;
          const int l0 = lb * LB, l1 = min(L, l0 + LB);
          wait_uniform(ab, &hfull[hb], hph);
          tc_fence_after();
          const uint64_t bwin = bdesc0 + (uint64_t)((hb * hbytes) >> 4);
          uint32_t brow8 = (uint32_t)(p.s * (l1 - 1 - l0)) * 8;              // (row shift * 128 B) >> 4
          for (int l = l0; l < l1; l += kReconLagsPerStage) {
            wait_uniform(ab, &full[ps.stage], ps.phase);
            tc_fence_after();
            const int nl = min(kReconLagsPerStage, l1 - l);
            uint64_t ad = adesc0 + (uint64_t)(ps.stage * (kReconStageBytes >> 4));
            for (int u = 0; u < nl; ++u) {
              const uint32_t dtm = si.begin_unit(ab, sempty, tmem);
              const uint64_t bd = bwin + brow8;
              if (elect_one()) {
                mma_tf32_ss(dtm, ad, bd, idesc, si.us != 0 ? 1u : 0u);
                mma_tf32_ss(dtm, ad + 2, bd + 2, idesc, 1u);
                mma_tf32_ss(dtm, ad + 4, bd + 4, idesc, 1u);
                mma_tf32_ss(dtm, ad + 6, bd + 6, idesc, 1u);
              }
              si.end_unit(sfull, plan, ++unit);
              ad += kReconABytes >> 4;
              brow8 -= (uint32_t)p.s * 8;
            }
            if (elect_one()) mma_commit(&empty[ps.stage]);
            ps.advance(kReconStages);
          }
          if (elect_one()) mma_commit(&hempty[hb]);
          hb ^= 1;
          hph ^= (hb == 0);
        }
      }
    }
  } else if (warp < 4) {
    reg_dec<kSCtlRegs>();
  } else {
    reg_inc<kSEpiRegs>();
    // ---------------- epilogue: fold sub-chunks, then master -> est^T, fused loss ----------------
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;             // column half of the 256-column buffers
    const int n_sub = plan.n_sub();
    float m[128];
    double loss_acc = 0.0;
    SubDrain sd;
    bool ok = true;
    for (long long tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      const int nt = (int)(tile % p.n_tiles_n);
      const long long tt = tile / p.n_tiles_n;
      const float* __restrict__ Xt = p.Xt;
      const float* __restrict__ Xlo = p.Xlo;
      const size_t np = (size_t)p.ld_out;
      // The X rows of this CTA's NEXT tile go to L2 now (one bulk prefetch of a 512-byte row segment per thread and
      // array): the loads of the residual below are issued in bursts between the folds and the stores, so on problems
      // whose reconstruction is short (config B: 96 MMAs per tile) their DRAM latency was part of the epilogue.
      // (warp-uniform addresses, one elected lane: the instruction takes uniform registers, per-lane addresses would be
      // issued one lane at a time)
      if (p.t_own > 0 && tile + gridDim.x < p.n_tiles) {
        const long long ntile = tile + gridDim.x;
        const int col = (int)(ntile % p.n_tiles_n) * 128;
        const long long row0 = (ntile / p.n_tiles_n) * 256 + (warp - 4) * 32;
        if (col < p.n_rows) {
          const uint32_t bytes = (uint32_t)min(128, p.n_rows - col) * 4;
          const float* px = Xt + (size_t)row0 * np + col;
          const float* pl = Xlo ? Xlo + (size_t)row0 * np + col : nullptr;
#pragma unroll 1
          for (int r = 0; r < 32; ++r) {
            if (row0 + r < p.t_own && elect_one()) {
              prefetch_l2_bulk(px, bytes);
              if (pl) prefetch_l2_bulk(pl, bytes);
            }
            px += np;
            if (pl) pl += np;
          }
        }
      }
      for (int sub = 0; sub < n_sub; ++sub, sd.next()) {
        if (!wait_relaxed(ab, &sfull[sd.b], sd.ph)) { ok = false; break; }
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sd.b * 256 + half * 128);
        if (sub == 0) fold_sub<true>(m, taddr); else fold_sub<false>(m, taddr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[sd.b]);
      }
      if (!ok) break;
      const int n = nt * 128 + q * 32 + lane;
      const bool n_ok = n < p.n_rows;
      float tile_loss = 0.f;
      float* __restrict__ Et = p.Et;
      float* __restrict__ Elo = p.Elo;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const long long tau0 = tt * 256 + half * 128 + c * 32;
        const size_t off0 = (size_t)tau0 * np + n;
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
                    make_float4(m[4 * j4], m[4 * j4 + 1], m[4 * j4 + 2], m[4 * j4 + 3]);
            }
          }
        } else if (n_ok) {
          // (t_own is 0 whenever there is no X: the plain GEMMs of the Gram route, the reconstruction between the W and
          // the H step)
          const int nv = rows_below(p.t_valid, tau0), no = rows_below(p.t_own, tau0);
          float x[32];
          if (no > 0) recon_load_x(x, no, Xt + off0, Xlo ? Xlo + off0 : nullptr, np);
          recon_finish(m, x, nv, no, Et + off0, Elo ? Elo + off0 : nullptr, np, !p.skip_store, p.round_out != 0, tile_loss);
        }
        rotate32(m);
      }
      loss_acc += (double)tile_loss;
    }
    // block partial of the loss (epilogue warps only)
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) red[warp - 4] = loss_acc;
    epi_bar();
    if (warp == 4 && lane == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i];
      p.loss_partials[blockIdx.x] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ==========================================================================
// K2, two-level accumulation.  Work item = (128 features, 8 virtual lags, X or est, a chunk of time): one
// 128 x 256 buffer per sub-chunk, so consecutive sub-chunks alternate between S0 and S1.  A unit is one 32-row
// time stage (4 MMAs).  Output as tc_wterms_kernel: part[chunk][src][L][Np][Kp].
// ==========================================================================
__host__ __device__ inline int wterms8_brows(int s) { return ((32 + 7 * s + 7) / 8) * 8; }
__host__ __device__ inline size_t wterms8_stage_bytes(int s) { return (size_t)kWtABytes + (size_t)wterms8_brows(s) * 128; }
__host__ __device__ inline size_t wterms8_smem_bytes(int s) { return 1024 + (size_t)kWtStages * wterms8_stage_bytes(s) + 256; }

__global__ void __launch_bounds__(kSThreads, 1)
tc_wterms_x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmE,
                    const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmXlo,
                    const __grid_constant__ CUtensorMap tmElo, const WTermsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* St = smem;                                            // [stages][A 16 KB | B brows x 128 B]
  const uint32_t stage_bytes = (uint32_t)wterms8_stage_bytes(p.s);
  uint64_t* bars = (uint64_t*)(St + (size_t)kWtStages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWtStages;
  uint64_t* sfull = bars + 2 * kWtStages;                        // [2]
  uint64_t* sempty = sfull + 2;                                  // [2]
  uint32_t* tmem_slot = (uint32_t*)(sempty + 2);
  volatile int* abort_flag = (volatile int*)(tmem_slot + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler, too
  if (tid == 0) {
    for (int i = 0; i < kWtStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 8); }
    *abort_flag = 0;
    fence_mbar_init();
    prefetch_tmap(&tmX); prefetch_tmap(&tmE); prefetch_tmap(&tmH); prefetch_tmap(&tmXlo); prefetch_tmap(&tmElo);
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const Abort ab{abort_flag, p.err};
  const int n_pass = p.x3 ? 3 : 1;

  // item -> (chunk, n tile, src, column block, lag group); lag group fastest so that the CTAs that stream the
  // same S^T rows run at the same time (L2 reuse)
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
  const int lpi = p.quad ? 32 : 8;                // virtual lags per item (quad mode: see WTermsParams)

  if (warp == 0) {
    reg_dec<kSCtlRegs>();
    {
      PipeState ps;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int lg, cb, src, nt, ch;
        decode(item, lg, cb, src, nt, ch);
        long long s0, s1;
        chunk_range(ch, s0, s1);
        // 3xTF32: (S lo, H hi), (S hi, H lo), (S hi, H hi)
        for (int combo = p.x3 ? 0 : 2; combo < 3; ++combo) {
          const CUtensorMap* tmS = combo == 0 ? (src ? &tmElo : &tmXlo) : (src ? &tmE : &tmX);
          const int hcol = cb * 32 + (combo == 1 ? p.lo_off : 0);
          for (long long s = s0; s < s1; ++s) {
            wait_uniform(ab, &empty[ps.stage], ps.phase ^ 1);
            uint8_t* dst = St + (size_t)ps.stage * stage_bytes;
            const int tau0 = (int)(s * 32);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full[ps.stage], stage_bytes);
              if (!p.quad) {
                tma_load_3d(dst, tmS, &full[ps.stage], 0, tau0, nt * 4);    // four 32-feature regions in one box
              } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)                                 // the same features, 8 r lags later
                  tma_load_2d(dst + r * 4096, tmS, &full[ps.stage], 0, tau0 + r * 8 * p.s);
              }
              // Hv rows tau0 - s*(l0 + 7) .. tau0 + 32; row index in Hv is tau + h
              tma_load_2d(dst + kWtABytes, &tmH, &full[ps.stage], hcol, tau0 - p.s * (lg * lpi + 7) + p.h);
            }
            ps.advance(kWtStages);
          }
        }
      }
    }
  } else if (warp == 1) {
    reg_dec<kSCtlRegs>();
    {
      const uint32_t idesc = make_idesc_tf32(128, 256, 1, 1);
      const uint64_t adesc0 = make_smem_desc(smem_u32(St), 4096, 512, 1 /*SW128_BASE32B*/);
      // N-atom a starts a * s rows further down: lag 8 lg + 7 - a
      const uint64_t bdesc0 = make_smem_desc(smem_u32(St) + kWtABytes, (uint32_t)p.s * 128, 512, 1);
      const uint32_t stage16 = stage_bytes >> 4;
      PipeState ps;
      SubIssue si;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int lg, cb, src, nt, ch;
        decode(item, lg, cb, src, nt, ch);
        long long s0, s1;
        chunk_range(ch, s0, s1);
        const int n_units = (int)(n_pass * (s1 - s0));
        const SubPlan plan(n_units, p.x3 ? 2 * (s1 - s0) : 0, p.sub_units);
        for (int u = 0; u < n_units; ++u) {
          wait_uniform(ab, &full[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t dtm = si.begin_unit(ab, sempty, tmem);
          const uint64_t ad = adesc0 + (uint64_t)(ps.stage * stage16);
          const uint64_t bd = bdesc0 + (uint64_t)(ps.stage * stage16);
          if (elect_one()) {
            mma_tf32_ss(dtm, ad, bd, idesc, si.us != 0 ? 1u : 0u);      // k-step ks: A +1 KB, B +8 rows
            mma_tf32_ss(dtm, ad + 64, bd + 64, idesc, 1u);
            mma_tf32_ss(dtm, ad + 128, bd + 128, idesc, 1u);
            mma_tf32_ss(dtm, ad + 192, bd + 192, idesc, 1u);
          }
          si.end_unit(sfull, plan, u + 1);
          if (elect_one()) mma_commit(&empty[ps.stage]);
          ps.advance(kWtStages);
        }
      }
    }
  } else if (warp < 4) {
    reg_dec<kSCtlRegs>();
  } else {
    reg_inc<kSEpiRegs>();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    float m[128];
    SubDrain sd;
    bool ok = true;
    for (long long item = blockIdx.x; item < p.n_items && ok; item += gridDim.x) {
      int lg, cb, src, nt, ch;
      decode(item, lg, cb, src, nt, ch);
      long long s0, s1;
      chunk_range(ch, s0, s1);
      const SubPlan plan(n_pass * (s1 - s0), p.x3 ? 2 * (s1 - s0) : 0, p.sub_units);
      const int n_sub = plan.n_sub();
      for (int sub = 0; sub < n_sub; ++sub, sd.next()) {
        if (!wait_relaxed(ab, &sfull[sd.b], sd.ph)) { ok = false; break; }
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sd.b * 256 + half * 128);
        if (sub == 0) fold_sub<true>(m, taddr); else fold_sub<false>(m, taddr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[sd.b]);
      }
      if (!ok) break;
      const int n = p.quad ? lane : nt * 128 + q * 32 + lane;
      float* obase = p.part + ((long long)ch * p.n_src + src) * p.per_src;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {              // 4 column blocks of 32 = one virtual lag each
        const int a = half * 4 + c;
        const int lv = lg * lpi + (p.quad ? 8 * q : 0) + 7 - a;
        if (n < p.Np && lv < p.Lv) {
          if (p.s == 1) {                         // 32 components of column block cb, real lag lv
            float4* o = reinterpret_cast<float4*>(obase + ((long long)lv * p.Np + n) * p.Kp + cb * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]);
          } else {                                // Kp = 16 / 8: s real lags of Kp components each
            const int kq = p.Kp / 4;              // float4 per real lag
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int dl = j / kq, l = p.s * lv + dl;
              if (l < p.L)
                reinterpret_cast<float4*>(obase + ((long long)l * p.Np + n) * p.Kp)[j - dl * kq] =
                    make_float4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]);
            }
          }
        }
        rotate32(m);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ==========================================================================
// K3  H terms (every tensor-core precision mode)
//   out[src][tau][k] = sum_l sum_n W[l][n][k] * S^T[tau+l][n],  S = X | est
//   (reference tensor_transconv, cmfpy/common.py:61-86, via mult.py:42-48)
// The output has only K rows, so the lag groups share the 128 MMA rows:
//   row (g,k) of D accumulates lags l = j + J*g (j = 0..J-1):
//   D[(g,k)][c] += W[j+J*g][n][k] * S^T[base + c + j][n]      (M=128, N=256)
//   => D[(g,k)][c] is the group-g part of out[k][base + c - s*J*g].
//   A = W rows of the lag groups, MN-major SWIZZLE_128B_BASE32B (4 regions = (lag group, column block))
//   B = S^T window, K-major SWIZZLE_128B rows of 32 features, row shift j = +128 B
// Work item = (time tile of 256 columns, source); a unit is one lag of one 32-feature chunk.  The eight epilogue
// warps hold the item's accumulator in registers and reduce the lag groups INSIDE the CTA: a shared-memory tile
// R[u][k], u = tau - base + hd, is zeroed and the warps add their rows at their group's shift one after the other
// (fixed order: deterministic, no atomics).  R covers hd = (groups-1) s J + s - 1 columns before the tile: the
// complete part goes straight to out, the head (which the previous tile's groups also feed) to a small carry
// buffer that hterms_carry_kernel adds afterwards - 6 % of the output at config C instead of the 16-fold
// partial round trip of a separate combine pass.
// ==========================================================================
struct HTermsParams {
  int Np, J, n_chunks_n, wrows;    // n chunks of 32 features; window rows >= 256 + s*(J-1)
  int s, CB, Kp;                   // lag stride in rows; column blocks (regions = (4/CB lag groups) x CB); padded K
  int n_src;                       // 2: X and est (numerator, denominator); 1: X only (denominator via Gram)
  long long n_time_tiles;          // TO / 256 + 1 (the last tile only feeds the carry of the final columns)
  long long n_items;               // n_time_tiles * n_src * n_split
  int n_split;                     // > 1: the feature chunks of a (time tile, source) are shared out over n_split items
                                   // (short shards, where whole tiles leave the last round of CTAs half empty); n_chunks_n
                                   // is a multiple of it; the items write partial outputs that the host sums
  long long TO;                    // output rows per source
  float* out;                      // [n_src][TO][Kp]  (n_split > 1: [n_src][n_split][TO][Kp] partials)
  float* carry;                    // [n_src][n_split][n_time_tiles][hd][Kp]
  int hd;                          // columns before the tile that its lag groups reach: (4/CB - 1) s J + s - 1
  int sub_units;                   // units per sub-chunk (0: the whole item in one tensor-memory chain)
  int n_stages;                    // W ring depth (2 lags per stage)
  int staged;                      // folded lags: reduce through a staged gather (needs hterms_stage_bytes of shared memory)
  int x3, lo_off;                  // 3xTF32: the feature chunks of an item are walked three times - (W lo, S hi),
                                   // (W hi, S lo), (W hi, S hi)
  int* err;
};

constexpr int kHtLagsPerStage = 2;
constexpr int kHtABytes = 4 * 32 * 128;            // 4 regions x 32 n x 32 k
constexpr int kHtStageBytes = kHtLagsPerStage * kHtABytes;
constexpr int kHtMaxStages = 4;

constexpr int kHtStageLd = 33;                     // staging row stride in floats (odd: conflict-free both ways)
__host__ __device__ inline size_t hterms_stage_bytes(bool staged) {
  return staged ? (size_t)2 * 128 * kHtStageLd * 4 : 0;       // folded lags: [column half][128 rows][32 columns]
}
__host__ __device__ inline size_t hterms_r_bytes(int Kp, int hd, bool direct, bool staged) {
  // (staged: R is component-major with an odd row length, at most 256 + hd + 1)
  return direct ? 0 : (size_t)(256 + hd + (staged ? 1 : 0)) * Kp * 4 + hterms_stage_bytes(staged);
}
__host__ __device__ inline size_t hterms_smem_bytes(int n_stages, int wrows, int Kp, int hd, bool direct, bool staged) {
  return 1024 + (size_t)n_stages * kHtStageBytes + 2 * (size_t)wrows * 128 + hterms_r_bytes(Kp, hd, direct, staged) + 256;
}

__global__ void __launch_bounds__(kSThreads, 1)
tc_hterms_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmXlo,
                 const __grid_constant__ CUtensorMap tmElo, const HTermsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int n_glag = 4 / p.CB;
  const bool direct = (n_glag == 1 && p.s == 1);                  // nothing to reduce: rows are whole outputs
  const int n_stages = p.n_stages;
  uint8_t* As = smem;                                             // [stages][2 lags x 16 KB]
  const uint32_t wbytes = (uint32_t)p.wrows * 128;                // one 32-feature chunk of the window
  uint8_t* Ws = As + (size_t)n_stages * kHtStageBytes;            // [2][wbytes]
  float* R = (float*)(Ws + 2 * (size_t)wbytes);                   // [256 + hd][Kp]
  uint64_t* bars = (uint64_t*)((uint8_t*)R + hterms_r_bytes(p.Kp, p.hd, direct, p.staged != 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + kHtMaxStages;
  uint64_t* wfull = bars + 2 * kHtMaxStages;                      // [2]
  uint64_t* wempty = wfull + 2;                                   // [2]
  uint64_t* sfull = wempty + 2;                                   // [2]
  uint64_t* sempty = sfull + 2;                                   // [2]
  uint32_t* tmem_slot = (uint32_t*)(sempty + 2);
  volatile int* abort_flag = (volatile int*)(tmem_slot + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler, too
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1);
      mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 8);
    }
    *abort_flag = 0;
    fence_mbar_init();
    prefetch_tmap(&tmW); prefetch_tmap(&tmX); prefetch_tmap(&tmE); prefetch_tmap(&tmXlo); prefetch_tmap(&tmElo);
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const Abort ab{abort_flag, p.err};
  const int J = p.J, wrows = p.wrows;
  const int n_pass = p.x3 ? 3 : 1;
  const int n_split = p.n_split > 1 ? p.n_split : 1;
  const int cpi = p.n_chunks_n / n_split;                        // feature chunks per item
  const int per_tile = p.n_src * n_split;                        // item = (time tile, source, split)
  const SubPlan plan((long long)n_pass * cpi * J, p.x3 ? 2ll * cpi * J : 0, p.sub_units);   // units per item

  if (warp == 0) {
    reg_dec<kSCtlRegs>();
    {
      PipeState ps;
      // chunks = (item, operand pass, 32-feature chunk) in execution order; chunk c uses window buffer c & 1
      struct Chunk { long long item; int nc, combo; bool valid; };
      auto make_chunk = [&](long long item) { return Chunk{item, 0, p.x3 ? 0 : 2, item < p.n_items}; };
      auto next_chunk = [&](Chunk c) {
        if (++c.nc >= cpi) {
          if (c.combo < 2) { ++c.combo; c.nc = 0; }
          else c = make_chunk(c.item + gridDim.x);
        }
        return c;
      };
      // window of chunk number wc (its buffer is known to be free)
      auto issue_window = [&](const Chunk& c, long long wc) {
        const int wb = (int)(wc & 1);
        const long long tile = c.item / per_tile;
        const int rem = (int)(c.item % per_tile);
        const int src = rem / n_split, nc0 = (rem % n_split) * cpi;
        uint8_t* wdst = Ws + (size_t)wb * wbytes;
        const CUtensorMap* tmS = src ? (c.combo == 1 ? &tmElo : &tmE) : (c.combo == 1 ? &tmXlo : &tmX);
        if (elect_one()) {
          mbar_arrive_expect_tx(&wfull[wb], wbytes);
          for (int rb = 0; rb < wrows / 32; ++rb)
            tma_load_2d(wdst + (size_t)rb * 32 * 128, tmS, &wfull[wb], (nc0 + c.nc) * 32, (int)(tile * 256 + rb * 32));
        }
      };
      Chunk cur = make_chunk(blockIdx.x);
      long long wc = 0;
      if (cur.valid) issue_window(cur, 0);
      while (cur.valid) {
        const Chunk nxt = next_chunk(cur);
        bool prefetched = !nxt.valid;
        const int nwb = (int)((wc + 1) & 1);
        const uint32_t npar = (uint32_t)(((wc + 1) >> 1) & 1) ^ 1;
        for (int j = 0; j < J; j += kHtLagsPerStage) {
          wait_uniform(ab, &empty[ps.stage], ps.phase ^ 1);
          // one box: lags j, j+1 x regions (lag group, column block) x 32 features x 32 components (lags >= J
          // and features >= Np arrive as zeros)
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[ps.stage], kHtStageBytes);
            tma_load_5d(As + (size_t)ps.stage * kHtStageBytes, &tmW, &full[ps.stage], 0,
                        ((int)(cur.item % n_split) * cpi + cur.nc) * 32,
                        cur.combo == 0 ? p.lo_off / 32 : 0, 0, j);
          }
          ps.advance(n_stages);
          // the next window goes out as soon as its buffer is free; waiting for it here would starve the W ring
          if (!prefetched && __shfl_sync(0xffffffffu, (int)mbar_test_wait(&wempty[nwb], npar), 0)) {
            issue_window(nxt, wc + 1);
            prefetched = true;
          }
        }
        if (!prefetched) {
          wait_uniform(ab, &wempty[nwb], npar);
          issue_window(nxt, wc + 1);
        }
        cur = nxt;
        ++wc;
      }
    }
  } else if (warp == 1) {
    reg_dec<kSCtlRegs>();
    {
      const uint32_t idesc = make_idesc_tf32(128, 256, 1, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(As), 4096, 512, 1 /*SW128_BASE32B*/);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(Ws), 16, 1024, kSwz128);
      PipeState ps;
      SubIssue si;
      uint32_t wb = 0, wph = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int unit = 0;
        for (int chunk = 0; chunk < n_pass * cpi; ++chunk) {
          wait_uniform(ab, &wfull[wb], wph);
          tc_fence_after();
          uint64_t bd = bdesc0 + (uint64_t)((wb * wbytes) >> 4);             // lag j: row shift s * j
          for (int j = 0; j < J; j += kHtLagsPerStage) {
            wait_uniform(ab, &full[ps.stage], ps.phase);
            tc_fence_after();
            const int nl = min(kHtLagsPerStage, J - j);
            uint64_t ad = adesc0 + (uint64_t)(ps.stage * (kHtStageBytes >> 4));
            for (int u = 0; u < nl; ++u) {
              const uint32_t dtm = si.begin_unit(ab, sempty, tmem);
              if (elect_one()) {
                mma_tf32_ss(dtm, ad, bd, idesc, si.us != 0 ? 1u : 0u);  // k-step ks: A +1 KB, B +32 B
                mma_tf32_ss(dtm, ad + 64, bd + 2, idesc, 1u);
                mma_tf32_ss(dtm, ad + 128, bd + 4, idesc, 1u);
                mma_tf32_ss(dtm, ad + 192, bd + 6, idesc, 1u);
              }
              si.end_unit(sfull, plan, ++unit);
              ad += kHtABytes >> 4;
              bd += (uint64_t)(p.s * 8);
            }
            if (elect_one()) mma_commit(&empty[ps.stage]);
            ps.advance(n_stages);
          }
          if (elect_one()) mma_commit(&wempty[wb]);
          wb ^= 1;
          wph ^= (wb == 0);
        }
      }
    }
  } else if (warp < 4) {
    reg_dec<kSCtlRegs>();
  } else {
    reg_inc<kSEpiRegs>();
    const int e = warp - 4;                   // epilogue warp 0..7
    const int q = warp & 3;                   // region (lag group, column block) of this warp's 32 TMEM lanes
    const int half = e >> 2;
    const int etid = tid - 128;               // 0..255
    const int n_sub = plan.n_sub();
    const int gl = q / p.CB;                  // lag group
    const int dl = p.s > 1 ? lane / p.Kp : 0; // real lag inside a folded virtual lag
    const int kcol = p.s > 1 ? lane % p.Kp : (q % p.CB) * 32 + lane;
    const int u0 = half * 128 + (n_glag - 1 - gl) * p.s * J + (p.s - 1 - dl);     // R row of this thread's column 0
    const int U = 256 + p.hd;
    float m[128];
    SubDrain sd;
    bool ok = true;
    for (long long item = blockIdx.x; item < p.n_items && ok; item += gridDim.x) {
      const long long tile = item / per_tile;
      const int src = (int)(item % per_tile);       // source * n_split + split: the index of this item's (partial) output
      for (int sub = 0; sub < n_sub; ++sub, sd.next()) {
        if (!wait_relaxed(ab, &sfull[sd.b], sd.ph)) { ok = false; break; }
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sd.b * 256 + half * 128);
        if (sub == 0) fold_sub<true>(m, taddr); else fold_sub<false>(m, taddr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[sd.b]);
      }
      if (!ok) break;
      float* outs = p.out + (size_t)src * p.TO * p.Kp;
      const long long base = tile * 256;
      if (direct) {
        // Kp = 128: row (cb, k) at column c is out[base + c][cb*32 + k] itself
        float* o = outs + (size_t)(base + half * 128) * p.Kp + kcol;
        const long long left = p.TO - (base + half * 128);
        const int rows = left > 128 ? 128 : (int)left;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c * 32 + i < rows) *o = m[i];
            o += p.Kp;
          }
          rotate32(m);
        }
        continue;
      }
      // ---- reduce the lag groups through R ----
      const int Up = U | 1;                        // staged: R is [Kp][Up] (component-major, odd row length)
      {
        float4* R4 = reinterpret_cast<float4*>(R);
        const int n4 = (p.staged ? Up : U) * p.Kp / 4;
        for (int i = etid; i < n4; i += kSEpiThreads) R4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      epi_bar();
      if (p.staged) {
        // Folded lags (K < 32): the 128 accumulator rows are (lag group, lag inside a virtual lag, k) and every one of
        // them lands on its own row offset (`shift`) of R.  All warps stage 32 accumulator columns at a time, row-major;
        // then warp e owns components e, e + 8, ... and lane r the R rows u = r (mod 32): staged column j of source row
        // (g, d) belongs to u = 32 c + j + shift, so for every source the lane reads column (r - shift) mod 32 and adds
        // it to the row 32 c + r + 32 ceil((shift - r) / 32) - two real adds per source (one per column half), no
        // predicates, staging reads and R accesses conflict-free, one owner per R element (fixed order: deterministic,
        // no atomics).  Its predecessors - the warps adding their rows one after the other, then a gather over (u, k)
        // with sixteen predicated sources per output - made this reduction, not the MMAs, the length of a work item on
        // small problems (config B: 44 k cycles per item against 25 k of MMAs).
        float* St = R + (size_t)Up * p.Kp;                        // [2][128][kHtStageLd]
        float* mine = St + ((size_t)half * 128 + q * 32 + lane) * kHtStageLd;
        const int sJ = p.s * J;
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mine[i] = m[i];
          rotate32(m);
          epi_bar();
          for (int k = e; k < p.Kp; k += 8) {
            float* rk = R + (size_t)k * Up + c * 32 + lane;
            const float* srow_g = St + (size_t)k * kHtStageLd;
            int shift_g = (n_glag - 1) * sJ + (p.s - 1);         // shift of source (g = 0, d = 0)
            for (int g = 0; g < n_glag; ++g, shift_g -= sJ, srow_g += 32 * kHtStageLd) {
              const float* srow = srow_g;
              int shift = shift_g;
              for (int d = 0; d < p.s; ++d, --shift, srow += p.Kp * kHtStageLd) {
                const int j0 = (lane - shift) & 31;
                float* ru = rk + ((shift - lane + 31) & ~31);
                ru[0] += srow[j0];
                ru[128] += srow[128 * kHtStageLd + j0];
              }
            }
          }
          epi_bar();
        }
      } else {
      for (int w = 0; w < 8; ++w) {
        if (w == e) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            float* r = R + (size_t)(u0 + c * 32) * p.Kp + kcol;
            for (int d = 0; d < p.s; ++d) {        // folded lags of one warp overlap in R: one at a time
              if (d == dl) {
                // the 32 rows are distinct addresses, which the compiler cannot know (Kp is a run-time stride): load
                // them all, then store them all, or every += waits for the store before it (~45 cycles each)
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = r[i * p.Kp];
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i * p.Kp] = t[i] + m[i];
              }
              __syncwarp();
            }
            rotate32(m);
          }
        }
        epi_bar();
      }
      }
      // ---- R -> carry (the hd columns before the tile) and out (the tile's own 256 columns) ----
      {
        const float4* R4 = reinterpret_cast<const float4*>(R);
        const int kp4 = p.Kp / 4;
        const int head4 = p.hd * kp4;
        float4* c4 = reinterpret_cast<float4*>(p.carry + ((size_t)src * p.n_time_tiles + tile) * p.hd * p.Kp);
        float4* o4 = reinterpret_cast<float4*>(outs + (size_t)base * p.Kp);
        long long rows_left = p.TO - base;
        if (rows_left > 256) rows_left = 256;
        const int main4 = rows_left > 0 ? (int)rows_left * kp4 : 0;
        if (p.staged) {
          // component-major R: element (u, 4 k4 .. 4 k4 + 3) -> one float4 of the row-major output (Kp is 8 or 16)
          const int sh = kp4 == 2 ? 1 : 2;
          for (int i = etid; i < head4 + main4; i += kSEpiThreads) {
            const int u = i >> sh, k0 = (i & (kp4 - 1)) * 4;
            const float* r = R + (size_t)k0 * Up + u;
            const float4 v = make_float4(r[0], r[Up], r[2 * Up], r[3 * Up]);
            if (i < head4) c4[i] = v; else o4[i - head4] = v;
          }
        } else {
          for (int i = etid; i < head4; i += kSEpiThreads) c4[i] = R4[i];
          for (int i = etid; i < main4; i += kSEpiThreads) o4[i] = R4[head4 + i];
        }
      }
      epi_bar();                                   // R is zeroed again by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// out[src][t][:] += sum over the tiles i > t / 256 whose head reaches t of carry[src][i][t - (256 i - hd)][:]
// (ascending i: deterministic).  One thread per float4 of the min(hd, 256) columns before each tile boundary.
__global__ void __launch_bounds__(256)
hterms_carry_kernel(float* __restrict__ out, const float* __restrict__ carry, long long TO, long long n_time_tiles,
                    int hd, int Kp, int n_src) {
  const int kp4 = Kp / 4;
  const int wmax = hd < 256 ? hd : 256;
  const long long per_src = (n_time_tiles - 1) * wmax * kp4;
  const long long total = per_src * n_src;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int src = (int)(idx / per_src);
    long long r = idx % per_src;
    const int k4 = (int)(r % kp4); r /= kp4;
    const int w = (int)(r % wmax) + 1;                 // distance to the next tile boundary
    const long long i0 = r / wmax + 1;                 // that tile
    const long long t = i0 * 256 - w;
    if (t >= TO) continue;
    float4* o = reinterpret_cast<float4*>(out + ((size_t)src * TO + t) * Kp) + k4;
    float4 acc = *o;
    for (long long i = i0; i < n_time_tiles && i * 256 - t <= hd; ++i) {
      const long long u = t - (i * 256 - hd);
      const float4 v = __ldcs(reinterpret_cast<const float4*>(carry + (((size_t)src * n_time_tiles + i) * hd + u) * Kp) + k4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *o = acc;
  }
}

}  // namespace tc
}  // namespace cmf
