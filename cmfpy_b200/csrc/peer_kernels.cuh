// Collectives of the time-sharded MU iteration over NVLink / NVSwitch peer
// memory (one process per GPU; buffers mapped into every peer with CUDA IPC).
//
//   W step   : every rank holds partial sums [num | den] over its own columns
//              (reference _compute_mult_W, cmfpy/algs/mult.py:27-40).  ONE kernel
//              per rank does reduce-scatter -> W * num / (den + eps) (mult.py:18)
//              -> all-gather: rank r reads slice r of every peer's partials over
//              NVLink, sums them in rank order, updates slice r of W and stores
//              it into EVERY rank's W.  Each element of W has one writer, so W is
//              bit-identical on all ranks by construction.
//   H halos  : after the H update (mult.py:22) each rank stores its first / last
//              L-1 columns of H into the neighbours' staging rows and waits for
//              theirs (cmf_predict needs H[:, t-l], common.py:50-58).
//   loss     : every rank stores its local sum of squared residuals into slot
//              [step][rank] of every peer's ring; the host adds the G values of
//              a step in rank order (base.py:90-97), so all ranks report the same
//              loss without a collective on the critical path.
//
// Synchronisation is by monotonically increasing epoch flags in each rank's own
// control block: a rank only ever SPINS ON ITS OWN MEMORY and only STORES to
// peers.  Every spin is bounded (~10 s) and reports through `err`; a dead peer
// cannot hang the GPU.
#pragma once
#include "common.cuh"

namespace cmf {
namespace peer {

constexpr int kMaxPeers = 8;
constexpr long long kSpinTimeoutCycles = 20000000000ll;   // ~10 s: ranks are only loosely in step
enum { kErrPeerTimeout = 2 };

// Control block at the head of each rank's shared allocation.
struct Control {
  uint32_t ready[kMaxPeers];     // ready[p] = e: peer p's W-term partials of exchange e are complete
  uint32_t done[kMaxPeers];      // done[p]  = e: peer p has stored its slice of W (exchange e) into my W
  uint32_t halo[2];              // halo[0] = e: left neighbour's columns arrived; halo[1]: right neighbour's
  uint32_t bar[kMaxPeers];       // end-of-call barrier
  uint32_t blocks_done;          // block counter of the running exchange kernel
  int err;
  // Epochs and the loss-ring slot live on the device and are advanced by tick_kernel, so that every kernel of a
  // sharded iteration has constant arguments and the iteration can be captured once and replayed as a CUDA graph.
  uint32_t ex_epoch, halo_epoch, slot, pad;
};
static_assert(sizeof(Control) % 16 == 0, "control block must keep the payload 16-byte aligned");

struct Peers {
  int rank, world;
  Control* ctl[kMaxPeers];       // ctl[rank] is local
  float* numden[kMaxPeers];      // [num | den] partials, 2 * wcount floats each
  float* W[kMaxPeers];           // W masters
  float* halo_in[kMaxPeers];     // staging: [0]: from the left neighbour, [1]: from the right; h * Kp floats each
  double* ring[kMaxPeers];       // [slot][kMaxPeers] local residual sums
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data is read with ld.volatile: never from a stale L1 line
__device__ __forceinline__ float4 ld_peer(const float4* p) { return __ldcv(p); }

// spin until *flag >= epoch (wrap-safe); false on timeout
// Once one wait has timed out (err != 0) every later wait on this rank gives up at once, so a broken exchange costs
// one timeout, not one per kernel; the host reads err at the end of the call and raises.
__device__ __forceinline__ bool wait_epoch(const uint32_t* flag, uint32_t epoch, int* err) {
  if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return true;
  if (*reinterpret_cast<volatile int*>(err) != 0) return false;
  const long long t0 = clock64();
  unsigned spins = 0;
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (clock64() - t0 > kSpinTimeoutCycles) {
      atomicExch(err, kErrPeerTimeout);
      return false;
    }
    if ((++spins & 1023u) == 0 && *reinterpret_cast<volatile int*>(err) != 0) return false;
    __nanosleep(64);
  }
  return true;
}

// advances the device-side counters of this rank (one thread): what the next exchange kernels will use
enum { kTickExchange = 1, kTickHalo = 2, kTickSlotReset = 4 };
__global__ void tick_kernel(Control* me, int what) {
  if (what & kTickExchange) me->ex_epoch += 1;
  if (what & kTickHalo) me->halo_epoch += 1;
  if (what & kTickSlotReset) me->slot = 0;
}

// --------------------------------------------------------------------------
// W step: reduce-scatter of the partials + multiplicative update + all-gather of W.
//   n4       : wcount / 4 (float4 elements of W); the den half of numden starts at float4 index n4
// The TF32 operand copies of W are rebuilt by each rank's own refresh kernel afterwards.
// Launch: grid <= number of SMs (all blocks co-resident: they spin), 256 threads.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wstep_exchange_kernel(const Peers P, long long n4) {
  Control* me = P.ctl[P.rank];
  const int G = P.world;
  const uint32_t epoch = me->ex_epoch;              // set by the tick kernel before this launch
  // (1) my partials are complete (previous kernels on this stream): tell every peer
  if (blockIdx.x == 0 && threadIdx.x < G) {
    __threadfence_system();
    st_release_sys(&P.ctl[threadIdx.x]->ready[P.rank], epoch);
  }
  // (2) wait until every peer's partials are complete (block-wide AND of the per-peer waits)
  int ok = 1;
  if (threadIdx.x < G) ok = wait_epoch(&me->ready[threadIdx.x], epoch, &me->err) ? 1 : 0;
  if (__syncthreads_and(ok)) {
    // (3) my slice: sum over ranks in rank order, update, store into every rank's W
    const long long s0 = n4 * P.rank / G, s1 = n4 * (P.rank + 1) / G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += stride) {
      float4 a[kMaxPeers], d[kMaxPeers];
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p) {
        if (p < G) {
          a[p] = ld_peer(reinterpret_cast<const float4*>(P.numden[p]) + i);
          d[p] = ld_peer(reinterpret_cast<const float4*>(P.numden[p]) + n4 + i);
        }
      }
      float4 num = a[0], den = d[0];
#pragma unroll
      for (int p = 1; p < kMaxPeers; ++p) {
        if (p < G) {
          num.x += a[p].x; num.y += a[p].y; num.z += a[p].z; num.w += a[p].w;
          den.x += d[p].x; den.y += d[p].y; den.z += d[p].z; den.w += d[p].w;
        }
      }
      float4 w = reinterpret_cast<const float4*>(P.W[P.rank])[i];
      w.x = w.x * num.x / (den.x + kEpsilon);
      w.y = w.y * num.y / (den.y + kEpsilon);
      w.z = w.z * num.z / (den.z + kEpsilon);
      w.w = w.w * num.w / (den.w + kEpsilon);
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < G) reinterpret_cast<float4*>(P.W[p])[i] = w;
    }
  }
  // (4) the last block to finish publishes "my slice is everywhere" and waits for the peers' slices
  __threadfence_system();
  __syncthreads();
  __shared__ int last_s;
  if (threadIdx.x == 0) last_s = (atomicAdd(&me->blocks_done, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (last_s) {
    if (threadIdx.x == 0) me->blocks_done = 0;
    if (threadIdx.x < G) {
      __threadfence_system();
      st_release_sys(&P.ctl[threadIdx.x]->done[P.rank], epoch);
      wait_epoch(&me->done[threadIdx.x], epoch, &me->err);
    }
  }
}

// --------------------------------------------------------------------------
// H halos.  Ht: local H^T master ((h + RT) x Kp); rows [h, h + Tloc) are owned.
//   to the left neighbour  : my first h owned rows  -> its staging slot 1 ("from the right")
//   to the right neighbour : my last h owned rows   -> its staging slot 0 ("from the left")
// then wait for the neighbours' rows and copy them from the staging slots into my halo rows
// (zeros at the global boundaries).  One block per direction.
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
halo_exchange_kernel(const Peers P, float* __restrict__ Ht, int h, int Kp, long long Tloc) {
  Control* me = P.ctl[P.rank];
  const uint32_t epoch = me->halo_epoch;            // set by the tick kernel before this launch
  const int dir = blockIdx.x;                       // 0: talk to the left neighbour, 1: to the right
  const int nb = dir == 0 ? P.rank - 1 : P.rank + 1;
  const long long n4 = (long long)h * Kp / 4;
  float4* halo_rows = reinterpret_cast<float4*>(dir == 0 ? Ht : Ht + (size_t)(h + Tloc) * Kp);
  if (nb < 0 || nb >= P.world) {                    // global boundary: the halo is zeros
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) halo_rows[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* src = reinterpret_cast<const float4*>(dir == 0 ? Ht + (size_t)h * Kp : Ht + (size_t)Tloc * Kp);
  float4* dst = reinterpret_cast<float4*>(P.halo_in[nb]) + (dir == 0 ? n4 : 0);
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  __shared__ int ok_s;
  if (threadIdx.x == 0) {
    st_release_sys(&P.ctl[nb]->halo[dir == 0 ? 1 : 0], epoch);
    ok_s = wait_epoch(&me->halo[dir], epoch, &me->err) ? 1 : 0;
  }
  __syncthreads();
  if (!ok_s) return;
  const float4* in = reinterpret_cast<const float4*>(P.halo_in[P.rank]) + (dir == 0 ? 0 : n4);
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) halo_rows[i] = ld_peer(in + i);
}

// every rank's ring[slot][me] = my local residual sum of squares; the slot is this rank's device counter
__global__ void sumsq_push_kernel(const Peers P, const double* __restrict__ sumsq) {
  Control* me = P.ctl[P.rank];
  const uint32_t slot = me->slot;
  if (threadIdx.x < P.world) {
    P.ring[threadIdx.x][(size_t)slot * kMaxPeers + P.rank] = *sumsq;
    __threadfence_system();
  }
  __syncwarp();
  if (threadIdx.x == 0) me->slot = slot + 1;
}

// end-of-call barrier: everything every peer stored before it (ring slots) is visible afterwards
__global__ void barrier_kernel(const Peers P, uint32_t epoch) {
  Control* me = P.ctl[P.rank];
  if (threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys(&P.ctl[threadIdx.x]->bar[P.rank], epoch);
    wait_epoch(&me->bar[threadIdx.x], epoch, &me->err);
  }
}

}  // namespace peer
}  // namespace cmf
