// Process-wide cache of freed device blocks (host side only).
#pragma once
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace cmf {

// The solver's larger device buffers (X^T, est^T and their lo halves; the operand copies, partial-sum and output
// buffers of the tensor-core path) come from a small process-wide cache of freed blocks: cudaMalloc of a few GiB costs
// anywhere from 10 ms to 0.4 s on these virtualised boxes (the end-to-end figure of bench.py moved between 19 and
// 25 it/s with identical code for that reason alone; the ~30 smaller allocations of one solver add another 40 ms),
// and a process that fits model after model - a sweep over K or L, CMF.fit in a loop - asks for the same sizes again
// and again.  A block is reused only for the same device and exactly the same size; blocks below 1 MiB are not kept;
// at most CMF_CACHE_GB GiB stay cached (default 16, 0 disables); cmf_release_cached_memory() returns them to the
// driver.  (Buffers whose CUDA-IPC handles are exported to peer processes do not come from here.)
struct BigCache {
  struct Block { int dev; void* p; size_t bytes; };
  std::mutex mu;
  std::vector<Block> free_blocks;
  std::unordered_map<void*, std::pair<size_t, int>> live;      // blocks handed out by cached_malloc: size, device
  size_t cached = 0;
  size_t cap() {
    static const size_t c = [] { const char* e = getenv("CMF_CACHE_GB"); return (size_t)((e ? atof(e) : 16.0) * (1ull << 30)); }();
    return c;
  }
};
inline BigCache& big_cache() { static BigCache* c = new BigCache(); return *c; }     // (never destroyed: outlives the runtime)
constexpr size_t kCacheMinBytes = 1u << 20;

inline void release_cached_blocks() {
  BigCache& c = big_cache();
  std::vector<BigCache::Block> blocks;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    blocks.swap(c.free_blocks);
    c.cached = 0;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  for (auto& b : blocks) {
    cudaSetDevice(b.dev);
    cudaFree(b.p);
  }
  if (prev >= 0) cudaSetDevice(prev);
}

inline int big_alloc(void** p, size_t bytes, int dev) {
  {
    BigCache& c = big_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    for (size_t i = 0; i < c.free_blocks.size(); ++i)
      if (c.free_blocks[i].dev == dev && c.free_blocks[i].bytes == bytes) {
        *p = c.free_blocks[i].p;
        c.cached -= bytes;
        c.free_blocks.erase(c.free_blocks.begin() + (long)i);
        return 0;
      }
  }
  if (cudaMalloc(p, bytes) != cudaSuccess) {
    cudaGetLastError();
    release_cached_blocks();                          // make room and try once more
    CMF_CUDA(cudaMalloc(p, bytes));
  }
  return 0;
}
inline void big_free(void* p, size_t bytes, int dev) {
  if (!p) return;
  BigCache& c = big_cache();
  {
    std::lock_guard<std::mutex> lock(c.mu);
    if (bytes >= kCacheMinBytes && c.cached + bytes <= c.cap()) {
      c.free_blocks.push_back({dev, p, bytes});
      c.cached += bytes;
      return;
    }
  }
  cudaFree(p);
}

// cudaMalloc / cudaFree look-alikes on the current device: the size of a block is remembered, so the matching free
// needs only the pointer.  cached_free also takes pointers that came from plain cudaMalloc.  Blocks of 1 MiB and more
// are handed out ZEROED, fresh or reused alike (a reused block holds the numbers of the previous solver, and padding
// rows that a TMA window reads must never hold a NaN): a solver behaves the same whatever the cache holds.  The
// memset is finished when the call returns (~0.2 ms per GiB), so every stream sees it.
inline int cached_malloc(void** p, size_t bytes) {
  if (bytes < kCacheMinBytes) {
    CMF_CUDA(cudaMalloc(p, bytes > 0 ? bytes : 1));
    return 0;
  }
  int dev = 0;
  CMF_CUDA(cudaGetDevice(&dev));
  CMF_TRY(big_alloc(p, bytes, dev));
  CMF_CUDA(cudaMemsetAsync(*p, 0, bytes, cudaStreamPerThread));
  CMF_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
  BigCache& c = big_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  c.live[*p] = {bytes, dev};
  return 0;
}
inline void cached_free(void* p) {
  if (!p) return;
  size_t bytes = 0;
  int dev = 0;
  {
    BigCache& c = big_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live.find(p);
    if (it != c.live.end()) { bytes = it->second.first; dev = it->second.second; c.live.erase(it); }
  }
  if (bytes) big_free(p, bytes, dev); else cudaFree(p);
}

}  // namespace cmf
