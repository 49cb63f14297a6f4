// Shared helpers for the cmf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace cmf {

// float64 machine epsilon, the constant the reference adds to the MU
// denominators (reference cmfpy/common.py:9).  Held in fp32 on the device.
constexpr float kEpsilon = 2.220446049250313e-16f;

inline int round_up(long long x, long long m) { return (int)(((x + m - 1) / m) * m); }
inline long long round_up_ll(long long x, long long m) { return ((x + m - 1) / m) * m; }
inline long long ceil_div_ll(long long x, long long m) { return (x + m - 1) / m; }

// thread-local error text behind cmf_last_error()
std::string& last_error();
void set_error(const char* fmt, ...);

#define CMF_CUDA(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ::cmf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                       __FILE__, __LINE__);                                     \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

#define CMF_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      ::cmf::set_error(__VA_ARGS__);         \
      return 2;                              \
    }                                        \
  } while (0)

#define CMF_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Round an fp32 value to the nearest TF32 (10 explicit mantissa bits), ties to
// even, and return it as fp32.  tcgen05 kind::tf32 ignores the low 13 mantissa
// bits of its operands; operands stored pre-rounded make that truncation exact.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  return __uint_as_float(u);
}

}  // namespace cmf
