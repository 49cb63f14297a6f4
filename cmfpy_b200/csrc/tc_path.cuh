// tcgen05 / TMEM / TMA contraction path (CMF_PREC_TF32) - interface.
#pragma once
#include "common.cuh"

namespace cmf {
namespace tc {

struct Dims {
  int N, K, L, Np, Kp, h;
  long long Tloc, TO, RT, RH, t_valid;
  int num_sms;
};

struct TcState {
  bool ready = false;
};

constexpr int kReconLaunches = 0, kWTermsLaunches = 0, kHTermsLaunches = 0;

inline bool shape_supported(int, int, int) { return false; }
inline int init(TcState&, const Dims&, float*, float*, float*, float*, float*, float*, double*, long long, double*,
                cudaStream_t) {
  set_error("tcgen05 path not built");
  return 2;
}
inline void destroy(TcState&) {}
inline int recon(TcState&, cudaStream_t) { return 2; }
inline int w_terms(TcState&, cudaStream_t) { return 2; }
inline int h_terms(TcState&, cudaStream_t) { return 2; }

}  // namespace tc
}  // namespace cmf
