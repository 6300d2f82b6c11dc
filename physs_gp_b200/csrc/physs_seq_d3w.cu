// physs_seq_d3w.cu -- register-resident sequential filter / smoother for one integrated Wiener block of state
// dim 3 (IWP(q = 2), kernels/wiener.py:60-149), closed-form A_k, Q_k on chip (PHYSS_DISC_IWP).
#include "physs_seq_impl.cuh"
namespace physs {
int seq_filter_d3w(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  return filter_by_m<3, 3, 2>(st, a, m, hid);
}
int seq_smooth_d3w(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  return smooth_by_mo<3, 3, 2>(st, a, mo);
}
}  // namespace physs
