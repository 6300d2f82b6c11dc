// physs_seq_d2w.cu -- register-resident sequential filter / smoother for one integrated Wiener block of state
// dim 2 (IWP(q = 1), kernels/wiener.py:60-149), closed-form A_k, Q_k on chip (PHYSS_DISC_IWP).
#include "physs_seq_impl.cuh"
namespace physs {
int seq_filter_d2w(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  return filter_by_m<2, 2, 2>(st, a, m, hid);
}
int seq_smooth_d2w(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  return smooth_by_mo<2, 2, 2>(st, a, mo);
}
}  // namespace physs
