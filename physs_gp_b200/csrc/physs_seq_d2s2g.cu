// physs_seq_d2s2g.cu -- instantiations of the register-resident sequential filter/smoother for
// state dim 2, transition block size 2, caller-given A_k/Q_k.
#include "physs_seq_impl.cuh"
namespace physs {
int seq_filter_d2s2g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  return filter_by_m<2, 2, true>(st, a, m, hid);
}
int seq_smooth_d2s2g(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  return smooth_by_mo<2, 2, true>(st, a, mo);
}
int seq_filter_summary_d2s2g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems) {
  return filter_summary_by_m<2, 2, true>(st, a, m, hid, elems);
}
int seq_smooth_summary_d2s2g(cudaStream_t st, const SeqSmoothArgs& a, double* elems) {
  return launch_smooth_summary<2, 2, true>(st, a, elems);
}
}  // namespace physs
