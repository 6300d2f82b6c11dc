// physs_seq_d3s1m.cu -- instantiations of the register-resident sequential filter/smoother for
// state dim 3, transition block size 1, closed-form Matern discretisation.
#include "physs_seq_impl.cuh"
namespace physs {
int seq_filter_d3s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  return filter_by_m<3, 1, false>(st, a, m, hid);
}
int seq_smooth_d3s1m(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  return smooth_by_mo<3, 1, false>(st, a, mo);
}
int seq_filter_summary_d3s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems) {
  return filter_summary_by_m<3, 1, false>(st, a, m, hid, elems);
}
int seq_smooth_summary_d3s1m(cudaStream_t st, const SeqSmoothArgs& a, double* elems) {
  return launch_smooth_summary<3, 1, false>(st, a, elems);
}
}  // namespace physs
