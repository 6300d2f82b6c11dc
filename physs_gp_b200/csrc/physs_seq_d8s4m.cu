// physs_seq_d8s4m.cu -- instantiations of the thread-per-series sequential filter/smoother for state dim 8
// (two Matern-7/2 blocks), closed-form Matern discretisation.  The 8 x 8 blocks exceed the register file:
// the compiler keeps the hot tiles in registers and places the rest in lane-interleaved local memory
// (L1-resident), which still beats the shared-memory lane-group path at this size.
#include "physs_seq_impl.cuh"
namespace physs {
int seq_filter_d8s4m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  return filter_by_m<8, 4, false>(st, a, m, hid);
}
int seq_smooth_d8s4m(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  return smooth_by_mo<8, 4, false>(st, a, mo);
}
}  // namespace physs
