// physs_rt2_d16.cu -- two-kernel smoother (physs_rt2_impl.cuh) instantiated for the padded dimension 16
#include "physs_rt2_impl.cuh"

namespace physs {
template int rt2_smooth_dm<16>(cudaStream_t, const SeqSmoothArgs&, double*, int64_t);
template int64_t rt2_workspace_doubles<16>(int64_t, int64_t);
}  // namespace physs
