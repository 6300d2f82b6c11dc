// physs_rt2_d32.cu -- two-kernel smoother (physs_rt2_impl.cuh) instantiated for the padded dimension 32
#include "physs_rt2_impl.cuh"

namespace physs {
template int rt2_smooth_dm<32>(cudaStream_t, const SeqSmoothArgs&, double*, int64_t);
template int64_t rt2_workspace_doubles<32>(int64_t, int64_t);
}  // namespace physs
