// physs_rt_sum_d16.cu -- chunk-summary instantiations of physs_rt_sum_impl.cuh for the padded dimension 16
#include "physs_rt_sum_impl.cuh"

namespace physs {
PHYSS_RT_SUM_INSTANTIATE(16)
}  // namespace physs
