// physs_rt_sum_d32.cu -- chunk-summary instantiations of physs_rt_sum_impl.cuh for the padded dimension 32
#include "physs_rt_sum_impl.cuh"

namespace physs {
PHYSS_RT_SUM_INSTANTIATE(32)
}  // namespace physs
