// physs_rt_sum_d8.cu -- chunk-summary instantiations of physs_rt_sum_impl.cuh for the padded dimension 8
#include "physs_rt_sum_impl.cuh"

namespace physs {
PHYSS_RT_SUM_INSTANTIATE(8)
}  // namespace physs
