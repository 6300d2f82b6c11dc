// physs_rt_d8.cu -- filter / smoother instantiations of physs_rt_impl.cuh for the padded dimension 8
#include "physs_rt_impl.cuh"

namespace physs {
PHYSS_RT_INSTANTIATE(8)
}  // namespace physs
