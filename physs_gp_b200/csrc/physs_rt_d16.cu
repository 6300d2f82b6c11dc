// physs_rt_d16.cu -- filter / smoother instantiations of physs_rt_impl.cuh for the padded dimension 16
#include "physs_rt_impl.cuh"

namespace physs {
PHYSS_RT_INSTANTIATE(16)
}  // namespace physs
