// physs_rt_d32.cu -- filter / smoother instantiations of physs_rt_impl.cuh for the padded dimension 32
#include "physs_rt_impl.cuh"

namespace physs {
PHYSS_RT_INSTANTIATE(32)
}  // namespace physs
