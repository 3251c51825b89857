// lpx_reg.cu — register-resident batched kernel (placeholder until the kernel lands; the
// dispatcher never selects it while reg_kernel_supports() is false).
#include "lpx_common.cuh"
#include "lpx_stream.hpp"

namespace lpx {

bool reg_kernel_supports(int, int, int, bool) { return false; }

int reg_launch_batched(int, int, int, int, const double*, const double*, const double*, const lpx_options&, int*, int*,
                       int*, double*, double*, double*, unsigned long long*, cudaStream_t) {
    set_error("LPX_KERNEL_CTA_REG: no register-resident kernel is built for this shape");
    return LPX_E_CAPACITY;
}

}  // namespace lpx
