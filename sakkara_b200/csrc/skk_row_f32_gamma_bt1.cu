// Row-kernel instantiations: float data, gamma likelihood, batch tile 1 (see skk_row_inst.cuh).
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f32_gamma_bt1(int ng, int np, int ns, int spec) { return row_kernel_lookup<float, SKK_LIK_GAMMA_LOG, 1>(ng, np, ns, spec); }
}  // namespace skk
