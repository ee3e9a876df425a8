// Row-kernel instantiations: double data, gamma likelihood, batch tile 4 (see skk_row_inst.cuh).
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f64_gamma_bt4(int ng, int np, int ns, int spec) { return row_kernel_lookup<double, SKK_LIK_GAMMA_LOG, 4>(ng, np, ns, spec); }
}  // namespace skk
