// Row-kernel instantiations: double data, poisson likelihood (see skk_row_inst.cuh).
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f64_poisson(int bt, int ng, int np, int ns) {
    return row_kernel_lookup<double, SKK_LIK_POISSON_LOG>(bt, ng, np, ns);
}
}  // namespace skk
