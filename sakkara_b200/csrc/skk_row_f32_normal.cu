// Row-kernel instantiations: float data, normal likelihood (see skk_row_inst.cuh).
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f32_normal(int bt, int ng, int np, int ns) {
    return row_kernel_lookup<float, SKK_LIK_NORMAL>(bt, ng, np, ns);
}
}  // namespace skk
