// Row-kernel instantiations: float data, bernoulli likelihood (see skk_row_inst.cuh).
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f32_bernoulli(int bt, int ng, int np, int ns) {
    return row_kernel_lookup<float, SKK_LIK_BERNOULLI_LOGIT>(bt, ng, np, ns);
}
}  // namespace skk
