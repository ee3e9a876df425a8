// Chain-per-lane kernel instantiations (double), see skk_chain_kernel.cuh.
#include "skk_chain_kernel.cuh"

namespace skk {
const void *chain_kernel_f64(int family, int ng, int np) {
    if (family == SKK_LIK_NORMAL) return chain_kernel_lookup<double, SKK_LIK_NORMAL>(ng, np);
    if (family == SKK_LIK_BERNOULLI_LOGIT) return chain_kernel_lookup<double, SKK_LIK_BERNOULLI_LOGIT>(ng, np);
    if (family == SKK_LIK_POISSON_LOG) return chain_kernel_lookup<double, SKK_LIK_POISSON_LOG>(ng, np);
    if (family == SKK_LIK_NEGBINOMIAL_LOG) return chain_kernel_lookup<double, SKK_LIK_NEGBINOMIAL_LOG>(ng, np);
    if (family == SKK_LIK_GAMMA_LOG) return chain_kernel_lookup<double, SKK_LIK_GAMMA_LOG>(ng, np);
    return nullptr;
}
}  // namespace skk
