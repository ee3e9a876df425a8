// Explicit instantiation table of the row kernel for one (floating type, likelihood family).
// Every translation unit skk_row_<type>_<family>_bt<B>.cu includes this header once; the lookup returns
// the kernel specialised for (#global terms, #primary terms, #secondary terms) at a fixed batch tile.
#pragma once

#include "skk_row_kernel.cuh"

namespace skk {

constexpr int kRowsPerLane = 17;  // R: odd => conflict-free lane-strided shared-memory reads

template <typename T, int FAM, int BT>
const void *row_kernel_lookup(int ng, int np, int ns) {
#define SKK_CASE(NG, NP, NS)              \
    if (ng == NG && np == NP && ns == NS) \
        return reinterpret_cast<const void *>(&skk_row_kernel<T, FAM, BT, NG, NP, NS, kRowsPerLane>);
#define SKK_CASES_NS(NG, NP) SKK_CASE(NG, NP, 0) SKK_CASE(NG, NP, 1)
#define SKK_CASES_NP(NG) SKK_CASES_NS(NG, 0) SKK_CASES_NS(NG, 1) SKK_CASES_NS(NG, 2)
    SKK_CASES_NP(0)
    SKK_CASES_NP(1)
    SKK_CASES_NP(2)
#undef SKK_CASES_NP
#undef SKK_CASES_NS
#undef SKK_CASE
    return nullptr;
}

}  // namespace skk
