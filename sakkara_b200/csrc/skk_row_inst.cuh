// Explicit instantiation table of the row kernel for one (floating type, likelihood family).
// Every translation unit skk_row_<type>_<family>.cu includes this header once; the lookup returns
// the kernel specialised for (batch tile, #global terms, #primary terms, #secondary terms).
#pragma once

#include "skk_row_kernel.cuh"

namespace skk {

constexpr int kRowsPerLane = 17;  // R: odd => conflict-free lane-strided shared-memory reads

template <typename T, int FAM>
const void *row_kernel_lookup(int bt, int ng, int np, int ns) {
#define SKK_CASE(BT, NG, NP, NS)                        \
    if (bt == BT && ng == NG && np == NP && ns == NS) \
        return reinterpret_cast<const void *>(&skk_row_kernel<T, FAM, BT, NG, NP, NS, kRowsPerLane>);
#define SKK_CASES_NS(BT, NG, NP) SKK_CASE(BT, NG, NP, 0) SKK_CASE(BT, NG, NP, 1)
#define SKK_CASES_NP(BT, NG) SKK_CASES_NS(BT, NG, 0) SKK_CASES_NS(BT, NG, 1) SKK_CASES_NS(BT, NG, 2)
#define SKK_CASES_NG(BT) SKK_CASES_NP(BT, 0) SKK_CASES_NP(BT, 1) SKK_CASES_NP(BT, 2)
    SKK_CASES_NG(1)
    SKK_CASES_NG(4)
#undef SKK_CASES_NG
#undef SKK_CASES_NP
#undef SKK_CASES_NS
#undef SKK_CASE
    return nullptr;
}

}  // namespace skk
