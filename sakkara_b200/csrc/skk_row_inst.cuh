// Explicit instantiation table of the row kernel for one (floating type, likelihood family).
// Every translation unit skk_row_<type>_<family>_bt<B>.cu includes this header once; the lookup returns
// the kernel specialised for (#global terms, #primary terms, #secondary terms) at a fixed batch tile.
#pragma once

#include "skk_row_kernel.cuh"

namespace skk {

#ifndef SKK_ROWS_PER_LANE
#define SKK_ROWS_PER_LANE 17
#endif
constexpr int kRowsPerLane = SKK_ROWS_PER_LANE;  // R: odd => conflict-free lane-strided shared-memory reads
                                  // (R = 9 with three resident CTAs per SM measured slower on the crossed model:
                                  //  the per-tile work does not shrink with the tile, profiles/README.md)

// spec bits (see RowSpec): x columns of global / primary / secondary terms, y as uint8, 16-bit codes
constexpr int kSpecXG0 = 1, kSpecXP0 = 4, kSpecY8 = 64, kSpecCode16 = 256;

template <typename T, int FAM, int BT>
const void *row_kernel_lookup(int ng, int np, int ns, int spec) {
    // shapes of the BASELINE configurations, fully specialised
#define SKK_FIXED(NG, NP, NS, SPEC)                                      \
    if (ng == NG && np == NP && ns == NS && spec == (SPEC))              \
        return reinterpret_cast<const void *>(&skk_row_kernel<T, FAM, BT, NG, NP, NS, kRowsPerLane, (SPEC)>);
    if constexpr (FAM != SKK_LIK_NORMAL_LOGSIGMA && FAM != SKK_LIK_STUDENTT_LOGSIGMA) {  // (run-time kernels only)
    SKK_FIXED(1, 1, 0, kSpecXP0)                         // intercept + group slope * x (README / C2)
    SKK_FIXED(0, 2, 0, kSpecXP0)                         // nested: parent slope * x + child intercept (C3)
    }
    if constexpr (FAM != SKK_LIK_NORMAL && FAM != SKK_LIK_NORMAL_LOGSIGMA && FAM != SKK_LIK_STUDENTT_LOGSIGMA) {
        SKK_FIXED(1, 1, 0, kSpecXP0 | kSpecY8)
        SKK_FIXED(0, 2, 0, kSpecXP0 | kSpecY8)
        SKK_FIXED(1, 1, 1, kSpecXG0 | kSpecY8 | kSpecCode16)  // beta * x + a[A] + b[B] (C4)
        SKK_FIXED(1, 1, 1, kSpecXG0 | kSpecCode16)
    }
#undef SKK_FIXED
    // every other model: one generic kernel per term-count shape
#define SKK_CASE(NG, NP, NS)              \
    if (ng == NG && np == NP && ns == NS) \
        return reinterpret_cast<const void *>(&skk_row_kernel<T, FAM, BT, NG, NP, NS, kRowsPerLane, -1>);
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
