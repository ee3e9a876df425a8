// K2: chain-per-lane kernels for large batches of parameter vectors (BASELINE config 5: 1,024
// chains over the same data).  ALU-bound, not HBM-bound: the rows of a CTA's range are staged once in
// shared memory (TMA bulk copies, double buffered) and broadcast to all lanes; lane = chain, so
// every lane accumulates its own sums over the rows serially, in row order, with no cross-lane
// traffic at all -- trivially deterministic.  On a group boundary (warp-uniform: all lanes see the
// same row) the lane's primary accumulators are flushed: groups that lie inside the CTA's row range
// go straight to direct[group][chain] (coalesced over chains), the first / last group of the range
// to head / tail slots per CTA, which the slot kernel adds in CTA order.
// Supports global and primary-level terms (no secondary level).
#pragma once

#include "skk_aux_kernels.cuh"
#include "skk_common.cuh"

namespace skk {

constexpr int kChainTileRows = 512;   // rows per staged tile
constexpr int kChainWarps = 4;        // chain groups (of 32) per CTA, all sharing the staged rows

template <typename T>
struct ChainParams {
    const T *x[kMaxXCols];
    const T *y;
    const uint8_t *y8;
    const T *eta_offset;
    int32_t n_x;
    int64_t n_rows;
    int32_t n_tiles;             // ceil(n_rows / kChainTileRows)
    const int2 *range_keys;      // [gridDim.x] {first, last} primary slot of every CTA's row range
    const int64_t *seg_offsets;  // [n_primary + 1]
    int32_t n_primary;
    int32_t xcol_g[kMaxTermsPerLevel];
    int32_t xcol_p[kMaxTermsPerLevel];
    const int32_t *theta_index_g[kMaxTermsPerLevel];  // [1]
    const int32_t *theta_index_p[kMaxTermsPerLevel];  // [n_primary]
    const T *theta;              // [batch][P]
    T fam_a, fam_b;              // constants of the likelihood family (NegativeBinomial: alpha, log(alpha))
    int64_t n_params;
    int32_t batch;
    int32_t bp;                  // batch rounded up to a multiple of 32: chain stride of all outputs
    double *direct;              // [NP][n_primary][bp]   (zeroed per evaluation)
    double *head;                // [ranges][NP][bp]
    double *tail;                // [ranges][NP][bp]
    double *glob;                // [ranges][NG + 1][bp]
};

template <typename T, int FAM, int NG, int NP>
__global__ void __launch_bounds__(kChainWarps * 32) skk_chain_kernel(const __grid_constant__ ChainParams<T> p) {
    constexpr int NGa = NG > 0 ? NG : 1, NPa = NP > 0 ? NP : 1;
    constexpr int TR = kChainTileRows;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chain = (blockIdx.y * kChainWarps + warp) * 32 + lane;
    const bool live = chain < p.batch;
    const T *th = p.theta + (int64_t)(live ? chain : p.batch - 1) * p.n_params;

    const bool y_u8 = p.y8 != nullptr, has_off = p.eta_offset != nullptr;
    const int x_stride = TR * (int)sizeof(T);
    const int y_off = p.n_x * x_stride;
    const int off_off = y_off + TR * (y_u8 ? 1 : (int)sizeof(T));
    const int stage_bytes = (off_off + (has_off ? x_stride : 0) + 127) & ~127;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    unsigned char *stage0 = smem + 128;

    const int t0 = (int)(((int64_t)p.n_tiles * blockIdx.x) / gridDim.x);
    const int t1 = (int)(((int64_t)p.n_tiles * (blockIdx.x + 1)) / gridDim.x);
    const int64_t r0 = (int64_t)t0 * TR, r1 = min(p.n_rows, (int64_t)t1 * TR);

    double accG64[NG + 1], headv[NPa], tailv[NPa];
#pragma unroll
    for (int t = 0; t <= NG; ++t) accG64[t] = 0.0;
#pragma unroll
    for (int t = 0; t < NPa; ++t) headv[t] = tailv[t] = 0.0;

    auto issue = [&](int tile, int s) {
        const int64_t row = (int64_t)tile * TR;
        unsigned char *dst = stage0 + (int64_t)s * stage_bytes;
        uint32_t total = (uint32_t)p.n_x * x_stride + (y_u8 ? TR : x_stride) + (has_off ? x_stride : 0);
        mbar_arrive_expect_tx(bars + s, total);
        for (int c = 0; c < p.n_x; ++c) tma_load_1d(dst + c * x_stride, p.x[c] + row, x_stride, bars + s);
        if (y_u8)
            tma_load_1d(dst + y_off, p.y8 + row, TR, bars + s);
        else
            tma_load_1d(dst + y_off, p.y + row, x_stride, bars + s);
        if (has_off) tma_load_1d(dst + off_off, p.eta_offset + row, x_stride, bars + s);
    };

    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (t0 < t1) issue(t0, 0);
        if (t0 + 1 < t1) issue(t0 + 1, 1);
    }

    if (r0 < r1) {
        T vg[NGa], vp[NPa];
#pragma unroll
        for (int t = 0; t < NGa; ++t) vg[t] = NG > 0 ? th[p.theta_index_g[t][0]] : T(0);
        int g = p.range_keys[blockIdx.x].x;
        const int kf = g;
        int64_t nextb = NP > 0 ? p.seg_offsets[g + 1] : INT64_MAX;
        bool inside = NP > 0 ? p.seg_offsets[g] >= r0 : false;
#pragma unroll
        for (int t = 0; t < NPa; ++t) vp[t] = NP > 0 ? th[p.theta_index_p[t][g]] : T(0);
        double accP64[NPa];
#pragma unroll
        for (int t = 0; t < NPa; ++t) accP64[t] = 0.0;

        for (int tile = t0, it = 0; tile < t1; ++tile, ++it) {
            const int s = it & 1;
            mbar_wait(bars + s, (uint32_t)(it >> 1) & 1u);
            const unsigned char *st = stage0 + (int64_t)s * stage_bytes;
            const T *sy = reinterpret_cast<const T *>(st + y_off);
            const uint8_t *sy8 = st + y_off;
            const T *soff = reinterpret_cast<const T *>(st + off_off);
            const T *pxg[NGa], *pxp[NPa];
#pragma unroll
            for (int t = 0; t < NGa; ++t)
                pxg[t] = (t < NG && p.xcol_g[t] >= 0) ? reinterpret_cast<const T *>(st + p.xcol_g[t] * x_stride) : nullptr;
#pragma unroll
            for (int t = 0; t < NPa; ++t)
                pxp[t] = (t < NP && p.xcol_p[t] >= 0) ? reinterpret_cast<const T *>(st + p.xcol_p[t] * x_stride) : nullptr;

            const int64_t row_base = (int64_t)tile * TR;
            const int n = (int)min((int64_t)TR, r1 - row_base);
            int i = 0;
            while (i < n) {
                // rows of the current group inside this tile: a branch-free run
                const int run_end = (int)min((int64_t)n, nextb - row_base);
                T accL = T(0), accG[NGa], accP[NPa];
#pragma unroll
                for (int t = 0; t < NGa; ++t) accG[t] = T(0);
#pragma unroll
                for (int t = 0; t < NPa; ++t) accP[t] = T(0);
#pragma unroll 4
                for (; i < run_end; ++i) {
                    const T yv = y_u8 ? (T)sy8[i] : sy[i];
                    T xg[NGa], xp[NPa];
                    T eta = has_off ? soff[i] : T(0);
#pragma unroll
                    for (int t = 0; t < NG; ++t) {
                        xg[t] = pxg[t] ? pxg[t][i] : T(1);
                        eta += vg[t] * xg[t];
                    }
#pragma unroll
                    for (int t = 0; t < NP; ++t) {
                        xp[t] = pxp[t] ? pxp[t][i] : T(1);
                        eta += vp[t] * xp[t];
                    }
                    T lp, d;
                    family_row<T, FAM>(yv, eta, lp, d, FamilyConst<T>{p.fam_a, p.fam_b});
                    accL += lp;
#pragma unroll
                    for (int t = 0; t < NG; ++t) accG[t] += d * xg[t];
#pragma unroll
                    for (int t = 0; t < NP; ++t) accP[t] += d * xp[t];
                }
                accG64[NG] += (double)accL;
#pragma unroll
                for (int t = 0; t < NG; ++t) accG64[t] += (double)accG[t];
#pragma unroll
                for (int t = 0; t < NP; ++t) accP64[t] += (double)accP[t];
                if constexpr (NP > 0) {
                    // group boundary (the same row for every lane)
                    while (row_base + i >= nextb && row_base + i < r1) {
#pragma unroll
                        for (int t = 0; t < NP; ++t) {
                            if (inside) {
                                if (live) p.direct[((int64_t)t * p.n_primary + g) * p.bp + chain] = accP64[t];
                            } else {
                                headv[t] = accP64[t];
                            }
                            accP64[t] = 0.0;
                        }
                        ++g;
                        inside = true;
                        nextb = p.seg_offsets[g + 1];
#pragma unroll
                        for (int t = 0; t < NP; ++t) vp[t] = th[p.theta_index_p[t][g]];
                    }
                }
            }
            __syncthreads();  // everybody is done with this stage
            if (threadIdx.x == 0 && tile + 2 < t1) {
                fence_proxy_async();
                issue(tile + 2, s);
            }
        }
        if constexpr (NP > 0) {
            // the group still open at the end of the range
            const bool complete = inside && p.seg_offsets[g + 1] <= r1;
#pragma unroll
            for (int t = 0; t < NP; ++t) {
                if (complete) {
                    if (live) p.direct[((int64_t)t * p.n_primary + g) * p.bp + chain] = accP64[t];
                } else if (g == kf) {
                    headv[t] = accP64[t];
                } else {
                    tailv[t] = accP64[t];
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int t = 0; t <= NG; ++t) p.glob[((int64_t)blockIdx.x * (NG + 1) + t) * p.bp + chain] = accG64[t];
#pragma unroll
        for (int t = 0; t < NP; ++t) {
            p.head[((int64_t)blockIdx.x * NP + t) * p.bp + chain] = headv[t];
            p.tail[((int64_t)blockIdx.x * NP + t) * p.bp + chain] = tailv[t];
        }
    }
}

// ---- reductions over the CTA ranges, and the final combination, chain-minor ----------------------
struct ChainReduceParams {
    const int4 *slot_desc;       // [GP] {c_lo, c_hi, kind of first, kind of last} in units of CTA ranges
    double *totals;              // [(NG+1) + NP*GP][bp]: global terms + logp partial | primary slots (= direct)
    const double *head, *tail, *glob;
    const int32_t *single_unit;  // [P] as in ReduceParams, units of this buffer
    const int32_t *multi_elem;
    const int64_t *multi_ptr;
    const int32_t *multi_unit;
    const int32_t *multi_of;     // [P] position in the multi list, -1
    int32_t ranges, ng, np, bp, batch;
    int64_t n_primary, n_params;
    const double *prior_grad;    // [batch][P]
    const double *prior_lp_part; // [batch][prior_ctas]
    int32_t prior_ctas, include_priors, family;
    int64_t n_rows_total, sigma_theta_index;
    double sigma_const, logp_const;
};

// blockIdx.y == 0: primary slots (thread = chain, blockIdx.x = slot); blockIdx.y == 1: global terms
static __global__ void skk_chain_slot_kernel(const __grid_constant__ ChainReduceParams rp) {
    double *direct = rp.totals + (int64_t)(rp.ng + 1) * rp.bp;
    if (blockIdx.y == 0) {
        for (int chain = threadIdx.x; chain < rp.batch; chain += blockDim.x)
            for (int64_t slot = blockIdx.x; slot < rp.n_primary; slot += gridDim.x) {
                const int4 d = rp.slot_desc[slot];
                if (d.y < d.x) continue;
                for (int t = 0; t < rp.np; ++t) {
                    double acc = direct[((int64_t)t * rp.n_primary + slot) * rp.bp + chain];
                    for (int c = d.x; c <= d.y; ++c) {
                        const int kind = c == d.x ? d.z : (c == d.y ? d.w : 1);
                        const double *src = kind == 1 ? rp.head : (kind == 2 ? rp.tail : nullptr);
                        if (src) acc += src[((int64_t)c * rp.np + t) * rp.bp + chain];
                    }
                    direct[((int64_t)t * rp.n_primary + slot) * rp.bp + chain] = acc;
                }
            }
    } else {
        // global terms and the logp partial: block = (term, 32 chains); the block's warps share the CTA ranges of the
        // chain kernel (warp w: ranges w, w + W, ...; sixteen loads in flight per lane), then the warps' partials are
        // added in warp order -- a fixed order, so the result is reproducible.  (One thread per chain walking all
        // ranges was 50 us of a 120 us evaluation at 128 chains: 592 dependent loads.)
        __shared__ double part[8][32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
        const int chunks = (rp.batch + 31) / 32;
        for (int item = blockIdx.x; item < (rp.ng + 1) * chunks; item += gridDim.x) {
            const int t = item % (rp.ng + 1), chain = (item / (rp.ng + 1)) * 32 + lane;
            const bool valid = chain < rp.batch;
            const double *src = rp.glob + (int64_t)t * rp.bp + (valid ? chain : 0);
            const int64_t stride = (int64_t)(rp.ng + 1) * rp.bp;
            double acc = 0.0;
            for (int c0 = warp; c0 < rp.ranges; c0 += 16 * warps) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = (valid && c0 + u * warps < rp.ranges) ? src[(int64_t)(c0 + u * warps) * stride] : 0.0;
#pragma unroll
                for (int u = 0; u < 16; ++u) acc += v[u];
            }
            part[warp][lane] = acc;
            __syncthreads();
            if (warp == 0 && valid) {
                double tot = 0.0;
                for (int w = 0; w < warps; ++w) tot += part[w][lane];
                rp.totals[(int64_t)t * rp.bp + chain] = tot;
            }
            __syncthreads();
        }
    }
}

// thread = chain (coalesced reads of the totals), blockIdx.x strides over theta elements
template <typename T>
__global__ void skk_chain_final_kernel(const __grid_constant__ ChainReduceParams rp, const T *__restrict__ theta,
                                       T *__restrict__ out_logp, T *__restrict__ out_grad) {
    const int64_t P = rp.n_params;
    for (int chain = threadIdx.x; chain < rp.batch; chain += blockDim.x) {
        double scale = 1.0, sigma = 1.0;
        if (rp.family == SKK_LIK_NORMAL) {
            sigma = rp.sigma_theta_index >= 0 ? exp((double)theta[(int64_t)chain * P + rp.sigma_theta_index])
                                              : rp.sigma_const;
            scale = 1.0 / (sigma * sigma);
        }
        const double L = rp.totals[(int64_t)rp.ng * rp.bp + chain];
        for (int64_t j = blockIdx.x; j < P; j += gridDim.x) {
            const int unit = rp.single_unit[j];
            double lik = 0.0;
            if (unit >= 0) {
                lik = rp.totals[(int64_t)unit * rp.bp + chain];
            } else if (unit == -2) {
                const int64_t m = rp.multi_of[j];
                for (int64_t e = rp.multi_ptr[m]; e < rp.multi_ptr[m + 1]; ++e)
                    lik += rp.totals[(int64_t)rp.multi_unit[e] * rp.bp + chain];
            }
            double gval = lik * scale;
            if (rp.family == SKK_LIK_NORMAL && j == rp.sigma_theta_index) gval += L * scale - (double)rp.n_rows_total;
            if (rp.include_priors) gval += rp.prior_grad[(int64_t)chain * P + j];
            out_grad[(int64_t)chain * P + j] = (T)gval;
        }
        if (blockIdx.x == 0) {
            double tot = 0.0;
            if (rp.include_priors)
                for (int k = 0; k < rp.prior_ctas; ++k) tot += rp.prior_lp_part[(int64_t)chain * rp.prior_ctas + k];
            double ll = L;
            if (rp.family == SKK_LIK_NORMAL)
                ll = -0.5 * L * scale - (double)rp.n_rows_total * (kLogSqrt2Pi + log(sigma));
            out_logp[chain] = (T)(tot + ll + rp.logp_const);
        }
    }
}

}  // namespace skk

namespace skk {
template <typename T, int FAM>
const void *chain_kernel_lookup(int ng, int np) {
#define SKK_CHAIN(NG, NP) \
    if (ng == NG && np == NP) return reinterpret_cast<const void *>(&skk_chain_kernel<T, FAM, NG, NP>);
    SKK_CHAIN(0, 0) SKK_CHAIN(0, 1) SKK_CHAIN(0, 2)
    SKK_CHAIN(1, 0) SKK_CHAIN(1, 1) SKK_CHAIN(1, 2)
    SKK_CHAIN(2, 0) SKK_CHAIN(2, 1) SKK_CHAIN(2, 2)
#undef SKK_CHAIN
    return nullptr;
}
}  // namespace skk
