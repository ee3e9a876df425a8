// Small per-evaluation kernels around the row kernel.  All are O(P) or O(#partial slots), use fp64
// internally and add in a fixed order (serial per thread, or lane-strided + fixed shuffle tree), so
// the whole evaluation is bit-reproducible.
//   K0  prepare      : gather theta into the value tables the row kernel reads (slot order)
//   K3a prior terms  : per theta element, its own prior logp term and gradient, and what it
//                      contributes to the gradient of its prior's mu / sigma hyper-parameters
//   K3b prior links  : per hyper-parameter element, the sum over its children (CSR, built at plan time)
//   K4a slot totals  : per (term, slot): direct_p + tile partials | sum over CTA partials
//   K4t theta reduce : per theta element, the sum over the slots that map to it  -> lik_vec
//   K4b final        : likelihood scale (Normal), + priors, + constants -> logp[B], dlogp[B, P]
#pragma once

#include "skk_common.cuh"

namespace skk {

constexpr double kLogSqrt2Pi = 0.91893853320467274178032973640562;
constexpr double kLogSqrt2OverPi = -0.22579135264472743236309761494744;
constexpr int kSerialLimit = 48;  // longer partial lists are summed by a whole warp

// Runs body(first, last, item) for every lane of the warp whose list [first, last) is long, with
// all 32 lanes cooperating on one list at a time: lane-strided partial sums + fixed shuffle tree.
// `value(e, b)` returns the e-th addend of chain b; `store(item, b, sum)` is called on lane 0.
template <int BT, typename Value, typename Store>
__device__ __forceinline__ void warp_cooperative_lists(bool heavy, int64_t first, int64_t last, int64_t item,
                                                       Value value, Store store) {
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(kFull, heavy);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t lo = __shfl_sync(kFull, first, src), hi = __shfl_sync(kFull, last, src);
        const int64_t it = __shfl_sync(kFull, item, src);
        double acc[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] = 0.0;
        for (int64_t e = lo + lane; e < hi; e += 32)
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] += value(it, e, b);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const double s = warp_sum(acc[b]);
            if (lane == 0) store(it, b, s);
        }
    }
}

// ---- K0 -------------------------------------------------------------------------------------
struct PrepTerm {
    const int32_t *theta_index;  // [n_slots]
    int64_t n_slots;
    int64_t out_offset;          // element offset (in units of BT values) into the level's table
    int32_t level;
};
constexpr int kMaxTerms = 3 * kMaxTermsPerLevel;
struct PrepParams {
    PrepTerm term[kMaxTerms];
    int32_t n_terms;
    int32_t bt;          // chains in this batch tile
    int32_t b_valid;     // chains of the tile that exist (others read chain 0 and are discarded)
    int64_t n_params;
};

template <typename T>
__global__ void skk_prepare_kernel(const __grid_constant__ PrepParams pp, const T *__restrict__ theta,
                                   T *__restrict__ val_g, T *__restrict__ val_p, T *__restrict__ val_s) {
    const PrepTerm &tm = pp.term[blockIdx.y];
    T *out = tm.level == SKK_LEVEL_GLOBAL ? val_g : tm.level == SKK_LEVEL_PRIMARY ? val_p : val_s;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < tm.n_slots;
         s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = tm.theta_index[s];
        for (int b = 0; b < pp.bt; ++b) {
            const int bs = b < pp.b_valid ? b : 0;
            out[(tm.out_offset + s) * pp.bt + b] = theta[(int64_t)bs * pp.n_params + j];
        }
    }
}

// ---- K3 -------------------------------------------------------------------------------------
struct DevBlock {
    int32_t family;
    int32_t pad;
    int64_t offset;
    int64_t size;
    const double *a_const;
    int64_t a_const_len;
    const int32_t *a_index;
    const double *b_const;
    int64_t b_const_len;
    const int32_t *b_index;
};

struct PriorParams {
    const DevBlock *blocks;
    const int32_t *block_of;    // [P]
    // hyper-parameter elements ("parents") and their children, heavy parents first
    const int32_t *parent;      // [n_parents] theta element
    const int64_t *link_ptr;    // [n_parents + 1]
    const int32_t *link_child;  // child theta element
    const int8_t *link_kind;    // 0: parent is the child's mu, 1: the child's (log) sigma
    int32_t n_parents;
    int32_t n_heavy;            // parents handled by a whole CTA
    int64_t n_params;
    int32_t batch;
};

constexpr int kPriorThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kPriorThreads)
skk_prior_kernel(const __grid_constant__ PriorParams pp, const T *__restrict__ theta, double *__restrict__ prior_grad,
                 double *__restrict__ prior_lp_part, double *__restrict__ to_mu, double *__restrict__ to_sigma) {
    __shared__ double part[kPriorThreads / 32];
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = j < pp.n_params;
    DevBlock blk{};
    int64_t g = 0;
    if (valid) {
        blk = pp.blocks[pp.block_of[j]];
        g = j - blk.offset;
    }
    for (int b = 0; b < pp.batch; ++b) {
        const T *th = theta + (int64_t)b * pp.n_params;
        double lp = 0.0;
        if (valid) {
            const double v = (double)th[j];
            double grad, cm = 0.0, cs = 0.0;
            if (blk.family == SKK_PRIOR_NORMAL) {
                const double mu = blk.a_index ? (double)th[blk.a_index[g]] : blk.a_const[blk.a_const_len == 1 ? 0 : g];
                const double sd =
                    blk.b_index ? exp((double)th[blk.b_index[g]]) : blk.b_const[blk.b_const_len == 1 ? 0 : g];
                const double z = (v - mu) / sd;
                lp = -0.5 * z * z - kLogSqrt2Pi - log(sd);
                grad = -z / sd;
                cm = z / sd;       // d/d mu
                cs = z * z - 1.0;  // d/d log(sigma)
            } else if (blk.family == SKK_PRIOR_HALFNORMAL) {
                const double tau = blk.a_const[blk.a_const_len == 1 ? 0 : g];
                const double q = exp(v) / tau;
                lp = -0.5 * q * q + kLogSqrt2OverPi - log(tau) + v;
                grad = -q * q + 1.0;
            } else {  // SKK_PRIOR_EXPONENTIAL
                const double lam = blk.a_const[blk.a_const_len == 1 ? 0 : g];
                const double sv = exp(v);
                lp = log(lam) - lam * sv + v;
                grad = -lam * sv + 1.0;
            }
            const int64_t o = (int64_t)b * pp.n_params + j;
            prior_grad[o] = grad;
            to_mu[o] = cm;
            to_sigma[o] = cs;
        }
        // logp of the priors: fixed tree per CTA, the CTA partials are added in order by the final kernel
        const double s = warp_sum(lp);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kPriorThreads / 32; ++w) tot += part[w];
            prior_lp_part[(int64_t)b * gridDim.x + blockIdx.x] = tot;
        }
        __syncthreads();
    }
}

// One CTA per heavy parent (blockIdx.x < n_heavy), one warp per light parent afterwards.
__global__ void skk_prior_link_kernel(const __grid_constant__ PriorParams pp, double *__restrict__ prior_grad,
                                      const double *__restrict__ to_mu, const double *__restrict__ to_sigma) {
    __shared__ double part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const bool heavy = (int)blockIdx.x < pp.n_heavy;
    const int64_t p = heavy ? blockIdx.x : pp.n_heavy + ((int64_t)blockIdx.x - pp.n_heavy) * warps + warp;
    if (p >= pp.n_parents) return;
    const int64_t lo = pp.link_ptr[p], hi = pp.link_ptr[p + 1];
    const int64_t first = heavy ? lo + threadIdx.x : lo + lane;
    const int64_t step = heavy ? blockDim.x : 32;
    for (int b = 0; b < pp.batch; ++b) {
        const double *cm = to_mu + (int64_t)b * pp.n_params, *cs = to_sigma + (int64_t)b * pp.n_params;
        double acc = 0.0;
        int64_t e = first;
        for (; e + 3 * step < hi; e += 4 * step) {  // four independent gathers in flight
            const int64_t c0 = pp.link_child[e], c1 = pp.link_child[e + step], c2 = pp.link_child[e + 2 * step],
                          c3 = pp.link_child[e + 3 * step];
            const int8_t k0 = pp.link_kind[e], k1 = pp.link_kind[e + step], k2 = pp.link_kind[e + 2 * step],
                         k3 = pp.link_kind[e + 3 * step];
            const double v0 = k0 == 0 ? cm[c0] : cs[c0], v1 = k1 == 0 ? cm[c1] : cs[c1];
            const double v2 = k2 == 0 ? cm[c2] : cs[c2], v3 = k3 == 0 ? cm[c3] : cs[c3];
            acc += v0;
            acc += v1;
            acc += v2;
            acc += v3;
        }
        for (; e < hi; e += step) {
            const int64_t c = pp.link_child[e];
            acc += pp.link_kind[e] == 0 ? cm[c] : cs[c];
        }
        acc = warp_sum(acc);
        if (heavy) {
            if (lane == 0) part[warp] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                double s = 0.0;
                for (int w = 0; w < warps; ++w) s += part[w];
                prior_grad[(int64_t)b * pp.n_params + pp.parent[p]] += s;
            }
            __syncthreads();
        } else if (lane == 0) {
            prior_grad[(int64_t)b * pp.n_params + pp.parent[p]] += acc;
        }
    }
}

// ---- K4a ------------------------------------------------------------------------------------
struct ReduceParams {
    const int64_t *csr_ptr;     // [P+1]: slots of theta element j
    const int8_t *entry_level;
    const int8_t *entry_t;      // term number within its level
    const int32_t *entry_slot;
    const int2 *tile_keys;
    const int64_t *seg_offsets;
    double *direct_p;           // [NP][GP][BT]  in: groups inside one tile; out: total per primary slot
    const double *tile_head;    // [tiles][NP][BT]
    const double *tile_tail;
    const double *cta_glob;     // [grid][NG+1][BT]
    const double *cta_sec;      // [grid][NS][GS][BT]
    double *glob_tot;           // [NG+1][BT]
    double *sec_tot;            // [NS][GS][BT]
    double *lik_vec;            // [B][P+2]
    int64_t n_params;
    int64_t n_primary, n_secondary;
    int32_t ng, np, ns;
    int32_t bt, b0, b_valid;
    int32_t grid;               // CTAs of the row kernel
    int32_t rows_per_tile;
    int32_t blocks_p, blocks_s; // CTAs of the slot-total kernel working on the primary / secondary level
    double logp_const;
};

__device__ __forceinline__ int64_t t_lo_of(const ReduceParams &rp, int64_t slot) {
    return rp.seg_offsets[slot] / rp.rows_per_tile;
}
__device__ __forceinline__ int64_t t_hi_of(const ReduceParams &rp, int64_t slot) {
    return (rp.seg_offsets[slot + 1] - 1) / rp.rows_per_tile;
}

// blockIdx.x in [0, blocks_p): primary slots, one per thread;  [blocks_p, blocks_p + blocks_s):
// secondary slots, 32 per CTA x 8 ranges of row-kernel CTAs;  last CTA: global terms and logp.
template <int BT>
__global__ void __launch_bounds__(256) skk_slot_total_kernel(const __grid_constant__ ReduceParams rp) {
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x < rp.blocks_p) {
        const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool valid = slot < rp.n_primary;
        int64_t t_lo = 0, t_hi = -1;
        if (valid) {
            const int64_t r0 = rp.seg_offsets[slot], r1 = rp.seg_offsets[slot + 1];
            if (r1 > r0) {
                t_lo = r0 / rp.rows_per_tile;
                t_hi = (r1 - 1) / rp.rows_per_tile;
            }
        }
        // a tile strictly inside the slot's row range lies wholly in the slot: its head partial is
        // the slot's; only the first and the last tile need the key check
        auto addend = [&](int64_t s, int64_t tile, int t, int b) -> double {
            const double *src = rp.tile_head;
            if (tile == t_lo_of(rp, s) || tile == t_hi_of(rp, s)) {
                const int2 k = rp.tile_keys[tile];
                src = k.x == s ? rp.tile_head : (k.y == s ? rp.tile_tail : nullptr);
            }
            return src ? src[(tile * rp.np + t) * BT + b] : 0.0;
        };
        const bool heavy = valid && (t_hi - t_lo + 1 > kSerialLimit);
        for (int t = 0; t < rp.np; ++t) {
            if (valid && !heavy && t_hi >= t_lo) {
                double acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = rp.direct_p[((int64_t)t * rp.n_primary + slot) * BT + b];
                const int2 k_lo = rp.tile_keys[t_lo], k_hi = rp.tile_keys[t_hi];
                const double *s_lo = k_lo.x == slot ? rp.tile_head : (k_lo.y == slot ? rp.tile_tail : nullptr);
                const double *s_hi = k_hi.x == slot ? rp.tile_head : (k_hi.y == slot ? rp.tile_tail : nullptr);
                if (s_lo)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += s_lo[(t_lo * rp.np + t) * BT + b];
#pragma unroll 4
                for (int64_t tile = t_lo + 1; tile < t_hi; ++tile)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += rp.tile_head[(tile * rp.np + t) * BT + b];
                if (s_hi && t_hi > t_lo)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += s_hi[(t_hi * rp.np + t) * BT + b];
#pragma unroll
                for (int b = 0; b < BT; ++b) rp.direct_p[((int64_t)t * rp.n_primary + slot) * BT + b] = acc[b];
            }
            warp_cooperative_lists<BT>(
                heavy, t_lo, t_hi + 1, slot, [&](int64_t s, int64_t tile, int b) { return addend(s, tile, t, b); },
                [&](int64_t s, int b, double sum) { rp.direct_p[((int64_t)t * rp.n_primary + s) * BT + b] += sum; });
        }
    } else if ((int)blockIdx.x < rp.blocks_p + rp.blocks_s) {
        // 32 consecutive slots (coalesced) x 8 sub-ranges of CTAs, combined in fixed order
        __shared__ double part[8][32];
        const int sub = threadIdx.x >> 5;
        const int64_t slot = ((int64_t)blockIdx.x - rp.blocks_p) * 32 + lane;
        const int per = (rp.grid + 7) / 8;
        const int c0 = sub * per, c1 = min(rp.grid, c0 + per);
        for (int t = 0; t < rp.ns; ++t)
            for (int b = 0; b < BT; ++b) {
                double acc = 0.0;
                if (slot < rp.n_secondary)
                    for (int c = c0; c < c1; ++c)
                        acc += rp.cta_sec[(((int64_t)c * rp.ns + t) * rp.n_secondary + slot) * BT + b];
                part[sub][lane] = acc;
                __syncthreads();
                if (sub == 0 && slot < rp.n_secondary) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) s += part[k][lane];
                    rp.sec_tot[((int64_t)t * rp.n_secondary + slot) * BT + b] = s;
                }
                __syncthreads();
            }
    } else {
        // global terms and the logp partial: one warp per (term, chain)
        const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
        for (int item = warp; item < (rp.ng + 1) * BT; item += warps) {
            double acc = 0.0;
            for (int c = lane; c < rp.grid; c += 32) acc += rp.cta_glob[(int64_t)c * (rp.ng + 1) * BT + item];
            acc = warp_sum(acc);
            if (lane == 0) rp.glob_tot[item] = acc;
        }
    }
}

// per theta element: sum of the totals of the slots that map to it; one warp per element, lanes
// strided over the element's slots, fixed shuffle tree
template <int BT>
__global__ void __launch_bounds__(256) skk_theta_reduce_kernel(const __grid_constant__ ReduceParams rp) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (j >= rp.n_params) return;
    const int64_t lo = rp.csr_ptr[j], hi = rp.csr_ptr[j + 1];
    double acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = 0.0;
    for (int64_t e = lo + lane; e < hi; e += 32) {
        const int level = rp.entry_level[e], t = rp.entry_t[e];
        const int64_t slot = rp.entry_slot[e];
        const double *src = level == SKK_LEVEL_GLOBAL    ? rp.glob_tot + (int64_t)t * BT
                            : level == SKK_LEVEL_PRIMARY ? rp.direct_p + ((int64_t)t * rp.n_primary + slot) * BT
                                                         : rp.sec_tot + ((int64_t)t * rp.n_secondary + slot) * BT;
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] += src[b];
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        const double sum = (hi - lo > 1) ? warp_sum(acc[b]) : acc[b];
        if (lane == 0 && b < rp.b_valid) rp.lik_vec[(int64_t)(rp.b0 + b) * (rp.n_params + 2) + j] = sum;
    }
    if (j == 0 && lane == 0)
        for (int b = 0; b < rp.b_valid; ++b) {
            double *row = rp.lik_vec + (int64_t)(rp.b0 + b) * (rp.n_params + 2);
            row[rp.n_params] = rp.glob_tot[rp.ng * BT + b];  // likelihood logp partial
            row[rp.n_params + 1] = rp.logp_const;
        }
}

// ---- K4b ------------------------------------------------------------------------------------
struct FinalParams {
    const double *lik_vec;      // [B][P+2] (all-reduced over ranks when sharded)
    const double *prior_grad;   // [B][P]
    const double *prior_lp_part; // [B][prior_ctas] per-CTA partial sums of the prior logp terms
    int32_t prior_ctas;
    int64_t n_params;
    int64_t n_rows_total;
    int64_t sigma_theta_index;  // -1: constant sigma
    double sigma_const;
    int32_t family;
    int32_t include_priors;
};

template <typename T>
__global__ void skk_final_kernel(const __grid_constant__ FinalParams fp, const T *__restrict__ theta,
                                 T *__restrict__ out_logp, T *__restrict__ out_grad) {
    const int b = blockIdx.y;
    const int64_t P = fp.n_params;
    const double *lik = fp.lik_vec + (int64_t)b * (P + 2);
    double scale = 1.0, sigma = 1.0;
    if (fp.family == SKK_LIK_NORMAL) {
        sigma = fp.sigma_theta_index >= 0 ? exp((double)theta[(int64_t)b * P + fp.sigma_theta_index]) : fp.sigma_const;
        scale = 1.0 / (sigma * sigma);
    }
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < P) {
        double g = lik[j] * scale;
        if (fp.family == SKK_LIK_NORMAL && j == fp.sigma_theta_index)
            g += lik[P] * scale - (double)fp.n_rows_total;  // sum(r^2 - 1), r = e / sigma
        if (fp.include_priors) g += fp.prior_grad[(int64_t)b * P + j];
        out_grad[(int64_t)b * P + j] = (T)g;
    }
    if (blockIdx.x == 0) {
        // logp: CTA partials of the prior kernel, added in fixed order (strided per thread, fixed tree)
        __shared__ double part[32];
        double s = 0.0;
        if (fp.include_priors)
            for (int k = threadIdx.x; k < fp.prior_ctas; k += blockDim.x)
                s += fp.prior_lp_part[(int64_t)b * fp.prior_ctas + k];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
            double ll;
            if (fp.family == SKK_LIK_NORMAL)
                ll = -0.5 * lik[P] * scale - (double)fp.n_rows_total * (kLogSqrt2Pi + log(sigma));
            else
                ll = lik[P];
            out_logp[b] = (T)(tot + ll + lik[P + 1]);
        }
    }
}

}  // namespace skk
