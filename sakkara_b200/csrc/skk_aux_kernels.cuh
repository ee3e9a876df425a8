// Small per-evaluation kernels around the row kernel (all O(P) or O(#partials), fp64 inside):
//   K0 prepare : gather theta into the value tables the row kernel reads (slot order)
//   K3 priors  : logp term and gradient of every parameter block, incl. hyper-parameter links
//   K4a reduce : add the row kernel's partial slots per theta element in fixed order
//   K4b final  : likelihood scale (Normal), + priors, + constants -> logp[B], dlogp[B, P]
#pragma once

#include "skk_common.cuh"

namespace skk {

constexpr double kLogSqrt2Pi = 0.91893853320467274178032973640562;
constexpr double kLogSqrt2OverPi = -0.22579135264472743236309761494744;

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
    const int64_t *link_ptr;    // [P+1] CSR: hyper-parameter element -> child elements
    const int32_t *link_child;  // child theta element
    const int8_t *link_kind;    // 0: parent is the child's a (mu), 1: the child's b (sigma)
    int64_t n_params;
    int32_t batch;
};

// standardised residual and scale of a Normal-prior element
template <typename T>
__device__ __forceinline__ void normal_prior(const DevBlock &blk, int64_t j, const T *__restrict__ th, double &z,
                                             double &sd) {
    const int64_t g = j - blk.offset;
    const double mu = blk.a_index ? (double)th[blk.a_index[g]] : blk.a_const[blk.a_const_len == 1 ? 0 : g];
    sd = blk.b_index ? exp((double)th[blk.b_index[g]]) : blk.b_const[blk.b_const_len == 1 ? 0 : g];
    z = ((double)th[j] - mu) / sd;
}

template <typename T>
__global__ void skk_prior_kernel(const __grid_constant__ PriorParams pp, const T *__restrict__ theta,
                                 double *__restrict__ prior_grad, double *__restrict__ prior_lp) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= pp.n_params) return;
    const DevBlock blk = pp.blocks[pp.block_of[j]];
    for (int b = 0; b < pp.batch; ++b) {
        const T *th = theta + (int64_t)b * pp.n_params;
        double lp, g;
        const double v = (double)th[j];
        if (blk.family == SKK_PRIOR_NORMAL) {
            double z, sd;
            normal_prior(blk, j, th, z, sd);
            lp = -0.5 * z * z - kLogSqrt2Pi - log(sd);
            g = -z / sd;
        } else if (blk.family == SKK_PRIOR_HALFNORMAL) {
            const double tau = blk.a_const[blk.a_const_len == 1 ? 0 : j - blk.offset];
            const double q = exp(v) / tau;
            lp = -0.5 * q * q + kLogSqrt2OverPi - log(tau) + v;
            g = -q * q + 1.0;
        } else {  // SKK_PRIOR_EXPONENTIAL
            const double lam = blk.a_const[blk.a_const_len == 1 ? 0 : j - blk.offset];
            const double s = exp(v);
            lp = log(lam) - lam * s + v;
            g = -lam * s + 1.0;
        }
        // children whose prior has this element as mu or (log) sigma, in fixed order
        for (int64_t e = pp.link_ptr[j]; e < pp.link_ptr[j + 1]; ++e) {
            const int64_t c = pp.link_child[e];
            double z, sd;
            normal_prior(pp.blocks[pp.block_of[c]], c, th, z, sd);
            g += pp.link_kind[e] == 0 ? z / sd : z * z - 1.0;
        }
        prior_grad[(int64_t)b * pp.n_params + j] = g;
        prior_lp[(int64_t)b * pp.n_params + j] = lp;
    }
}

// ---- K4a ------------------------------------------------------------------------------------
struct ReduceParams {
    const int64_t *csr_ptr;     // [P+2]: entries of theta element j; element P = the logp partial
    const int8_t *entry_level;
    const int8_t *entry_t;      // term number within its level
    const int32_t *entry_slot;
    const int2 *tile_keys;
    const int64_t *seg_offsets;
    const double *direct_p;     // [NP][GP][BT]
    const double *tile_head;    // [tiles][NP][BT]
    const double *tile_tail;
    const double *cta_glob;     // [grid][NG+1][BT]
    const double *cta_sec;      // [grid][NS][GS][BT]
    double *lik_vec;            // [B][P+2]
    int64_t n_params;
    int64_t n_primary, n_secondary;
    int32_t ng, np, ns;
    int32_t bt, b0, b_valid;
    int32_t grid;               // CTAs of the row kernel
    int32_t rows_per_tile;
    double logp_const;
};

template <int BT>
__global__ void skk_reduce_kernel(const __grid_constant__ ReduceParams rp) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;  // one warp per element
    if (j > rp.n_params) return;
    double acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = 0.0;

    if (j == rp.n_params) {
        // likelihood logp partial: last entry of every CTA's global block
        for (int c = lane; c < rp.grid; c += 32)
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] += rp.cta_glob[((int64_t)c * (rp.ng + 1) + rp.ng) * BT + b];
    } else {
        for (int64_t e = rp.csr_ptr[j]; e < rp.csr_ptr[j + 1]; ++e) {
            const int level = rp.entry_level[e], t = rp.entry_t[e];
            const int64_t slot = rp.entry_slot[e];
            if (level == SKK_LEVEL_GLOBAL) {
                for (int c = lane; c < rp.grid; c += 32)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += rp.cta_glob[((int64_t)c * (rp.ng + 1) + t) * BT + b];
            } else if (level == SKK_LEVEL_SECONDARY) {
                for (int c = lane; c < rp.grid; c += 32)
#pragma unroll
                    for (int b = 0; b < BT; ++b)
                        acc[b] += rp.cta_sec[(((int64_t)c * rp.ns + t) * rp.n_secondary + slot) * BT + b];
            } else {
                const int64_t r0 = rp.seg_offsets[slot], r1 = rp.seg_offsets[slot + 1];
                if (r1 > r0) {
                    if (lane == 0)
#pragma unroll
                        for (int b = 0; b < BT; ++b) acc[b] += rp.direct_p[((int64_t)t * rp.n_primary + slot) * BT + b];
                    const int64_t t_lo = r0 / rp.rows_per_tile, t_hi = (r1 - 1) / rp.rows_per_tile;
                    for (int64_t tile = t_lo + lane; tile <= t_hi; tile += 32) {
                        const int2 k = rp.tile_keys[tile];
                        const double *src = k.x == slot ? rp.tile_head : (k.y == slot ? rp.tile_tail : nullptr);
                        if (src)
#pragma unroll
                            for (int b = 0; b < BT; ++b) acc[b] += src[((int64_t)tile * rp.np + t) * BT + b];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        const double s = warp_sum(acc[b]);
        if (lane == 0 && b < rp.b_valid) rp.lik_vec[(int64_t)(rp.b0 + b) * (rp.n_params + 2) + j] = s;
    }
    if (j == rp.n_params && lane == 0)
        for (int b = 0; b < rp.b_valid; ++b)
            rp.lik_vec[(int64_t)(rp.b0 + b) * (rp.n_params + 2) + rp.n_params + 1] = rp.logp_const;
}

// ---- K4b ------------------------------------------------------------------------------------
struct FinalParams {
    const double *lik_vec;      // [B][P+2] (all-reduced over ranks when sharded)
    const double *prior_grad;   // [B][P]
    const double *prior_lp;     // [B][P]
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
        // logp: priors summed in fixed order (strided per thread, then a fixed tree)
        __shared__ double part[32];
        double s = 0.0;
        if (fp.include_priors)
            for (int64_t k = threadIdx.x; k < P; k += blockDim.x) s += fp.prior_lp[(int64_t)b * P + k];
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
