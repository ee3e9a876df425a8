// Shared definitions of the sm_100a kernels: parameter structs, PTX wrappers (mbarrier, TMA bulk
// copy), warp primitives and the likelihood families.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skk_b200.h"

namespace skk {

constexpr int kMaxXCols = 6;      // float data columns a plan may stage (besides y / offset)
constexpr int kMaxTermsPerLevel = 2;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Parameters of the row kernel (K1).  Passed by value (__grid_constant__).
//   Row data lives in device columns padded to n_tiles * rows_per_tile rows (zero filled), so a
//   whole tile can always be fetched with one bulk copy per column.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct RowParams {
    // row data
    const T *x[kMaxXCols];
    const T *y;              // observed column stored as T ...
    const uint8_t *y8;       // ... or as uint8 (exactly one of y / y8 is non-null)
    const T *eta_offset;     // optional constant part of eta
    const void *sec_codes;   // uint16 or int32 per row
    int32_t n_x;
    int32_t sec_code_bytes;
    int64_t n_rows;
    int32_t n_tiles;
    int32_t stages;          // smem ring depth per warp
    // structure
    const int32_t *sched;    // static schedule: [grid * warps][sched_len + 4] tiles of every warp in order, -1 padded
    int32_t sched_len;       // multiple of 4
    int32_t sched_folds;     // folds of the fp32 accumulator tables every kFoldTiles tiles (same for all warps)
    const int4 *tile_meta;   // per tile {first, last primary slot, theta element of primary term 0 / 1 at the first slot}
    const int32_t *tile_split;  // per tile: row (inside the tile) where its second group starts when the tile
                                // holds exactly two groups, else -1
    const int64_t *seg_offsets;  // [n_primary + 1]
    int32_t n_primary;
    int32_t n_secondary;
    int32_t xcol_g[kMaxTermsPerLevel];
    int32_t xcol_p[kMaxTermsPerLevel];
    int32_t xcol_s[kMaxTermsPerLevel];
    // parameter values are read straight from theta through the slot -> theta tables of the terms
    const T *theta;          // [b_valid][n_params], first chain of this batch tile
    int64_t n_params;
    int32_t b_valid;         // chains of the batch tile that exist (the others repeat chain 0)
    const int32_t *ti_g[kMaxTermsPerLevel];  // [1]
    const int32_t *ti_p[kMaxTermsPerLevel];  // [n_primary]
    const int32_t *ti_s[kMaxTermsPerLevel];  // [n_secondary]
    // outputs, all fp64
    double *direct_p;        // [NP][n_primary][BT]   groups that lie inside one tile; the set of groups written
                             //                       here is fixed by the tiling, all others stay zero
    double *tile_head;       // [n_tiles][NP][BT]     partial of the tile's first group
    double *tile_tail;       // [n_tiles][NP][BT]     partial of the tile's last group (if != first)
    double *cta_glob;        // [grid][NG + 1][BT]    global-term partials, last entry = logp partial
    double *cta_sec;         // [grid][NS][n_secondary][BT]
    // prior terms, evaluated by the row kernel while its first tiles are in flight (single-GPU path;
    // prior_family == nullptr: the prior kernels run instead).  Flat per theta element, see PriorParams.
    const int8_t *prior_family;
    const int32_t *prior_a_src, *prior_b_src;
    const double *prior_a_const, *prior_b_const;
    double *prior_grad;      // [b_valid][n_params]      d prior / d element (own term only)
    double *prior_lp_part;   // [b_valid][grid * warps]  prior logp, one partial per warp
    // What the children add to the gradient of their priors' parameters, in link order: the links of
    // a hyper-parameter are contiguous and padded to whole chunks of 32 (child -1), so a warp sums one
    // chunk and the finish kernel adds 32 times fewer values per hyper-parameter.
    const int32_t *link_child;   // [n_chunks * 32] theta element of the child (Normal prior), -1: padding
    const int8_t *link_kind;     //                 0: link to the prior's mu, 1: to its sigma (+4: the child's mu is exp(u))
    const int32_t *link_a_src, *link_b_src;   //    the child's prior parameters, as in PriorParams
    const double *link_a_const, *link_b_const;
    double *link_part;       // [b_valid][n_chunks]
    int64_t n_chunks;
    // >= 0: the secondary term's slot -> theta table is `sec_theta_base + slot` (the planner numbers the slots that
    // way when it can): the value table is filled straight from theta, without the gather through ti_s
    int32_t sec_theta_base;
    // constants of the likelihood family (NegativeBinomial: alpha and log(alpha))
    T fam_a, fam_b;
    // row-sharded runs over peer memory: the evaluation counter of the exchange, advanced by one thread of the
    // evaluation's first row kernel (null otherwise); the kernels behind it in the stream read it
    unsigned long long *peer_seq;
    // optional per-warp timeline (skk_trace_enable): [grid * warps][kTraceWords], null in normal operation
    unsigned long long *trace;
    // when the kernel that follows in the stream (launched with the PDL attribute) may start: 0 at this
    // kernel's end as usual, 1 as soon as this kernel runs, 2 when the tile loops are through
    int32_t pdl_trigger;
};

// words of one warp's trace record
//   0 %globaltimer at kernel entry     1 ... when the prologue is done (before the first tile wait)
//   2 ... when the first tile has arrived   3 ... at the end of the tile loop   4 ... at kernel exit
//   5, 6, 7 tiles of the warp with one group / two groups / more
constexpr int kTraceWords = 16;  // (row kernel: 0-4 timeline, 5-7 tiles per class, 8-13 inside prologue and epilogue)

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

// One lane of a converged warp, chosen by the hardware (SASS ELECT): unlike `lane == 0` the
// predicate is warp-uniform for the compiler, so the address arithmetic of the TMA issue that
// follows stays in the uniform datapath.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// shared-memory load / store by 32-bit shared address (keeps a table base in one register)
__device__ __forceinline__ float lds(uint32_t a, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds(uint32_t a, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// The same accesses as pure operations for the compiler, for the read-modify-write of the secondary
// accumulator table inside the row loop.  There every table entry is touched by one
// read-modify-write per tile (codes ascend inside a primary segment, so a lane never returns to an
// entry it has stored), the table aliases none of the plain loads of the loop (staging ring, value
// table), and an entry's address depends on data loaded after the tile's mbarrier wait.  So the load
// may be scheduled freely inside the tile (`lds_free_if`), and the predicated store (`sts_free_if`) is
// tied only to its operands and to a token that `table_stores_done` consumes before the __syncwarp
// that precedes the plain accesses to the table after the loop.  The loads of the following rows
// can then overtake a row's store instead of waiting for its load -> add -> store chain.
// (The predicates are applied inside the asm: a pure asm under an `if` could be if-converted.)
// The load is predicated too (the value is zero where the predicate is false): lanes that will not
// store their entry do not take part in the gather, which removes most of its bank conflicts.
__device__ __forceinline__ float lds_free_if(bool pred, uint32_t a, float) {
    float v;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@p ld.shared.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "r"(a), "r"((uint32_t)pred));
    return v;
}
__device__ __forceinline__ double lds_free_if(bool pred, uint32_t a, double) {
    double v;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "mov.f64 %0, 0d0000000000000000;\n\t"
        "@p ld.shared.f64 %0, [%1];\n\t}"
        : "=d"(v)
        : "r"(a), "r"((uint32_t)pred));
    return v;
}
__device__ __forceinline__ void sts_free_if(bool pred, uint32_t a, float v, uint32_t &token) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %3, 0;\n\t"
        "@p st.shared.f32 [%1], %2;\n\t}"
        : "+r"(token)
        : "r"(a), "f"(v), "r"((uint32_t)pred));
}
__device__ __forceinline__ void sts_free_if(bool pred, uint32_t a, double v, uint32_t &token) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %3, 0;\n\t"
        "@p st.shared.f64 [%1], %2;\n\t}"
        : "+r"(token)
        : "r"(a), "d"(v), "r"((uint32_t)pred));
}
__device__ __forceinline__ void table_stores_done(uint32_t token) {
    asm volatile("" ::"r"(token) : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may start while its predecessor in the stream is still running -- as soon as every CTA of the predecessor
// has executed grid_dep_launch_dependents() (or exited) and SM resources are free -- and must call
// grid_dep_wait() before it touches anything the predecessor writes; the wait returns when the predecessor
// has completed and its memory is visible.  Both are no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void grid_dep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// order generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// warp primitives (all reductions use a fixed tree => bit-reproducible)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(kFull, v, d);
    return v;  // valid in lane 0
}

// Among the lanes with `has`, group lanes holding the same key; the lowest lane of each group
// receives the group's sum (added in ascending lane order).  Returns true on that lane.
// Works for any arrangement of keys.  Must be called by all 32 lanes.
template <int NV>
__device__ __forceinline__ bool warp_match_sum(bool has, int key, double (&v)[NV]) {
    const unsigned active = __ballot_sync(kFull, has);
    bool leader = false;
    if (has) {
        const unsigned peers = __match_any_sync(active, key);
        const int lane = threadIdx.x & 31;
        const int first = __ffs(peers) - 1;
        leader = (lane == first);
        if (peers & (peers - 1)) {  // more than one lane holds this key
            double tot[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) tot[k] = 0.0;
            for (unsigned rest = peers; rest; rest &= rest - 1) {
                const int src = __ffs(rest) - 1;
#pragma unroll
                for (int k = 0; k < NV; ++k) tot[k] += __shfl_sync(peers, v[k], src);
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = tot[k];
        }
    }
    __syncwarp();
    return leader;
}

// Same contract for keys that are non-decreasing along the participating lanes (equal keys occupy
// consecutive lanes): a fixed five-step shuffle tree, no loops and no divergence.
template <typename V, int NV>
__device__ __forceinline__ bool warp_sorted_run_sum(bool has, int key, V (&v)[NV]) {
    const int lane = threadIdx.x & 31;
    const int k = has ? key : -1 - lane;  // lanes without a value never merge with a neighbour
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int other = __shfl_down_sync(kFull, k, d);
        const bool take = (lane + d < 32) && (other == k);
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const V w = __shfl_down_sync(kFull, v[q], d);
            if (take) v[q] += w;
        }
    }
    const int before = __shfl_up_sync(kFull, k, 1);
    return has && (lane == 0 || before != k);
}

// ---------------------------------------------------------------------------------------------
// likelihood families: per row, contribution to logp and d logp / d eta
//   Normal   : the kernel accumulates e = y - eta and e^2; 1/sigma^2 and the constants are applied
//              once per evaluation in the finalise kernel.
//   Bernoulli: y*eta - softplus(eta),            d = y - sigmoid(eta)
//   Poisson  : y*eta - exp(eta) (-lgamma(y+1) is a per-plan constant),  d = y - exp(eta)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Math;
// float: hardware approximations (MUFU.EX2 / LG2 / RCP, ~2^-21 relative error) -- three orders of
// magnitude inside the 1e-4 tolerance stated for fp32; double: the accurate library functions.
constexpr double kLogSqrt2Pi = 0.91893853320467274178032973640562;
constexpr double kLogSqrt2OverPi = -0.22579135264472743236309761494744;

// The per-element family byte of the flattened priors: SKK_PRIOR_* in the low bits, plus a flag when the
// Normal prior's mu is a positive (log-transformed) variable u, i.e. mu = exp(u).
constexpr int kPriorFamilyMask = 15;
constexpr int kPriorMuIsLog = 16;
__device__ __forceinline__ double prior_mu_of(int family_byte, double raw) {
    return (family_byte & kPriorMuIsLog) ? exp(raw) : raw;
}

// positive families: PyMC samples them on the log scale (value variable <name>_log__, LogTransform)
__host__ __device__ __forceinline__ bool prior_is_log(int family) {
    return family == SKK_PRIOR_HALFNORMAL || family == SKK_PRIOR_EXPONENTIAL || family == SKK_PRIOR_HALFCAUCHY ||
           family == SKK_PRIOR_GAMMA || family == SKK_PRIOR_INVGAMMA || family == SKK_PRIOR_LOGNORMAL ||
           family == SKK_PRIOR_HALFSTUDENTT;
}
// families >= SKK_PRIOR_HALFCAUCHY: constant parameters only, evaluated without their normaliser (which the plan
// carries in logp_const); plans that contain one evaluate their priors in the separate prior kernels, so that the
// row kernel's prologue keeps the three basic families only (EXT = false)
__host__ __device__ __forceinline__ bool prior_is_extended(int family) { return family >= SKK_PRIOR_HALFCAUCHY; }

// One prior term in fp64.  v: the (transformed) element; a, b: mu and sd (Normal) or the rate /
// scale in a (Exponential, HalfNormal; sampled as log, the log-Jacobian v is included).
// Returns logp; grad = d/dv; for the Normal also d/d mu and d/d log(sd).
template <bool EXT = true>
__device__ __forceinline__ double prior_term(int family, double v, double a, double b, double &grad, double &d_mu,
                                             double &d_log_sd) {
    d_mu = 0.0;
    d_log_sd = 0.0;
    if (family == SKK_PRIOR_NORMAL) {
        const double z = (v - a) / b;
        grad = -z / b;
        d_mu = z / b;
        d_log_sd = z * z - 1.0;
        return -0.5 * z * z - kLogSqrt2Pi - log(b);
    }
    if (family == SKK_PRIOR_HALFNORMAL) {
        const double q = exp(v) / a;
        grad = -q * q + 1.0;
        return -0.5 * q * q + kLogSqrt2OverPi - log(a) + v;
    }
    if (!EXT || family == SKK_PRIOR_EXPONENTIAL) {
        const double sv = exp(v);
        grad = -a * sv + 1.0;
        return log(a) - a * sv + v;
    }
    if constexpr (EXT) {
        // theta-dependent parts only (see prior_is_extended); x = exp(v) for the positive families
        switch (family) {
            case SKK_PRIOR_HALFCAUCHY: {  // -log1p((x/beta)^2) + v
                const double q = exp(v) / a, q2 = q * q;
                grad = -2.0 * q2 / (1.0 + q2) + 1.0;
                return -log1p(q2) + v;
            }
            case SKK_PRIOR_GAMMA: {  // (alpha - 1) v - beta x + v
                const double x = exp(v);
                grad = a - b * x;
                return a * v - b * x;
            }
            case SKK_PRIOR_INVGAMMA: {  // -(alpha + 1) v - beta / x + v
                const double r = b * exp(-v);
                grad = -a + r;
                return -a * v - r;
            }
            case SKK_PRIOR_CAUCHY: {  // -log1p(z^2)
                const double z = (v - a) / b;
                grad = -2.0 * z / (b * (1.0 + z * z));
                return -log1p(z * z);
            }
            case SKK_PRIOR_LAPLACE: {  // -|v - mu| / b
                const double e = v - a;
                grad = e > 0.0 ? -1.0 / b : (e < 0.0 ? 1.0 / b : 0.0);
                return -fabs(e) / b;
            }
            case SKK_PRIOR_LOGNORMAL: {  // -z^2 / 2 with z = (log x - mu) / sigma; -log x and the Jacobian cancel
                const double z = (v - a) / b;
                grad = -z / b;
                return -0.5 * z * z;
            }
            default: {  // SKK_PRIOR_HALFSTUDENTT: -(nu + 1)/2 log1p(x^2 / (nu sigma^2)) + v
                const double x = exp(v) / b, q = x * x / a;
                grad = -(a + 1.0) * q / (1.0 + q) + 1.0;
                return -0.5 * (a + 1.0) * log1p(q) + v;
            }
        }
    }
    grad = 0.0;
    return 0.0;
}

template <>
struct Math<float> {
    static __device__ __forceinline__ float ex2(float x) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float lg2(float x) {
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float exp(float x) { return ex2(x * 1.4426950408889634f); }
    static __device__ __forceinline__ float log1p_of_exp_neg(float t) {  // t in (0, 1]
        return lg2(1.0f + t) * 0.6931471805599453f;
    }
    static __device__ __forceinline__ float log1p_pos(float x) { return ::log1pf(x); }  // x >= 0, any size
    static __device__ __forceinline__ float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
};
template <>
struct Math<double> {
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log1p_of_exp_neg(double t) { return ::log1p(t); }
    static __device__ __forceinline__ double log1p_pos(double x) { return ::log1p(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double abs(double x) { return fabs(x); }
    static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
};

// constants of a likelihood family (NegativeBinomial: a = alpha, b = log(alpha); Gamma, StudentT: a = alpha / nu)
template <typename T>
struct FamilyConst {
    T a, b;
};

// Normal with a log-scale predictor: sigma = exp(zeta).  lp = -r^2 / 2 - zeta, d lp / d eta = r / sigma,
// d lp / d zeta = r^2 - 1, with r = (y - eta) / sigma
template <typename T>
__device__ __forceinline__ void normal_logsigma_row(T y, T eta, T zeta, T &lp, T &d, T &dz) {
    const T inv = Math<T>::exp(-zeta);
    const T r = (y - eta) * inv;
    lp = T(-0.5) * r * r - zeta;
    d = r * inv;
    dz = r * r - T(1);
}

// Student-t with a log-scale predictor (nu constant): lp = -(nu+1)/2 log(1 + r^2/nu) - zeta,
// d lp / d eta = (nu+1) r / (nu + r^2) / sigma,  d lp / d zeta = (nu+1) r^2 / (nu + r^2) - 1
template <typename T>
__device__ __forceinline__ void studentt_logsigma_row(T y, T eta, T zeta, T nu, T &lp, T &d, T &dz) {
    const T inv = Math<T>::exp(-zeta);
    const T r = (y - eta) * inv;
    const T w = (nu + T(1)) * Math<T>::rcp(nu + r * r);
    lp = T(-0.5) * (nu + T(1)) * Math<T>::log1p_pos(r * r * Math<T>::rcp(nu)) - zeta;
    d = w * r * inv;
    dz = w * r * r - T(1);
}

template <typename T, int FAM>
__device__ __forceinline__ void family_row(T y, T eta, T &lp, T &d, const FamilyConst<T> &fc) {
    if constexpr (FAM == SKK_LIK_NORMAL) {
        d = y - eta;
        lp = d * d;
    } else if constexpr (FAM == SKK_LIK_BERNOULLI_LOGIT) {
        const T t = Math<T>::exp(-Math<T>::abs(eta));
        const T inv = Math<T>::rcp(T(1) + t);
        const T sig = eta >= T(0) ? inv : t * inv;
        lp = y * eta - (Math<T>::max(eta, T(0)) + Math<T>::log1p_of_exp_neg(t));
        d = y - sig;
    } else if constexpr (FAM == SKK_LIK_POISSON_LOG) {
        const T mu = Math<T>::exp(eta);
        lp = y * eta - mu;
        d = y - mu;
    } else if constexpr (FAM == SKK_LIK_GAMMA_LOG) {
        // Gamma(alpha, beta = alpha / mu), mu = exp(eta): lp = -alpha (eta + y / mu) (+ constants in logp_const),
        // d lp / d eta = alpha (y / mu - 1)
        const T w = y * Math<T>::exp(-eta);
        lp = -fc.a * (eta + w);
        d = fc.a * (w - T(1));
    } else {
        static_assert(FAM == SKK_LIK_NEGBINOMIAL_LOG, "family without a row function");
        // NegativeBinomial(mu = exp(eta), alpha): with u = eta - log(alpha),
        //   log(alpha + mu) = log(alpha) + softplus(u),   mu / (alpha + mu) = sigmoid(u)
        //   lp = y eta - (y + alpha) log(alpha + mu)  (+ per-row constants in logp_const),  d = y - (y + alpha) sigmoid(u)
        const T u = eta - fc.b;
        const T t = Math<T>::exp(-Math<T>::abs(u));
        const T inv = Math<T>::rcp(T(1) + t);
        const T sig = u >= T(0) ? inv : t * inv;
        const T ya = y + fc.a;
        lp = y * eta - ya * (fc.b + Math<T>::max(u, T(0)) + Math<T>::log1p_of_exp_neg(t));
        d = y - ya * sig;
    }
}

}  // namespace skk
