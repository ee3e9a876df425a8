// Small per-evaluation kernels around the row kernel.  All are O(P) or O(#partial slots), use fp64
// internally and add in a fixed order (serial per thread, or lane-strided + fixed shuffle tree), so
// the whole evaluation is bit-reproducible.  Every kernel is latency-bound (a few dependent loads
// per thread), so the data they walk is flattened at plan time to at most two dependent levels.
//   (the row kernel reads theta itself through the terms' slot -> theta tables; there is no prepare kernel)
//   K3a prior terms  : per theta element, its own prior logp term and gradient, and what it
//                      contributes to the gradient of its prior's mu / sigma hyper-parameters
//   K3b prior links  : per hyper-parameter element, the sum over its children (contiguous)
//   K4a slot totals  : per (term, slot): direct_p + tile partials | sum over CTA partials
//   K4f theta/final  : per theta element: sum of its slots' totals; then either the final
//                      combination (single GPU) or the vector handed to the all-reduce
//   K4b final        : (sharded runs) Normal scale, + priors, + constants -> logp[B], dlogp[B, P]
//   finish           : (single GPU, no element fed by several slots) K4a + K4f + the hyper-parameter sums
//                      in one launch; the prior terms then come from the row kernel's prologue
#pragma once

#include "skk_common.cuh"

namespace skk {

constexpr int kSerialLimit = 48;
constexpr int kSecInFlight = 40;  // CTA partials of a secondary slot that a thread of the finish kernel loads at once  // longer partial lists are summed by a whole warp

// All 32 lanes cooperate on the long lists of the warp's lanes, one list at a time: lane-strided
// partial sums + fixed shuffle tree.  `value(item, e, b)` is the e-th addend of chain b.
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

// sum_{k = first, first + step, ... < n} at(k), in that order, with the loads issued eight at a time: the chain of
// dependent memory latencies is n / (8 step) long instead of n / step (a thread of the finish kernels is latency-bound)
template <typename F>
__device__ __forceinline__ double strided_sum(int64_t first, int64_t step, int64_t n, F at) {
    double s = 0.0;
    for (int64_t k0 = first; k0 < n; k0 += 8 * step) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = k0 + u * step < n ? at(k0 + u * step) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    return s;
}

constexpr int kMaxTerms = 3 * kMaxTermsPerLevel;

// ---- K3 -------------------------------------------------------------------------------------
// Priors, flattened per theta element (structure of arrays) so that a thread needs one level of
// descriptor loads and one level of theta loads.
struct PriorParams {
    const int8_t *family;     // [P] SKK_PRIOR_* (| kPriorMuIsLog)
    const int32_t *a_src;     // [P] theta element holding parameter a (mu), -1: constant
    const int32_t *b_src;     // [P] theta element holding log(parameter b) (sigma), -1: constant
    const double *a_const;    // [P]
    const double *b_const;    // [P]
    const int32_t *e_mu;      // [P] where this element's contribution to its mu parent goes, -1: none
    const int32_t *e_sigma;   // [P] ... to its sigma parent
    const int32_t *parent;    // [n_parents] hyper-parameter elements, those with many children first
    const int64_t *link_ptr;  // [n_parents + 1] ranges of link_val, children in creation order
    int32_t n_parents;
    int32_t n_heavy;          // parents handled by a whole CTA
    int64_t n_params;
    int64_t n_links;
    int32_t batch;
};

constexpr int kPriorThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kPriorThreads)
skk_prior_kernel(const __grid_constant__ PriorParams pp, const T *__restrict__ theta, double *__restrict__ prior_grad,
                 double *__restrict__ prior_lp_part, double *__restrict__ link_val) {
    __shared__ double part[kPriorThreads / 32];
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool valid = j < pp.n_params;
    int fam = 0, a_src = -1, b_src = -1, e_mu = -1, e_sigma = -1;
    double a_c = 0.0, b_c = 1.0;
    if (valid) {
        fam = pp.family[j];
        a_src = pp.a_src[j];
        b_src = pp.b_src[j];
        a_c = pp.a_const[j];
        b_c = pp.b_const[j];
        e_mu = pp.e_mu[j];
        e_sigma = pp.e_sigma[j];
    }
    {
        const int b = blockIdx.y;  // one chain per grid row
        const T *th = theta + (int64_t)b * pp.n_params;
        double lp = 0.0;
        if (valid) {
            const double v = (double)th[j];
            double grad, d_mu, d_log_sd;
            const double a = a_src >= 0 ? prior_mu_of(fam, (double)th[a_src]) : a_c;
            const double sd = b_src >= 0 ? exp((double)th[b_src]) : b_c;
            lp = prior_term(fam & kPriorFamilyMask, v, a, sd, grad, d_mu, d_log_sd);
            // (mu = exp(u) of a positive variable: d/du = d/dmu * mu)
            if (e_mu >= 0) link_val[(int64_t)b * pp.n_links + e_mu] = (fam & kPriorMuIsLog) ? d_mu * a : d_mu;
            if (e_sigma >= 0) link_val[(int64_t)b * pp.n_links + e_sigma] = d_log_sd;
            prior_grad[(int64_t)b * pp.n_params + j] = grad;
        }
        // logp of the priors: fixed tree per CTA; the CTA partials are added in order by the final kernel
        const double s = warp_sum(lp);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kPriorThreads / 32; ++w) tot += part[w];
            prior_lp_part[(int64_t)b * gridDim.x + blockIdx.x] = tot;
        }
    }
}

// One CTA per heavy parent (blockIdx.x < n_heavy), one warp per light parent afterwards.  The
// children's contributions are contiguous in link_val.
static __global__ void skk_prior_link_kernel(const __grid_constant__ PriorParams pp, double *__restrict__ prior_grad,
                                      const double *__restrict__ link_val) {
    __shared__ double part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const bool heavy = (int)blockIdx.x < pp.n_heavy;
    const int64_t p = heavy ? blockIdx.x : pp.n_heavy + ((int64_t)blockIdx.x - pp.n_heavy) * warps + warp;
    if (p >= pp.n_parents) return;
    const int64_t lo = pp.link_ptr[p], hi = pp.link_ptr[p + 1];
    const int64_t first = heavy ? lo + threadIdx.x : lo + lane;
    const int64_t step = heavy ? blockDim.x : 32;
    const int32_t element = pp.parent[p];
    {
        const int b = blockIdx.y;  // one chain per grid row
        const double *val = link_val + (int64_t)b * pp.n_links;
        double acc = 0.0;
        int64_t e = first;
        for (; e + 3 * step < hi; e += 4 * step) {  // four independent loads in flight
            const double v0 = val[e], v1 = val[e + step], v2 = val[e + 2 * step], v3 = val[e + 3 * step];
            acc += v0;
            acc += v1;
            acc += v2;
            acc += v3;
        }
        for (; e < hi; e += step) acc += val[e];
        acc = warp_sum(acc);
        if (heavy) {
            if (lane == 0) part[warp] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                double s = 0.0;
                for (int w = 0; w < warps; ++w) s += part[w];
                prior_grad[(int64_t)b * pp.n_params + element] += s;
            }
        } else if (lane == 0) {
            prior_grad[(int64_t)b * pp.n_params + element] += acc;
        }
    }
}

// ---- K4 -------------------------------------------------------------------------------------
// `totals` is one buffer of BT-wide units: [NG+1 global terms + logp | NP*GP primary slots |
// NS*GS secondary slots], written by the slot-total kernel.
struct ReduceParams {
    // primary slots: tile range and which partial of the first / last tile belongs to the slot
    const int4 *slot_desc;      // [GP] {t_lo, t_hi (t_hi < t_lo: empty), kind of first, kind of last}; kind 0 none, 1 head, 2 tail
    const double *tile_head;    // [tiles][NP][BT]
    const double *tile_tail;
    const double *cta_glob;     // [grid][NG+1][BT]
    const double *cta_sec;      // [grid][NS][GS][BT]
    const double *direct_p;     // [NP][GP][BT] written by the row kernel (groups inside one tile), else zero
    double *totals;
    // theta elements
    const int32_t *single_unit; // [P] unit of `totals` for elements fed by exactly one slot, -1: none, -2: several
    const int32_t *multi_elem;  // [n_multi] elements fed by several slots
    const int64_t *multi_ptr;   // [n_multi + 1]
    const int32_t *multi_unit;  // units, per multi element
    int32_t n_multi;
    int32_t theta_blocks;       // CTAs of the theta kernel that work thread-per-element
    double *lik_vec;            // [B][P+2] (sharded runs)
    int64_t n_params;
    int64_t n_primary, n_secondary;
    int32_t ng, np, ns;
    int32_t bt, b0, b_valid;
    int32_t batch_total;        // chains of the whole evaluation (the last batch tile publishes)
    int32_t grid;               // CTAs of the row kernel
    int32_t blocks_p, blocks_s; // CTAs of the slot-total kernel working on the primary / secondary level
    double logp_const;
    // final combination
    const double *prior_grad;   // [B][P]
    const double *prior_lp_part;  // [B][prior_ctas]
    int32_t prior_ctas;
    int32_t include_priors;
    int64_t n_rows_total;
    int64_t sigma_theta_index;  // -1: constant sigma
    double sigma_const;
    int32_t family;
    // single-GPU fused finish (skk_finish_kernel)
    const int32_t *ti_g[kMaxTermsPerLevel];  // slot -> theta element of every term
    const int32_t *ti_p[kMaxTermsPerLevel];
    const int32_t *ti_s[kMaxTermsPerLevel];
    const int32_t *none_elem;   // theta elements that no term feeds and that have no children
    int32_t n_none, none_blocks;
    // hyper-parameters (elements with children in the priors; no term feeds them on this path)
    const int32_t *parent;      // [n_parents] those with many children first (a CTA each, then a warp each)
    const int64_t *chunk_ptr;   // [n_parents + 1] ranges of link_part (chunks of 32 children)
    const double *link_part;    // [B][n_chunks] written by the row kernel's prior prologue
    int64_t n_chunks;
    int32_t n_parents, n_heavy, parent_blocks;
    int32_t n_lp_parts;         // warp partials of the prior logp that the row kernel's prologue writes per chain
    int32_t multi_blocks;       // CTAs of the finish kernel for the elements fed by several slots (8 each)
    // per term: 0 none of its elements is fed by several slots, 1 all of them are, 2 some (look at single_unit)
    int8_t multi_g[kMaxTermsPerLevel], multi_p[kMaxTermsPerLevel], multi_s[kMaxTermsPerLevel];
    // elements fed by several slots: the raw partials of their primary slots, flattened at plan time --
    // (kind << 30) | index with kind 0: direct_p[index], 1: tile_head[index], 2: tile_tail[index] (x BT + chain)
    const int64_t *multi_src_ptr;  // [n_multi + 1]
    const uint32_t *multi_src;
    const uint8_t *is_parent;      // [P] 1: the element is a hyper-parameter (has children in the priors)
    // optional per-CTA timeline of the finish kernels (skk_trace_enable): [kFinishTraceCtas][kTraceWords], else null
    //   0 %globaltimer at entry   1 ... when the row kernel's results are visible (after griddepcontrol.wait)
    //   2 ... when the CTA's own sums are stored / pushed   3 ... at exit   4 CTA class
    unsigned long long *trace;
};
constexpr int kFinishTraceCtas = 1024;

// blockIdx.x in [0, blocks_p): primary slots, one per thread;  [blocks_p, blocks_p + blocks_s):
// secondary slots, 32 per CTA x 8 ranges of row-kernel CTAs;  last CTA: global terms and logp.
template <int BT>
__global__ void __launch_bounds__(256) skk_slot_total_kernel(const __grid_constant__ ReduceParams rp) {
    const int lane = threadIdx.x & 31;
    const double *direct_p = rp.direct_p;
    double *prim_tot = rp.totals + (int64_t)(rp.ng + 1) * BT;
    double *sec_tot = prim_tot + (int64_t)rp.np * rp.n_primary * BT;
    if ((int)blockIdx.x < rp.blocks_p) {
        const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool valid = slot < rp.n_primary;
        int4 d = make_int4(0, -1, 0, 0);
        if (valid) d = rp.slot_desc[slot];
        const int64_t t_lo = d.x, t_hi = d.y;
        const bool heavy = valid && (t_hi - t_lo + 1 > kSerialLimit);
        // a tile strictly inside the slot's row range lies wholly in the slot: its head partial is
        // the slot's; the first and the last tile contribute what the descriptor says
        auto addend = [&](int4 dd, int64_t tile, int t, int b) -> double {
            const int kind = tile == dd.x ? dd.z : (tile == dd.y ? dd.w : 1);
            const double *src = kind == 1 ? rp.tile_head : (kind == 2 ? rp.tile_tail : nullptr);
            return src ? src[(tile * rp.np + t) * BT + b] : 0.0;
        };
        for (int t = 0; t < rp.np; ++t) {
            if (valid && !heavy && t_hi < t_lo)  // empty group
#pragma unroll
                for (int b = 0; b < BT; ++b) prim_tot[((int64_t)t * rp.n_primary + slot) * BT + b] = 0.0;
            if (valid && !heavy && t_hi >= t_lo) {
                double acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = direct_p[((int64_t)t * rp.n_primary + slot) * BT + b];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] += addend(d, t_lo, t, b);
#pragma unroll 4
                for (int64_t tile = t_lo + 1; tile < t_hi; ++tile)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += rp.tile_head[(tile * rp.np + t) * BT + b];
                if (t_hi > t_lo)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += addend(d, t_hi, t, b);
#pragma unroll
                for (int b = 0; b < BT; ++b) prim_tot[((int64_t)t * rp.n_primary + slot) * BT + b] = acc[b];
            }
            warp_cooperative_lists<BT>(
                heavy, t_lo, t_hi + 1, slot,
                [&](int64_t s, int64_t tile, int b) { return addend(rp.slot_desc[s], tile, t, b); },
                [&](int64_t s, int b, double sum) {
                    const int64_t at = ((int64_t)t * rp.n_primary + s) * BT + b;
                    prim_tot[at] = direct_p[at] + sum;
                });
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
                if (slot < rp.n_secondary) {
                    const double *src = rp.cta_sec + ((int64_t)t * rp.n_secondary + slot) * BT + b;
                    const int64_t stride = (int64_t)rp.ns * rp.n_secondary * BT;
#pragma unroll 4
                    for (int c = c0; c < c1; ++c) acc += src[c * stride];
                }
                part[sub][lane] = acc;
                __syncthreads();
                if (sub == 0 && slot < rp.n_secondary) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) s += part[k][lane];
                    sec_tot[((int64_t)t * rp.n_secondary + slot) * BT + b] = s;
                }
                __syncthreads();
            }
    } else {
        // global terms and the logp partial: one warp per (term, chain)
        const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
        for (int item = warp; item < (rp.ng + 1) * BT; item += warps) {
            double acc = strided_sum(lane, 32, rp.grid, [&](int64_t c) { return rp.cta_glob[c * (rp.ng + 1) * BT + item]; });
            acc = warp_sum(acc);
            if (lane == 0) rp.totals[item] = acc;
        }
    }
}

// ---- one-shot all-reduce over NVLink peer memory ----------------------------------------------
// Every rank owns a receive buffer of 2 (parity) x R slots of `stride` 16-byte lines, mapped into all
// peers with CUDA IPC.  A value travels as one line {low word, flag, high word, flag} with
// flag = evaluation number: a 16-byte store may be split into two 8-byte halves on its way, each
// half carries its own flag, and 8-byte stores are atomic -- so the receiver needs no fence and no
// separate "vector complete" flag; it polls the line until both flags show the current evaluation
// (the LL protocol of NCCL, here written by the kernel that produces the values).
//   push   : the thread that completes the likelihood part of a theta element stores the line into
//            slot [seq & 1][rank] of EVERY peer, straight from its registers;
//   combine: the thread that finishes element j polls the lines of the ranks that feed j and adds
//            them in rank order -- the same order on every rank, so all ranks obtain bit-identical
//            results.  Which ranks feed which element is static: every rank writes its own byte of
//            fed8[j] into all peers when the exchange is set up (skk_peer_init).
// A rank can be at most one evaluation ahead of a peer (it needs that peer's lines to finish), hence
// two parities suffice.  The evaluation number lives on the device (so that the exchange replays inside a CUDA
// graph) and is advanced by the row kernel at the start of every evaluation.  Polls are bounded: on timeout a status word in mapped host memory is set
// and the kernel carries on, so a lost peer can never hang the GPU.
constexpr int kMaxPeers = 8;
struct PeerParams {
    uint4 *peer_lines[kMaxPeers];             // peer_lines[r]: rank r's receive buffer as seen from this rank
    const unsigned long long *fed8;           // [P + 2] own buffer: byte r != 0 when rank r feeds the element
    unsigned long long *seq;                  // this rank's evaluation counter (device)
    int *status;                              // mapped host word, != 0 after a timeout
    int64_t stride;                           // lines per slot
    int32_t rank, nranks;
};

// number of the current evaluation: advanced by the row kernel of the evaluation's first batch tile (one thread, at
// its start), read by the kernels behind it in the stream once the row kernel's writes are visible -- nothing has
// to close an evaluation, so the exchanging kernels end without a fence or a ticket
__device__ __forceinline__ unsigned long long peer_seq(const PeerParams &pp) {
    return *reinterpret_cast<volatile unsigned long long *>(pp.seq);
}

// stores `value` of evaluation `seq` at line `at` of this rank's slot in every peer
__device__ __forceinline__ void peer_push(const PeerParams &pp, unsigned long long seq, int64_t at, double value) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
    const uint32_t flag = (uint32_t)seq, lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
    const int64_t line = ((int64_t)(seq & 1) * pp.nranks + pp.rank) * pp.stride + at;
    for (int r = 0; r < pp.nranks; ++r)
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pp.peer_lines[r] + line), "r"(lo), "r"(flag),
                     "r"(hi), "r"(flag)
                     : "memory");
}

// sum over the ranks that feed line `at` (rank order).  The lines of all those ranks are loaded at once and the
// ones that do not show evaluation `seq` yet are loaded again, so the wait costs one memory latency after the last
// arrival instead of one per rank.
__device__ __forceinline__ double peer_combine(const PeerParams &pp, unsigned long long seq, int64_t at, unsigned long long fed) {
    const uint32_t flag = (uint32_t)seq;
    const uint4 *mine = pp.peer_lines[pp.rank] + (int64_t)(seq & 1) * pp.nranks * pp.stride + at;
    uint32_t lo[kMaxPeers], hi[kMaxPeers];
    unsigned wanted = 0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) {
        lo[r] = hi[r] = 0u;
        if (r < pp.nranks && ((fed >> (8 * r)) & 0xffull) != 0) wanted |= 1u << r;
    }
    unsigned pending = wanted;
    const long long t0 = clock64();
    while (pending != 0) {
        uint32_t l[kMaxPeers], f1[kMaxPeers], h[kMaxPeers], f2[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
            l[r] = f1[r] = h[r] = f2[r] = 0u;
            if ((pending >> r) & 1u)
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(l[r]), "=r"(f1[r]), "=r"(h[r]), "=r"(f2[r])
                             : "l"(mine + (int64_t)r * pp.stride)
                             : "memory");
        }
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (((pending >> r) & 1u) && f1[r] == flag && f2[r] == flag) {
                lo[r] = l[r];
                hi[r] = h[r];
                pending &= ~(1u << r);
            }
        if (pending != 0 && clock64() - t0 > (1ll << 36)) {  // ~35 s: report and go on
            *pp.status = 1;
            break;
        }
    }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
        if ((wanted >> r) & 1u) s += __longlong_as_double((long long)(((unsigned long long)hi[r] << 32) | lo[r]));
    return s;
}

// What the last step does with the likelihood gradient of one theta element.
// MODE 0: single GPU, the final combination; 1: hand lik_vec to ncclAllReduce; 2: store the value
// into this rank's slot of every peer (one-shot all-reduce, see above)
template <typename T, int BT, int MODE>
__device__ __forceinline__ void finish_element(const ReduceParams &rp, const PeerParams &pp, const T *__restrict__ theta,
                                               T *__restrict__ out_grad, int64_t j, const double (&lik)[BT]) {
    const int64_t P = rp.n_params;
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        if (b >= rp.b_valid) continue;
        const int64_t row = rp.b0 + b;
        if constexpr (MODE == 1) {
            rp.lik_vec[row * (P + 2) + j] = lik[b];
        } else if constexpr (MODE == 2) {
            peer_push(pp, peer_seq(pp), row * (P + 2) + j, lik[b]);
        } else {
            double g = lik[b];
            if (rp.family == SKK_LIK_NORMAL) {
                const double sigma =
                    rp.sigma_theta_index >= 0 ? exp((double)theta[row * P + rp.sigma_theta_index]) : rp.sigma_const;
                const double scale = 1.0 / (sigma * sigma);
                g *= scale;
                if (j == rp.sigma_theta_index)
                    g += rp.totals[rp.ng * BT + b] * scale - (double)rp.n_rows_total;  // sum(r^2 - 1)
            }
            if (rp.include_priors) g += rp.prior_grad[row * P + j];
            out_grad[row * P + j] = (T)g;
        }
    }
}

// K4f: CTAs [0, theta_blocks): one thread per theta element (elements fed by at most one slot);
// the following CTAs: one warp per element fed by several slots (lane-strided, fixed tree).
// CTA 0 also produces logp (MODE 0) or the two scalar entries of the exchanged vector.  In MODE 2
// the last CTA to finish raises this rank's flag on every peer.
template <typename T, int BT, int MODE>
__global__ void __launch_bounds__(256)
skk_theta_kernel(const __grid_constant__ ReduceParams rp, const __grid_constant__ PeerParams pp,
                 const T *__restrict__ theta, T *__restrict__ out_logp, T *__restrict__ out_grad) {
    const int lane = threadIdx.x & 31;
    const int64_t P = rp.n_params;
    if ((int)blockIdx.x < rp.theta_blocks) {
        const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        if (j < P) {
            const int unit = rp.single_unit[j];
            if (unit != -2 && (MODE != 2 || unit >= 0)) {   // (over peer memory only the elements this rank feeds travel)
                double lik[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) lik[b] = unit >= 0 ? rp.totals[(int64_t)unit * BT + b] : 0.0;
                finish_element<T, BT, MODE>(rp, pp, theta, out_grad, j, lik);
            }
        }
        if (blockIdx.x == 0) {
            if constexpr (MODE == 0) {
                // logp: CTA partials of the prior kernel in fixed order (strided per thread, fixed tree)
                __shared__ double part[8];
                for (int b = 0; b < rp.b_valid; ++b) {
                    const int64_t row = rp.b0 + b;
                    double s = 0.0;
                    if (rp.include_priors)
                        for (int k = threadIdx.x; k < rp.prior_ctas; k += blockDim.x)
                            s += rp.prior_lp_part[row * rp.prior_ctas + k];
                    s = warp_sum(s);
                    if (lane == 0) part[threadIdx.x >> 5] = s;
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        double tot = 0.0;
                        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
                        const double L = rp.totals[rp.ng * BT + b];
                        double ll = L;
                        if (rp.family == SKK_LIK_NORMAL) {
                            const double sigma = rp.sigma_theta_index >= 0
                                                     ? exp((double)theta[row * P + rp.sigma_theta_index])
                                                     : rp.sigma_const;
                            ll = -0.5 * L / (sigma * sigma) - (double)rp.n_rows_total * (kLogSqrt2Pi + log(sigma));
                        }
                        out_logp[row] = (T)(tot + ll + rp.logp_const);
                    }
                    __syncthreads();
                }
            } else if (threadIdx.x == 0) {
                for (int b = 0; b < rp.b_valid; ++b) {
                    const int64_t row = rp.b0 + b;
                    const double L = rp.totals[rp.ng * BT + b];  // likelihood logp partial
                    if constexpr (MODE == 1) {
                        rp.lik_vec[row * (P + 2) + P] = L;
                        rp.lik_vec[row * (P + 2) + P + 1] = rp.logp_const;
                    } else {
                        peer_push(pp, peer_seq(pp), row * (P + 2) + P, L);
                        peer_push(pp, peer_seq(pp), row * (P + 2) + P + 1, rp.logp_const);
                    }
                }
            }
        }
    } else {
        const int64_t m = ((int64_t)blockIdx.x - rp.theta_blocks) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const bool have = m < rp.n_multi;
        const int64_t lo = have ? rp.multi_ptr[m] : 0, hi = have ? rp.multi_ptr[m + 1] : 0;
        double acc[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] = 0.0;
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const int64_t unit = rp.multi_unit[e];
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[b] += rp.totals[unit * BT + b];
        }
        double lik[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) lik[b] = __shfl_sync(kFull, warp_sum(acc[b]), 0);
        if (lane == 0 && have) finish_element<T, BT, MODE>(rp, pp, theta, out_grad, rp.multi_elem[m], lik);
    }
}

// ---- K4 fused: slot totals -> theta elements -> outputs (or the exchange) in one launch ---------------
// Used unless a term feeds a hyper-parameter (those models run the separate kernels above).
// CTA classes by blockIdx.x: [0, blocks_p) primary slots (one per thread); then blocks_s CTAs of
// secondary slots (32 per CTA x 8 ranges of row-kernel CTAs); one CTA for the global terms, logp and
// the likelihood's sigma; parent_blocks CTAs for the hyper-parameters (sum over their children's
// chunk partials, a CTA or a warp per element); none_blocks CTAs for the elements nothing feeds;
// multi_blocks CTAs for the elements fed by several slots (a term on a parent of a grouping level: a
// warp per element re-adds the raw partials of its slots, lane-strided, fixed tree).  The thread that
// completes the total of a slot finishes the slot's theta element on the spot.  Every thread is
// latency-bound, so whatever it will need (descriptor, element, the element's own prior gradient) is
// loaded up front and the chain of dependent loads stays short.
//   MODE 0 (skk_finish_kernel, single GPU): the element's gradient and logp go to the outputs.
//   MODE 1 / 2 (skk_finish_push_kernel): a row-sharded run -- the element's likelihood part is handed to
//   the exchange instead (1: lik_vec for ncclAllReduce, 2: this rank's slot in every peer, see the
//   one-shot all-reduce above), the hyper-parameter CTAs add their children's chunk partials to
//   prior_grad (read by K4b after the exchange).  Elements that nothing feeds on this rank are sent as
//   zeros in MODE 1 and not at all over peer memory (their entries of the slot are zero since the
//   allocation and are never written).
//   MODE 3 (skk_finish_exchange_kernel): MODE 2 and K4b in one launch -- after the push every CTA waits
//   for the flags of all ranks and then combines a grid-strided slice of theta from the R slots of its
//   own buffer (rank order), adds the priors and writes the outputs.  All CTAs must be co-resident
//   (the host checks the occupancy and falls back to MODE 2 + K4b otherwise).  A sharded evaluation is
//   then two launches: the row kernel (priors in its prologue) and this one.
__device__ __forceinline__ double load_volatile(const double *p) { return *reinterpret_cast<const volatile double *>(p); }

template <typename T, int BT, int MODE>
__device__ __forceinline__ void finish_body(const ReduceParams &rp, const PeerParams &pp, const T *__restrict__ theta,
                                            T *__restrict__ out_logp, T *__restrict__ out_grad) {
    const int lane = threadIdx.x & 31;
    const int64_t P = rp.n_params;
    auto stamp = [&](int word, unsigned long long value) {
        if (rp.trace != nullptr && threadIdx.x == 0 && blockIdx.x < kFinishTraceCtas)
            rp.trace[(int64_t)blockIdx.x * kTraceWords + word] = value;
    };
    stamp(0, global_timer_ns());
    // descriptors are in flight; everything after this reads what the row kernel wrote
    auto dep_wait = [&]() {
        grid_dep_wait();
        stamp(1, global_timer_ns());
    };
    // 1 / sigma^2 of the Normal likelihood per chain (1 for the other families)
    double scale[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        scale[b] = 1.0;
        if (MODE == 0 && rp.family == SKK_LIK_NORMAL && b < rp.b_valid) {
            const double sigma = rp.sigma_theta_index >= 0
                                     ? exp((double)theta[(int64_t)(rp.b0 + b) * P + rp.sigma_theta_index])
                                     : rp.sigma_const;
            scale[b] = 1.0 / (sigma * sigma);
        }
    }
    // gradient of element j = likelihood part (scaled) + own prior term (+ children, hyper-parameters)
    auto prior_of = [&](int64_t j, double (&pg)[BT]) {
#pragma unroll
        for (int b = 0; b < BT; ++b)
            pg[b] = (MODE == 0 && rp.include_priors && j >= 0 && b < rp.b_valid)
                        ? rp.prior_grad[(int64_t)(rp.b0 + b) * P + j] : 0.0;
    };
    // k: a theta element, or (MODE >= 1) P for the logp partial and P + 1 for the constant
    auto store = [&](int64_t k, const double (&lik)[BT], const double (&pg)[BT]) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            if (b >= rp.b_valid) continue;
            const int64_t row = rp.b0 + b;
            if constexpr (MODE == 0) {
                out_grad[row * P + k] = (T)(lik[b] * scale[b] + pg[b]);
            } else if constexpr (MODE == 1) {
                rp.lik_vec[row * (P + 2) + k] = lik[b];
            } else {
                peer_push(pp, peer_seq(pp), row * (P + 2) + k, lik[b]);
            }
        }
    };
    // an element this rank's rows do not feed: its own prior term (MODE 0), a zero for ncclAllReduce, nothing
    // over peer memory
    auto store_unfed = [&](int64_t k, const double (&pg)[BT]) {
        if constexpr (MODE <= 1) {
            double zero[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) zero[b] = 0.0;
            store(k, zero, pg);
        }
    };
    const bool sigma_is_param = rp.family == SKK_LIK_NORMAL && rp.sigma_theta_index >= 0;
    const int b_glob = rp.blocks_p + rp.blocks_s;
    const int b_none = b_glob + 1 + rp.parent_blocks;
    const int b_multi = b_none + rp.none_blocks;
    auto addend = [&](int4 dd, int64_t tile, int t, int b) -> double {
        const int kind = tile == dd.x ? dd.z : (tile == dd.y ? dd.w : 1);
        const double *src = kind == 1 ? rp.tile_head : (kind == 2 ? rp.tile_tail : nullptr);
        return src ? src[(tile * rp.np + t) * BT + b] : 0.0;
    };
    if ((int)blockIdx.x < rp.blocks_p) {
        const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool valid = slot < rp.n_primary;
        int4 d = make_int4(0, -1, 0, 0);
        if (valid) d = rp.slot_desc[slot];
        int32_t elem[kMaxTermsPerLevel];
        bool single[kMaxTermsPerLevel];
#pragma unroll
        for (int t = 0; t < kMaxTermsPerLevel; ++t) {
            elem[t] = (valid && t < rp.np && rp.multi_p[t] != 1) ? rp.ti_p[t][slot] : -1;
            // (an element fed by several slots is finished by its own warp below)
            single[t] = elem[t] >= 0 && (rp.multi_p[t] == 0 || rp.single_unit[elem[t]] != -2);
            if (!single[t]) elem[t] = -1;
        }
        dep_wait();
        double pg[kMaxTermsPerLevel][BT];
#pragma unroll
        for (int t = 0; t < kMaxTermsPerLevel; ++t) prior_of(elem[t], pg[t]);
        const int64_t t_lo = d.x, t_hi = d.y;
        const bool heavy = valid && (t_hi - t_lo + 1 > kSerialLimit);
#pragma unroll
        for (int t = 0; t < kMaxTermsPerLevel; ++t) {
            if (t >= rp.np) break;
            if (valid && !heavy && single[t]) {
                double acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = 0.0;
                if (t_hi >= t_lo) {
#pragma unroll
                    for (int b = 0; b < BT; ++b)
                        acc[b] = rp.direct_p[((int64_t)t * rp.n_primary + slot) * BT + b] + addend(d, t_lo, t, b);
                    // (tiles in order, their loads requested sixteen at a time for a batch tile of one)
                    constexpr int kAhead = BT == 1 ? 16 : 4;
                    for (int64_t tile0 = t_lo + 1; tile0 < t_hi; tile0 += kAhead) {
                        double v[kAhead][BT];
#pragma unroll
                        for (int u = 0; u < kAhead; ++u)
#pragma unroll
                            for (int b = 0; b < BT; ++b)
                                v[u][b] = tile0 + u < t_hi ? rp.tile_head[((tile0 + u) * rp.np + t) * BT + b] : 0.0;
#pragma unroll
                        for (int u = 0; u < kAhead; ++u)
#pragma unroll
                            for (int b = 0; b < BT; ++b) acc[b] += v[u][b];
                    }
                    if (t_hi > t_lo)
#pragma unroll
                        for (int b = 0; b < BT; ++b) acc[b] += addend(d, t_hi, t, b);
                }
                store(elem[t], acc, pg[t]);
            }
            // groups that span many tiles: the whole warp sums one list at a time, then its owner finishes
            unsigned todo = __ballot_sync(kFull, heavy && single[t]);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t s2 = __shfl_sync(kFull, slot, src);
                const int4 dd = rp.slot_desc[s2];
                double acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = 0.0;
                for (int64_t tile = dd.x + lane; tile <= dd.y; tile += 32)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += addend(dd, tile, t, b);
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    acc[b] = __shfl_sync(kFull, warp_sum(acc[b]), 0) +
                             rp.direct_p[((int64_t)t * rp.n_primary + s2) * BT + b];
                if (lane == src) store(elem[t], acc, pg[t]);
            }
        }
    } else if ((int)blockIdx.x < b_glob) {
        __shared__ double part[8][32];
        const int sub = threadIdx.x >> 5;
        const int64_t slot = ((int64_t)blockIdx.x - rp.blocks_p) * 32 + lane;
        const bool valid = slot < rp.n_secondary;
        const int per = (rp.grid + 7) / 8;
        const int c0 = sub * per, c1 = min(rp.grid, c0 + per);
        int32_t elem[kMaxTermsPerLevel];
#pragma unroll
        for (int t = 0; t < kMaxTermsPerLevel; ++t) {
            elem[t] = (sub == 0 && valid && t < rp.ns && rp.multi_s[t] != 1) ? rp.ti_s[t][slot] : -1;
            if (elem[t] >= 0 && rp.multi_s[t] == 2 && rp.single_unit[elem[t]] == -2) elem[t] = -1;
        }
        dep_wait();
        double pg[kMaxTermsPerLevel][BT];
#pragma unroll
        for (int t = 0; t < kMaxTermsPerLevel; ++t) prior_of(elem[t], pg[t]);
#pragma unroll
        for (int t = 0; t < kMaxTermsPerLevel; ++t) {
            if (t >= rp.ns) break;
            double lik[BT];
            for (int b = 0; b < BT; ++b) {
                double acc = 0.0;
                if (valid) {
                    const double *src = rp.cta_sec + ((int64_t)t * rp.n_secondary + slot) * BT + b;
                    const int64_t stride = (int64_t)rp.ns * rp.n_secondary * BT;
                    if (per <= kSecInFlight) {
                        // the usual grid (two CTAs per SM: 37 partials per sub-range): every load is issued before
                        // the first addition, so the sub-range costs one memory latency; same order of additions
                        double v[kSecInFlight];
#pragma unroll
                        for (int k = 0; k < kSecInFlight; ++k) v[k] = c0 + k < c1 ? src[(c0 + k) * stride] : 0.0;
#pragma unroll
                        for (int k = 0; k < kSecInFlight; ++k) acc += v[k];
                    } else {
#pragma unroll 8
                        for (int c = c0; c < c1; ++c) acc += src[c * stride];
                    }
                }
                part[sub][lane] = acc;
                __syncthreads();
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) s += part[k][lane];
                lik[b] = s;
                __syncthreads();
            }
            if (elem[t] >= 0) store(elem[t], lik, pg[t]);
        }
    } else if ((int)blockIdx.x == b_glob) {
        // global terms and the likelihood's logp partial: one warp per (term, chain), CTAs of the row kernel in order
        __shared__ double glob[(kMaxTermsPerLevel + 1) * BT];
        __shared__ double lp_part[8];
        const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
        // (static descriptors of the global terms: in flight before the row kernel's results are waited for)
        int32_t jg = -1;
        if ((int)threadIdx.x < rp.ng) {
            jg = rp.ti_g[threadIdx.x][0];
            const int multi = rp.multi_g[threadIdx.x];
            if (multi == 1 || (multi == 2 && rp.single_unit[jg] == -2)) jg = -1;
        }
        dep_wait();
        for (int item = warp; item < (rp.ng + 1) * BT; item += warps) {
            double acc = 0.0;
            for (int c = lane; c < rp.grid; c += 32) acc += rp.cta_glob[(int64_t)c * (rp.ng + 1) * BT + item];
            acc = warp_sum(acc);
            if (lane == 0) glob[item] = acc;
        }
        // prior logp: the warp partials of the row kernel's prologue, 8 / BT warps per chain
        // (sharded runs: added after the exchange)
        constexpr int WPC = 8 / BT;  // BT is 1 or 4
        if constexpr (MODE == 0) {
            const int chain = warp / WPC, sub = warp % WPC;
            double s = 0.0;
            if (rp.include_priors && chain < rp.b_valid)
                s = strided_sum(sub * 32 + lane, WPC * 32, rp.n_lp_parts,
                                [&](int64_t k) { return rp.prior_lp_part[(int64_t)(rp.b0 + chain) * rp.prior_ctas + k]; });
            s = warp_sum(s);
            if (lane == 0) lp_part[warp] = s;
        }
        __syncthreads();
        if ((int)threadIdx.x < rp.ng) {
            double lik[BT], pg[BT];
            const int32_t j = jg;
#pragma unroll
            for (int b = 0; b < BT; ++b) lik[b] = glob[threadIdx.x * BT + b];
            prior_of(j, pg);
            if (j >= 0 && !(sigma_is_param && j == rp.sigma_theta_index)) store(j, lik, pg);
        }
        if (threadIdx.x == 32 && sigma_is_param) {
            // d/d log(sigma) = sum(r^2) / sigma^2 - N (+ its own prior term; sigma has no children on this path)
            // (sharded runs: nothing here, its part travels as the logp partial)
            double lik[BT], pg[BT];
            prior_of(rp.sigma_theta_index, pg);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                lik[b] = MODE == 0 ? glob[rp.ng * BT + b] : 0.0;
                pg[b] -= (double)rp.n_rows_total;
            }
            if constexpr (MODE == 0) store(rp.sigma_theta_index, lik, pg);
            else store_unfed(rp.sigma_theta_index, pg);
        }
        if constexpr (MODE == 0) {
            if ((int)threadIdx.x >= 64 && (int)threadIdx.x < 64 + BT && (int)threadIdx.x - 64 < rp.b_valid) {
                const int b = threadIdx.x - 64;
                const int64_t row = rp.b0 + b;
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < WPC; ++w) tot += lp_part[b * WPC + w];
                const double L = glob[rp.ng * BT + b];
                double ll = L;
                if (rp.family == SKK_LIK_NORMAL) {
                    const double log_sigma =
                        sigma_is_param ? (double)theta[row * P + rp.sigma_theta_index] : log(rp.sigma_const);
                    double sc = 1.0;
#pragma unroll
                    for (int k = 0; k < BT; ++k) sc = k == b ? scale[k] : sc;  // (no dynamically indexed register array)
                    ll = -0.5 * L * sc - (double)rp.n_rows_total * (kLogSqrt2Pi + log_sigma);
                }
                out_logp[row] = (T)(tot + ll + rp.logp_const);
            }
        } else if (threadIdx.x == 64) {
            double L[BT], c[BT], none[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                L[b] = glob[rp.ng * BT + b];
                c[b] = rp.logp_const;
                none[b] = 0.0;
            }
            store(P, L, none);
            store(P + 1, c, none);
        }
    } else if ((int)blockIdx.x < b_none) {
        // hyper-parameters: own prior term + the sum over the children's chunk partials (contiguous)
        __shared__ double hp_part[8][BT];
        const int warp = threadIdx.x >> 5;
        const int pb = (int)blockIdx.x - b_glob - 1;
        const bool heavy = pb < rp.n_heavy;
        const int64_t q = heavy ? pb : rp.n_heavy + ((int64_t)pb - rp.n_heavy) * 8 + warp;
        const bool active = heavy || q < rp.n_parents;
        int32_t j = 0;
        double pg[BT], sum[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) sum[b] = pg[b] = 0.0;
        int64_t lo = 0, hi = 0;
        if (active) {
            lo = rp.chunk_ptr[q];
            hi = rp.chunk_ptr[q + 1];
            j = rp.parent[q];
        }
        dep_wait();
        if (active) {
            prior_of(j, pg);
            const int64_t first = lo + (heavy ? threadIdx.x : lane), step = heavy ? blockDim.x : 32;
            const double *part = rp.link_part + (int64_t)rp.b0 * rp.n_chunks;
            for (int64_t e = first; e < hi; e += step)
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    if (b < rp.b_valid) sum[b] += part[(int64_t)b * rp.n_chunks + e];
        }
#pragma unroll
        for (int b = 0; b < BT; ++b) sum[b] = warp_sum(sum[b]);
        if (heavy) {
            if (lane == 0)
#pragma unroll
                for (int b = 0; b < BT; ++b) hp_part[warp][b] = sum[b];
            __syncthreads();
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                sum[b] = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) sum[b] += hp_part[w][b];
            }
        }
        if (active && (heavy ? threadIdx.x == 0 : lane == 0)) {
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                if constexpr (MODE == 0) {
                    pg[b] += rp.include_priors ? sum[b] : 0.0;
                } else if constexpr (MODE == 3) {
                    // nothing feeds a hyper-parameter on any rank: own prior term + the children, straight to the output
                    if (b < rp.b_valid)
                        out_grad[(int64_t)(rp.b0 + b) * P + j] =
                            (T)(rp.include_priors ? rp.prior_grad[(int64_t)(rp.b0 + b) * P + j] + sum[b] : 0.0);
                } else {
                    // own term written by the row kernel's prologue; this thread is the element's only other writer
                    if (b < rp.b_valid && rp.include_priors)
                        const_cast<double *>(rp.prior_grad)[(int64_t)(rp.b0 + b) * P + j] += sum[b];
                }
            }
            store_unfed(j, pg);
        }
    } else if ((int)blockIdx.x < b_multi) {
        // elements that no term feeds and that have no children: own prior term only
        const int64_t k = ((int64_t)blockIdx.x - b_none) * blockDim.x + threadIdx.x;
        const int64_t j = k < rp.n_none ? rp.none_elem[k] : -1;
        dep_wait();
        if (j >= 0) {
            if (!(sigma_is_param && j == rp.sigma_theta_index)) {
                double pg[BT];
                prior_of(j, pg);
                store_unfed(j, pg);
            }
        }
    } else {
        // elements fed by several slots: one warp each.  The raw partials of the element's primary slots were
        // flattened at plan time (one index load, one value load, lane-strided); global / secondary units --
        // a block used by terms of different levels -- are walked the long way.
        const int64_t m = ((int64_t)blockIdx.x - b_multi) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const bool have = m < rp.n_multi;
        const int64_t lo = have ? rp.multi_ptr[m] : 0, hi = have ? rp.multi_ptr[m + 1] : 0;
        const int64_t lo2 = have ? rp.multi_src_ptr[m] : 0, hi2 = have ? rp.multi_src_ptr[m + 1] : 0;
        const int32_t j = have ? rp.multi_elem[m] : -1;
        constexpr uint32_t kNoSrc = 0xffffffffu;
        // the list of raw partials is static: its first kMultiAhead * 32 entries are fetched before the row kernel
        // has finished, so that afterwards only the values remain (one memory latency for the usual lengths)
        constexpr int kMultiAhead = BT == 1 ? 16 : 8;
        uint32_t src0[kMultiAhead];
#pragma unroll
        for (int u = 0; u < kMultiAhead; ++u) src0[u] = lo2 + lane + 32 * u < hi2 ? rp.multi_src[lo2 + lane + 32 * u] : kNoSrc;
        dep_wait();
        double pg[BT], acc[BT];
        prior_of(j, pg);
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[b] = 0.0;
        auto value_of = [&](uint32_t src, int b) -> double {
            if (src == kNoSrc) return 0.0;
            const int kind = (int)(src >> 30);
            const int64_t at = (int64_t)(src & 0x3fffffffu) * BT;
            const double *base = kind == 0 ? rp.direct_p : (kind == 1 ? rp.tile_head : rp.tile_tail);
            return base[at + b];
        };
        {
            double v[kMultiAhead][BT];
#pragma unroll
            for (int u = 0; u < kMultiAhead; ++u)
#pragma unroll
                for (int b = 0; b < BT; ++b) v[u][b] = value_of(src0[u], b);
#pragma unroll
            for (int u = 0; u < kMultiAhead; ++u)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] += v[u][b];
        }
        for (int64_t e0 = lo2 + lane + 32 * kMultiAhead; e0 < hi2; e0 += 128) {
            uint32_t src[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) src[u] = e0 + 32 * u < hi2 ? rp.multi_src[e0 + 32 * u] : kNoSrc;
            double v[4][BT];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int b = 0; b < BT; ++b) v[u][b] = value_of(src[u], b);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] += v[u][b];
        }
        const int64_t first_p = rp.ng + 1, first_s = first_p + (int64_t)rp.np * rp.n_primary;
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const int64_t unit = rp.multi_unit[e];
            if (unit < first_p) {
                // a global term: CTA partials of the row kernel in order
                for (int c = 0; c < rp.grid; ++c)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += rp.cta_glob[((int64_t)c * (rp.ng + 1) + unit) * BT + b];
            } else if (unit >= first_s) {
                const int64_t at = unit - first_s;  // t * n_secondary + slot
                const int64_t stride = (int64_t)rp.ns * rp.n_secondary * BT;
                for (int c = 0; c < rp.grid; ++c)
#pragma unroll
                    for (int b = 0; b < BT; ++b) acc[b] += rp.cta_sec[c * stride + at * BT + b];
            }
        }
        double lik[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) lik[b] = warp_sum(acc[b]);
        if (lane == 0 && have) store(j, lik, pg);
    }
    stamp(2, global_timer_ns());
    stamp(4, (unsigned long long)((int)blockIdx.x < rp.blocks_p ? 0 : (int)blockIdx.x < b_glob ? 1 : (int)blockIdx.x == b_glob ? 2
                                  : (int)blockIdx.x < b_none ? 3 : (int)blockIdx.x < b_multi ? 4 : 5));
    if constexpr (MODE == 3) {
        // ---- combine a slice of theta from the R slots of the own buffer (rank order), as the lines arrive ----
        // (hyper-parameters were finished above: nothing feeds them on any rank)
        const unsigned long long seq = peer_seq(pp);
        const bool normal = rp.family == SKK_LIK_NORMAL;
        __shared__ double lp_tail[8];
        __shared__ double lp_pair[2];
        const bool last = blockIdx.x == gridDim.x - 1;  // this CTA also produces logp
        for (int b = 0; b < rp.b_valid; ++b) {
            const int64_t row = rp.b0 + b;
            const int64_t base = row * (P + 2);
            double sc = 1.0, sigma = 1.0;
            if (normal) {
                sigma = rp.sigma_theta_index >= 0 ? exp((double)theta[row * P + rp.sigma_theta_index]) : rp.sigma_const;
                sc = 1.0 / (sigma * sigma);
            }
            if (last) {
                // the priors' part of logp needs nothing from the exchange: summed before the waits below
                double s = 0.0;
                if (rp.include_priors)
                    s = strided_sum(threadIdx.x, blockDim.x, rp.prior_ctas,
                                    [&](int64_t k) { return rp.prior_lp_part[row * rp.prior_ctas + k]; });
                s = warp_sum(s);
                if (lane == 0) lp_tail[threadIdx.x >> 5] = s;
            }
            for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < P; j += (int64_t)gridDim.x * blockDim.x) {
                if (rp.is_parent[j]) continue;
                // (the element's own prior term travels next to the first poll, not behind the last)
                const double own = rp.include_priors ? rp.prior_grad[row * P + j] : 0.0;
                double g = peer_combine(pp, seq, base + j, pp.fed8[j]) * sc;
                if (normal && j == rp.sigma_theta_index)
                    g += peer_combine(pp, seq, base + P, pp.fed8[P]) * sc - (double)rp.n_rows_total;
                g += own;
                out_grad[row * P + j] = (T)g;
            }
            stamp(5, global_timer_ns());
            if (last) {
                // the likelihood's logp partial and the constant: two lanes of one warp poll side by side
                if (threadIdx.x < 2) lp_pair[threadIdx.x] = peer_combine(pp, seq, base + P + threadIdx.x, pp.fed8[P + threadIdx.x]);
                stamp(6, global_timer_ns());
                __syncthreads();
                if (threadIdx.x == 0) {
                    double tot = 0.0;
                    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += lp_tail[w];
                    const double L = lp_pair[0];
                    const double ll = normal ? -0.5 * L * sc - (double)rp.n_rows_total * (kLogSqrt2Pi + log(sigma)) : L;
                    out_logp[row] = (T)(tot + ll + lp_pair[1]);
                }
                __syncthreads();
            }
        }
    }
    stamp(3, global_timer_ns());
}

template <typename T, int BT>
__global__ void __launch_bounds__(256)
skk_finish_kernel(const __grid_constant__ ReduceParams rp, const __grid_constant__ PeerParams pp,
                  const T *__restrict__ theta, T *__restrict__ out_logp, T *__restrict__ out_grad) {
    finish_body<T, BT, 0>(rp, pp, theta, out_logp, out_grad);
}

template <typename T, int BT, int MODE>
__global__ void __launch_bounds__(256)
skk_finish_push_kernel(const __grid_constant__ ReduceParams rp, const __grid_constant__ PeerParams pp) {
    static_assert(MODE == 1 || MODE == 2, "MODE 0 is skk_finish_kernel, MODE 3 skk_finish_exchange_kernel");
    finish_body<T, BT, MODE>(rp, pp, static_cast<const T *>(nullptr), static_cast<T *>(nullptr), static_cast<T *>(nullptr));
}

template <typename T, int BT>
__global__ void __launch_bounds__(256)
skk_finish_exchange_kernel(const __grid_constant__ ReduceParams rp, const __grid_constant__ PeerParams pp,
                           const T *theta, T *out_logp, T *out_grad) {
    finish_body<T, BT, 3>(rp, pp, theta, out_logp, out_grad);
}

// ---- K4b (sharded runs: after the all-reduce) ------------------------------------------------
struct FinalParams {
    const double *lik_vec;      // [B][P+2], all-reduced over ranks
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

// PEER: the likelihood vector is the rank-ordered sum of the peers' slots (one-shot all-reduce
// above); otherwise it is read from lik_vec (single rank, or after ncclAllReduce).
template <typename T, bool PEER>
__global__ void skk_final_kernel(const __grid_constant__ FinalParams fp, const __grid_constant__ PeerParams pp,
                                 const T *__restrict__ theta, T *__restrict__ out_logp, T *__restrict__ out_grad) {
    const int b = blockIdx.y;
    const int64_t P = fp.n_params;
    const double *lik_in = fp.lik_vec + (int64_t)b * (P + 2);
    unsigned long long seq = 0;
    if constexpr (PEER) seq = peer_seq(pp);
    auto lik_at = [&](int64_t k) -> double {
        if constexpr (PEER) {
            return peer_combine(pp, seq, (int64_t)b * (P + 2) + k, pp.fed8[k]);
        } else {
            return lik_in[k];
        }
    };
    double scale = 1.0, sigma = 1.0;
    if (fp.family == SKK_LIK_NORMAL) {
        sigma = fp.sigma_theta_index >= 0 ? exp((double)theta[(int64_t)b * P + fp.sigma_theta_index]) : fp.sigma_const;
        scale = 1.0 / (sigma * sigma);
    }
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < P) {
        double g = lik_at(j) * scale;
        if (fp.family == SKK_LIK_NORMAL && j == fp.sigma_theta_index)
            g += lik_at(P) * scale - (double)fp.n_rows_total;  // sum(r^2 - 1), r = e / sigma
        if (fp.include_priors) g += fp.prior_grad[(int64_t)b * P + j];
        out_grad[(int64_t)b * P + j] = (T)g;
    }
    if (blockIdx.x == 0) {
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
                ll = -0.5 * lik_at(P) * scale - (double)fp.n_rows_total * (kLogSqrt2Pi + log(sigma));
            else
                ll = lik_at(P);
            out_logp[b] = (T)(tot + ll + lik_at(P + 1));
        }
    }
}

// set-up of the exchange: this rank's byte of fed8[j] on every peer (1: the rank feeds element j)
static __global__ void skk_peer_fed_kernel(const __grid_constant__ PeerParams pp, const uint8_t *__restrict__ fed, int64_t n,
                                           int64_t fed8_line_offset) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int r = 0; r < pp.nranks; ++r) {
        uint8_t *dst = reinterpret_cast<uint8_t *>(pp.peer_lines[r] + fed8_line_offset) + j * 8 + pp.rank;
        *dst = fed[j];
    }
}


// Integrator update of the many-chain samplers between two evaluations (sakkara_b200/sampling.py: kick_drift):
//   r[c][j] += a * eps[c] * g[c][j];   q[c][j] += b * eps[c] * inv_mass[j] * r[c][j]      (b == 0: the kick only)
// eps is signed per chain (NUTS extends every chain's trajectory in its own direction).  Elementwise over
// [chains, P], HBM-bound (3 reads + 2 writes per element); blockIdx.y = chain, so no index division.
template <typename T>
__global__ void __launch_bounds__(256)
skk_leapfrog_kernel(T *__restrict__ q, T *__restrict__ r, const T *__restrict__ g, const T *__restrict__ eps,
                    const T *__restrict__ inv_mass, int n_params, T a, T b) {
    const int64_t base = (int64_t)blockIdx.y * n_params;
    const T e = eps[blockIdx.y];
    const T ae = a * e, be = b * e;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_params; j += gridDim.x * blockDim.x) {
        const T rn = r[base + j] + ae * g[base + j];
        r[base + j] = rn;
        if (b != T(0)) q[base + j] += be * inv_mass[j] * rn;
    }
}

}  // namespace skk
