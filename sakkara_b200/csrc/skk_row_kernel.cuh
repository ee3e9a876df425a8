// K1: the row kernel.  One fused pass over the observation rows computes, for BT parameter
// vectors at once, the likelihood's logp and the gradient w.r.t. every term of the linear
// predictor, reduced per group without atomics.
//
// Design:
//  * rows are pre-sorted by the primary grouping level (and by the secondary code inside a primary
//    segment); a tile = 32*R consecutive rows belongs to one warp; a lane owns R consecutive rows.
//  * each warp runs its own TMA pipeline: lane 0 issues one cp.async.bulk per data column into the
//    warp's private shared-memory ring and all lanes wait on the stage's mbarrier -- no
//    __syncthreads in the main loop, loads of `stages-1` tiles are in flight per warp.
//  * a lane walks its R rows serially and merges runs of equal keys in registers.  Reductions
//    across lanes use fixed shuffle trees or __match_any groups summed in lane order, and every
//    partial lands in a slot with exactly one writer, so results are bit-reproducible:
//      global terms / logp : per-lane fp64 accumulators over the whole CTA range -> cta_glob[cta]
//      primary level       : tile inside one segment (the common case) -> one warp sum ->
//                            tile_head[tile]; tile with segment boundaries -> groups that lie inside
//                            the tile go to direct_p[group], the first/last group of the tile to
//                            tile_head/tile_tail.  K4 adds the slots of every group in fixed order.
//      secondary level     : warp-private accumulator table in shared memory (plain
//                            read-modify-write: equal codes are contiguous inside a segment, so
//                            runs that start and end inside a lane are owned by that lane; runs
//                            shared between lanes are combined with warp_match_sum) -> cta_sec[cta].
//  * R is odd so that lane-strided shared-memory reads of 4- and 8-byte columns are conflict free.
#pragma once

#include "skk_common.cuh"

namespace skk {

struct StageLayout {
    int x_stride;  // x column c starts at c * x_stride
    int y_off;
    int off_off;
    int code_off;
    int bytes;  // per warp-stage, multiple of 128
};

template <typename T>
__host__ __device__ inline StageLayout make_stage_layout(int rows_per_tile, int n_x, bool y_is_u8, bool has_offset,
                                                         int code_bytes) {
    StageLayout l{};
    l.x_stride = rows_per_tile * (int)sizeof(T);
    int o = n_x * l.x_stride;
    l.y_off = o;
    o += rows_per_tile * (y_is_u8 ? 1 : (int)sizeof(T));
    l.off_off = o;
    if (has_offset) o += rows_per_tile * (int)sizeof(T);
    l.code_off = o;
    o += rows_per_tile * code_bytes;
    l.bytes = (o + 127) & ~127;
    return l;
}

struct CtaLayout {
    int bars_off;
    int vals_off;     // secondary value table  [NS][GS][BT] of T
    int grads_off;    // secondary gradient tables [W][NS][GS][BT] of T
    int scratch_off;  // per-warp scratch of the straddling-tile path
    int scratch_bytes;
    int stage_off;
    int total;
};

#ifndef SKK_PIPE_ROWS
#define SKK_PIPE_ROWS 4
#endif
// rows per step of the hand-pipelined common path (0: plain loop); fewer where a row holds more
// registers (fp64, batch tile of 4) so that the step's loads stay in registers
constexpr int kPipeRows = SKK_PIPE_ROWS;
template <typename T, int BT>
struct PipeRows {
    static constexpr int value =
        kPipeRows == 0 ? 0 : ((int)sizeof(T) * BT <= 4 ? kPipeRows : ((int)sizeof(T) * BT <= 8 ? 2 : 1));
};

// Layout of a staged tile.  The lane owns R consecutive rows of the tile (li = lane * R ...), but a column stores the
// tile row-step major: the lane's row i at position i * 32 + lane.  At every step the 32 lanes then read 32 consecutive
// elements -- one shared-memory wavefront whatever the element size -- where lane-contiguous storage costs two to
// three (strides of 17 bytes / 34 bytes fold several lanes onto one bank: 65 % excess wavefronts in the c4 profile,
// profiles/r02).  The library permutes the columns accordingly when the plan is created (skk_api.cu: transpose_tiles).
#ifdef SKK_NATURAL_STAGE
constexpr bool kStageTransposed = false;
#else
constexpr bool kStageTransposed = true;
#endif

constexpr int kFoldTiles = 128;    // (power of two) tiles of a warp between two folds of the fp32 accumulator tables into fp64
constexpr int kStagedGroups = 32;  // straddling tiles with up to this many segments are handled on chip

// scratch of one warp: segment offsets [33] | values [32][NP][BT] T | accumulators [32][NP][BT] f64 | touched [32]
template <typename T>
__host__ __device__ inline int warp_scratch_bytes(int np, int bt) {
    if (np == 0) return 0;
    const int vals = (kStagedGroups * np * bt * (int)sizeof(T) + 7) & ~7;
    return ((kStagedGroups + 1) * 8 + vals + kStagedGroups * np * bt * 8 + kStagedGroups + 15) & ~15;
}

template <typename T>
__host__ __device__ inline CtaLayout make_cta_layout(int warps, int stages, int stage_bytes, int ns, int gs, int bt,
                                                     int np) {
    CtaLayout l{};
    int o = 0;
    l.bars_off = o;
    o += warps * stages * 8;
    o = (o + 15) & ~15;
    l.vals_off = o;
    o += ns * gs * bt * (int)sizeof(T);
    o = (o + 15) & ~15;
    l.grads_off = o;
    o += warps * ns * gs * bt * (int)sizeof(T);
    o = (o + 15) & ~15;
    l.scratch_off = o;
    l.scratch_bytes = warp_scratch_bytes<T>(np, bt);
    o += warps * l.scratch_bytes;
    o = (o + 127) & ~127;
    l.stage_off = o;
    o += warps * stages * stage_bytes;
    l.total = o;
    return l;
}

template <int N>
struct AtLeast1 {
    static constexpr int value = N > 0 ? N : 1;
};

// Compile-time description of which terms carry a data column and how y / codes are stored.
// SPEC < 0: everything is read from the kernel parameters at run time (generic kernels, one per
// term-count shape).  SPEC >= 0: bit t (global term t), bit 2+t (primary), bit 4+t (secondary) = the
// term has an x column; bit 6 = y stored as uint8; bit 7 = constant offset column; bit 8 = 16-bit
// secondary codes.  The hot shapes of the BASELINE configurations are compiled with a fixed SPEC,
// which removes the predicated loads and the multiplications by one from the row loop.
template <int SPEC>
struct RowSpec {
    static constexpr bool fixed = SPEC >= 0;
    static constexpr int x_bits = (SPEC & 1) + ((SPEC >> 1) & 1) + ((SPEC >> 2) & 1) + ((SPEC >> 3) & 1) +
                                  ((SPEC >> 4) & 1) + ((SPEC >> 5) & 1);
    template <typename P>
    static __device__ __forceinline__ bool xg(const P &p, int t) { return fixed ? ((SPEC >> t) & 1) : p.xcol_g[t] >= 0; }
    template <typename P>
    static __device__ __forceinline__ bool xp(const P &p, int t) { return fixed ? ((SPEC >> (2 + t)) & 1) : p.xcol_p[t] >= 0; }
    template <typename P>
    static __device__ __forceinline__ bool xs(const P &p, int t) { return fixed ? ((SPEC >> (4 + t)) & 1) : p.xcol_s[t] >= 0; }
    // number of staged x columns: with a fixed SPEC every term that has a column has its own (the host
    // falls back to the generic kernel otherwise)
    template <typename P>
    static __device__ __forceinline__ int n_x(const P &p) { return fixed ? x_bits : p.n_x; }
    template <typename P>
    static __device__ __forceinline__ bool y_u8(const P &p) { return fixed ? ((SPEC >> 6) & 1) : p.y8 != nullptr; }
    template <typename P>
    static __device__ __forceinline__ bool has_off(const P &p) { return fixed ? ((SPEC >> 7) & 1) : p.eta_offset != nullptr; }
    template <typename P>
    static __device__ __forceinline__ bool code16(const P &p) { return fixed ? ((SPEC >> 8) & 1) : p.sec_code_bytes == 2; }
};

// ---------------------------------------------------------------------------------------------
template <typename T, int FAM, int BT, int NG, int NP, int NS, int R, int SPEC = -1>
struct RowKernel {
    using Spec = RowSpec<SPEC>;
    static constexpr int WT = 32 * R;  // rows per tile
    static constexpr int NGa = AtLeast1<NG>::value;
    static constexpr int NPa = AtLeast1<NP>::value;
    static constexpr int NSa = AtLeast1<NS>::value;

    struct Tile {
        const unsigned char *stage;
        StageLayout sl;
        int tile;
        int kf, kl;
    };

    // position of (term t, code, chain b) in the shared-memory tables of the secondary level
    static __device__ __forceinline__ int s_at(const RowParams<T> &p, int t, int code, int b) {
        if constexpr (NS <= 1)
            return code * BT + b;  // one term: no multiplication by the table length
        else
            return (t * p.n_secondary + code) * BT + b;
    }

    // theta[chain b of the batch tile][element]; chains beyond b_valid repeat chain 0 (results discarded)
    static __device__ __forceinline__ T theta_at(const RowParams<T> &p, int b, int element) {
        return p.theta[(int64_t)(b < p.b_valid ? b : 0) * p.n_params + element];
    }

    // eta = offset + sum of the terms.  With a fixed SPEC the terms without a data column are added
    // first and those with one are fused on top (eta = fma(value, x, eta)), and nothing is added to a
    // literal zero: one instruction per row less than the generic order below.
    static __device__ __forceinline__ T eta_of(const RowParams<T> &p, T off, const T (&vgb)[NGa], const T (&xg)[NGa],
                                               const T (&vpb)[NPa], const T (&xp)[NPa], const T (&vsb)[NSa],
                                               const T (&xs)[NSa]) {
        if constexpr (Spec::fixed) {
            T eta = off;
            bool have = Spec::has_off(p);
            auto add = [&](T v) {
                eta = have ? eta + v : v;
                have = true;
            };
            auto fuse = [&](T v, T x) {
                eta = have ? v * x + eta : v * x;
                have = true;
            };
#pragma unroll
            for (int t = 0; t < NG; ++t)
                if (!Spec::xg(p, t)) add(vgb[t]);
#pragma unroll
            for (int t = 0; t < NP; ++t)
                if (!Spec::xp(p, t)) add(vpb[t]);
#pragma unroll
            for (int t = 0; t < NS; ++t)
                if (!Spec::xs(p, t)) add(vsb[t]);
#pragma unroll
            for (int t = 0; t < NG; ++t)
                if (Spec::xg(p, t)) fuse(vgb[t], xg[t]);
#pragma unroll
            for (int t = 0; t < NP; ++t)
                if (Spec::xp(p, t)) fuse(vpb[t], xp[t]);
#pragma unroll
            for (int t = 0; t < NS; ++t)
                if (Spec::xs(p, t)) fuse(vsb[t], xs[t]);
            return have ? eta : T(0);
        } else {
            T eta = off;
#pragma unroll
            for (int t = 0; t < NG; ++t) eta += vgb[t] * xg[t];
#pragma unroll
            for (int t = 0; t < NP; ++t) eta += vpb[t] * xp[t];
#pragma unroll
            for (int t = 0; t < NS; ++t) eta += vsb[t] * xs[t];
            return eta;
        }
    }

    // SKK_LIK_NORMAL_LOGSIGMA: a term whose xcol is SKK_XCOL_LOG_SIGMA belongs to zeta = log(sigma), the others to eta
    static constexpr bool kLogSigma = FAM == SKK_LIK_NORMAL_LOGSIGMA || FAM == SKK_LIK_STUDENTT_LOGSIGMA;
    static __device__ __forceinline__ bool zg(const RowParams<T> &p, int t) { return kLogSigma && p.xcol_g[t] == SKK_XCOL_LOG_SIGMA; }
    static __device__ __forceinline__ bool zp(const RowParams<T> &p, int t) { return kLogSigma && p.xcol_p[t] == SKK_XCOL_LOG_SIGMA; }
    static __device__ __forceinline__ bool zs(const RowParams<T> &p, int t) { return kLogSigma && p.xcol_s[t] == SKK_XCOL_LOG_SIGMA; }
    static __device__ __forceinline__ T zeta_of(const RowParams<T> &p, const T (&vgb)[NGa], const T (&vpb)[NPa],
                                                const T (&vsb)[NSa]) {
        T zeta = T(0);
#pragma unroll
        for (int t = 0; t < NG; ++t) zeta += zg(p, t) ? vgb[t] : T(0);
#pragma unroll
        for (int t = 0; t < NP; ++t) zeta += zp(p, t) ? vpb[t] : T(0);
#pragma unroll
        for (int t = 0; t < NS; ++t) zeta += zs(p, t) ? vsb[t] : T(0);
        return zeta;
    }
    // one row: eta (without the zeta terms), the family's logp term and the derivatives, accumulated per term
    static __device__ __forceinline__ void row_terms(const RowParams<T> &p, const FamilyConst<T> &fc, T y, T off,
                                                     const T (&vgb)[NGa], const T (&xg)[NGa], const T (&vpb)[NPa],
                                                     const T (&xp)[NPa], const T (&vsb)[NSa], const T (&xs)[NSa],
                                                     T &lp, T (&dg)[NGa], T (&dp)[NPa], T (&ds)[NSa]) {
        if constexpr (kLogSigma) {
            T eg[NGa], ep[NPa], es[NSa];   // values of the eta terms only
#pragma unroll
            for (int t = 0; t < NGa; ++t) eg[t] = (t < NG && zg(p, t)) ? T(0) : vgb[t];
#pragma unroll
            for (int t = 0; t < NPa; ++t) ep[t] = (t < NP && zp(p, t)) ? T(0) : vpb[t];
#pragma unroll
            for (int t = 0; t < NSa; ++t) es[t] = (t < NS && zs(p, t)) ? T(0) : vsb[t];
            const T eta = eta_of(p, off, eg, xg, ep, xp, es, xs);
            const T zeta = zeta_of(p, vgb, vpb, vsb);
            T d, dz;
            if constexpr (FAM == SKK_LIK_STUDENTT_LOGSIGMA)
                studentt_logsigma_row<T>(y, eta, zeta, fc.a, lp, d, dz);
            else
                normal_logsigma_row<T>(y, eta, zeta, lp, d, dz);
#pragma unroll
            for (int t = 0; t < NGa; ++t) dg[t] = (t < NG && zg(p, t)) ? dz : d * xg[t];
#pragma unroll
            for (int t = 0; t < NPa; ++t) dp[t] = (t < NP && zp(p, t)) ? dz : d * xp[t];
#pragma unroll
            for (int t = 0; t < NSa; ++t) ds[t] = (t < NS && zs(p, t)) ? dz : d * xs[t];
        } else {
            const T eta = eta_of(p, off, vgb, xg, vpb, xp, vsb, xs);
            T d;
            family_row<T, FAM>(y, eta, lp, d, fc);
#pragma unroll
            for (int t = 0; t < NGa; ++t) dg[t] = d * xg[t];
#pragma unroll
            for (int t = 0; t < NPa; ++t) dp[t] = d * xp[t];
#pragma unroll
            for (int t = 0; t < NSa; ++t) ds[t] = d * xs[t];
        }
    }

    // -----------------------------------------------------------------------------------------
    // FASTP: no per-row segment bookkeeping (the tile lies in one segment, or -- MID -- in exactly two:
    // rows from `split` on belong to the second group, whose values are vp_second)
    template <bool FASTP, bool FULL, bool MID = false>
    static __device__ __forceinline__ void process(const RowParams<T> &p, const Tile &tl, const T (&vp_first)[NPa][BT],
                                                   const T (&vp_second)[NPa][BT], int split,
                                                   const int (&staged_element)[NPa],
                                                   const T (&vg)[NGa][BT], double (&accG64)[NG + 1][BT],
                                                   const T *__restrict__ valS, T *__restrict__ gradS, uint32_t gradS_addr,
                                                   unsigned char *scratch, int lane) {
        const FamilyConst<T> fc{p.fam_a, p.fam_b};
        const int li = lane * R;
        // where the lane's row i sits in every staged column: at0 + i * kLaneStep (see kStageTransposed)
        const int at0 = kStageTransposed ? lane : li;
        constexpr int kLaneStep = kStageTransposed ? 32 : 1;
        const int64_t row0 = (int64_t)tl.tile * WT + li;
        const int64_t left = p.n_rows - row0;
        const int nvalid = FULL ? R : (left >= R ? R : (left > 0 ? (int)left : 0));

        // data columns of the terms (x == 1 when the term has no column)
        const T *pxg[NGa], *pxp[NPa], *pxs[NSa];
#pragma unroll
        for (int t = 0; t < NGa; ++t)
            pxg[t] = (t < NG && Spec::xg(p, t))
                         ? reinterpret_cast<const T *>(tl.stage + p.xcol_g[t] * tl.sl.x_stride) + at0 : nullptr;
#pragma unroll
        for (int t = 0; t < NPa; ++t)
            pxp[t] = (t < NP && Spec::xp(p, t))
                         ? reinterpret_cast<const T *>(tl.stage + p.xcol_p[t] * tl.sl.x_stride) + at0 : nullptr;
#pragma unroll
        for (int t = 0; t < NSa; ++t)
            pxs[t] = (t < NS && Spec::xs(p, t))
                         ? reinterpret_cast<const T *>(tl.stage + p.xcol_s[t] * tl.sl.x_stride) + at0 : nullptr;
        const T *sy = reinterpret_cast<const T *>(tl.stage + tl.sl.y_off) + at0;
        const uint8_t *sy8 = reinterpret_cast<const uint8_t *>(tl.stage + tl.sl.y_off) + at0;
        const T *soff = reinterpret_cast<const T *>(tl.stage + tl.sl.off_off) + at0;
        const uint16_t *sc16 = reinterpret_cast<const uint16_t *>(tl.stage + tl.sl.code_off) + at0;
        const int32_t *sc32 = reinterpret_cast<const int32_t *>(tl.stage + tl.sl.code_off) + at0;
        const bool y_u8 = Spec::y_u8(p);
        const bool has_off = Spec::has_off(p);
        const bool code16 = Spec::code16(p);

        // per-span accumulators in T; folded into fp64 at run / span boundaries
        T accL[BT], accG[NGa][BT], accP[NPa][BT], accS[NSa][BT], accP2[NPa][BT];
#pragma unroll
        for (int t = 0; t < NPa; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) accP2[t][b] = T(0);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            accL[b] = T(0);
#pragma unroll
            for (int t = 0; t < NGa; ++t) accG[t][b] = T(0);
#pragma unroll
            for (int t = 0; t < NPa; ++t) accP[t][b] = T(0);
#pragma unroll
            for (int t = 0; t < NSa; ++t) accS[t][b] = T(0);
        }

        // ---- primary level state ----------------------------------------------------------------
        T vp[NPa][BT];
#pragma unroll
        for (int t = 0; t < NPa; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) vp[t][b] = vp_first[t][b];
        int kcur = tl.kf;
        int64_t nextb = INT64_MAX;
        bool p_changed = false, p_have_head = false, p_inside = false;
        int p_head_key = 0;
        double p_head[NPa * BT];
        // Straddling tiles: the tile's segment offsets, values and accumulators are staged in the warp's
        // shared-memory scratch (one coalesced global load each) so that the per-lane search, the value
        // reloads at boundaries and the cross-lane combine never wait on global memory.
        const int ng = tl.kl - tl.kf + 1;
        const bool staged = ng <= kStagedGroups;
        int64_t *s_off = reinterpret_cast<int64_t *>(scratch);
        T *s_val = reinterpret_cast<T *>(scratch + (kStagedGroups + 1) * 8);
        double *s_acc = reinterpret_cast<double *>(scratch + (kStagedGroups + 1) * 8 +
                                                   ((kStagedGroups * NP * BT * (int)sizeof(T) + 7) & ~7));
        unsigned char *s_touch = reinterpret_cast<unsigned char *>(s_acc + kStagedGroups * NP * BT);
        auto seg_off = [&](int k) -> int64_t { return staged ? s_off[k - tl.kf] : p.seg_offsets[k]; };
        auto load_vals = [&](int k) {
#pragma unroll
            for (int t = 0; t < NP; ++t)
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    vp[t][b] = staged ? s_val[((k - tl.kf) * NP + t) * BT + b] : theta_at(p, b, p.ti_p[t][k]);
        };
        if constexpr (NP > 0 && !FASTP) {
            if (staged) {
                for (int j = lane; j <= ng; j += 32) s_off[j] = p.seg_offsets[tl.kf + j];
                if (lane < ng) {
#pragma unroll
                    for (int t = 0; t < NP; ++t)
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            s_val[(lane * NP + t) * BT + b] = theta_at(p, b, staged_element[t]);
                            s_acc[(lane * NP + t) * BT + b] = 0.0;
                        }
                    s_touch[lane] = 0;
                }
                __syncwarp();
            } else {
                // more groups than the scratch holds: partial sums are accumulated in direct_p itself,
                // so the tile's inner groups are cleared first (nothing else ever writes them)
                for (int k = tl.kf + 1 + lane; k < tl.kl; k += 32)
#pragma unroll
                    for (int t = 0; t < NP; ++t)
#pragma unroll
                        for (int b = 0; b < BT; ++b) p.direct_p[((int64_t)t * p.n_primary + k) * BT + b] = 0.0;
                __syncwarp();
            }
            if (nvalid > 0) {
                // segment of the lane's first row: largest k in [kf, kl] with seg_off(k) <= row0
                int lo = tl.kf, hi = tl.kl;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (seg_off(mid) <= row0) lo = mid; else hi = mid - 1;
                }
                kcur = lo;
                nextb = seg_off(kcur + 1);
                p_inside = seg_off(kcur) >= row0;
                load_vals(kcur);
            }
        }

        // ---- secondary level state --------------------------------------------------------------
        int scur = 0;
        bool s_changed = false;
        int s_head_key = 0;
        double s_head[NSa * BT];
        T vs[NSa][BT];
#pragma unroll
        for (int t = 0; t < NSa; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) vs[t][b] = T(0);
        if constexpr (NS > 0) {
            if (nvalid > 0) scur = code16 ? (int)sc16[0] : sc32[0];
        }

        uint32_t table_token = 0;  // ties the in-loop stores to the accumulator table to the fence after the loop
        // ---- the rows ---------------------------------------------------------------------------
        if constexpr (NS > 0 && FASTP && FULL && !MID && PipeRows<T, BT>::value > 0) {
            // The common tile of a model with a secondary level (whole tile inside one primary segment,
            // all rows valid), software-pipelined by hand.  The hardware keeps shared-memory accesses in
            // program order and the assembler cannot prove that the accumulator table aliases nothing
            // else, so in the plain loop below every row's loads wait behind the previous row's
            // load -> add -> store of its table entry.  Here the rows go in steps of kPipeRows: the
            // loads of the next step (data, gathered values, table entries) are issued before the
            // stores of the current one, and within a step nothing is stored before everything is loaded.
            // A row's table entry belongs to the run that ends in front of it, i.e. to the previous row's
            // code; it is only stored when the code changes (same arithmetic and order as the loop below).
            constexpr int K = PipeRows<T, BT>::value;
            int code[R];
#pragma unroll
            for (int i = 0; i < R; ++i) code[i] = code16 ? (int)sc16[i * kLaneStep] : sc32[i * kLaneStep];
            struct Raw {
                T y, off, xg[NGa], xp[NPa], xs[NSa], vs[NSa][BT], tab[NSa][BT];
            };
            auto load_raw = [&](int i, Raw &r) {
                r.y = y_u8 ? (T)sy8[i * kLaneStep] : sy[i * kLaneStep];
                r.off = has_off ? soff[i * kLaneStep] : T(0);
#pragma unroll
                for (int t = 0; t < NG; ++t) r.xg[t] = Spec::xg(p, t) ? pxg[t][i * kLaneStep] : T(1);
#pragma unroll
                for (int t = 0; t < NP; ++t) r.xp[t] = Spec::xp(p, t) ? pxp[t][i * kLaneStep] : T(1);
#pragma unroll
                for (int t = 0; t < NS; ++t) r.xs[t] = Spec::xs(p, t) ? pxs[t][i * kLaneStep] : T(1);
#pragma unroll
                for (int t = 0; t < NS; ++t)
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        r.vs[t][b] = valS[s_at(p, t, code[i], b)];
                        // (only the lanes whose run ends here take part in the gather: fewer bank conflicts)
                        r.tab[t][b] = i > 0 ? lds_free_if(code[i] != code[i - 1],
                                                          gradS_addr + (uint32_t)s_at(p, t, code[i - 1], b) * (uint32_t)sizeof(T), T(0))
                                            : T(0);
                    }
            };
            Raw cur[K], nxt[K];
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (k < R) load_raw(k, cur[k]);
#pragma unroll
            for (int i0 = 0; i0 < R; i0 += K) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (i0 + K + k < R) load_raw(i0 + K + k, nxt[k]);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int i = i0 + k;
                    if (i < R) {
                        const Raw &r = cur[k];
                        if (i > 0) {
                            const bool change = code[i] != code[i - 1];
#pragma unroll
                            for (int t = 0; t < NS; ++t)
#pragma unroll
                                for (int b = 0; b < BT; ++b) {
                                    const uint32_t slot =
                                        gradS_addr + (uint32_t)s_at(p, t, code[i - 1], b) * (uint32_t)sizeof(T);
                                    sts_free_if(change, slot, r.tab[t][b] + accS[t][b], table_token);
                                    if (change) accS[t][b] = T(0);
                                }
                        }
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            T vgb[NGa], vpb[NPa], vsb[NSa];
#pragma unroll
                            for (int t = 0; t < NGa; ++t) vgb[t] = vg[t][b];
#pragma unroll
                            for (int t = 0; t < NPa; ++t) vpb[t] = vp[t][b];
#pragma unroll
                            for (int t = 0; t < NSa; ++t) vsb[t] = r.vs[t][b];
                            if constexpr (kLogSigma) {
                                T lp, dg[NGa], dp[NPa], ds[NSa];
                                row_terms(p, fc, r.y, r.off, vgb, r.xg, vpb, r.xp, vsb, r.xs, lp, dg, dp, ds);
                                accL[b] += lp;
#pragma unroll
                                for (int t = 0; t < NG; ++t) accG[t][b] += dg[t];
#pragma unroll
                                for (int t = 0; t < NP; ++t) accP[t][b] += dp[t];
#pragma unroll
                                for (int t = 0; t < NS; ++t) accS[t][b] += ds[t];
                            } else {
                            const T eta = eta_of(p, r.off, vgb, r.xg, vpb, r.xp, vsb, r.xs);
                            T lp, d;
                            family_row<T, FAM>(r.y, eta, lp, d, fc);
                            accL[b] += lp;
#pragma unroll
                            for (int t = 0; t < NG; ++t) accG[t][b] += d * r.xg[t];
#pragma unroll
                            for (int t = 0; t < NP; ++t) accP[t][b] += d * r.xp[t];
#pragma unroll
                            for (int t = 0; t < NS; ++t) accS[t][b] += d * r.xs[t];
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) cur[k] = nxt[k];
            }
            scur = code[R - 1];
        } else
        // Only the common path (tile inside one segment, all rows valid) is unrolled; the other
        // variants stay rolled to keep the kernel inside the instruction cache.  (Round 2: two-group tiles whose
        // parts share no secondary code were tried on the pipelined path as well -- 150.1 vs 149.2 us on c4, same
        // box, profiles/r02/p_ab_two_group_pipe.log -- and stay here.)
#pragma unroll(FASTP && FULL && !(MID && NS > 0) ? R : 1)
        for (int i = 0; i < R; ++i) {
            const bool valid = FULL || i < nvalid;

            if constexpr (NP > 0 && !FASTP) {
                // segment boundaries inside the lane's span (tiles that straddle segments only)
                while (valid && row0 + i >= nextb) {
                    if (p_inside) {
                        // the group started and ended inside this lane: it has no other writer
#pragma unroll
                        for (int t = 0; t < NP; ++t)
#pragma unroll
                            for (int b = 0; b < BT; ++b)
                                p.direct_p[((int64_t)t * p.n_primary + kcur) * BT + b] = (double)accP[t][b];
                    } else {
                        p_have_head = true;
                        p_head_key = kcur;
#pragma unroll
                        for (int t = 0; t < NP; ++t)
#pragma unroll
                            for (int b = 0; b < BT; ++b) p_head[t * BT + b] = (double)accP[t][b];
                    }
                    p_changed = true;
                    p_inside = true;
                    ++kcur;
                    nextb = seg_off(kcur + 1);
#pragma unroll
                    for (int t = 0; t < NP; ++t)
#pragma unroll
                        for (int b = 0; b < BT; ++b) accP[t][b] = T(0);
                    load_vals(kcur);
                }
                if constexpr (NS > 0) __syncwarp();
            }

            if constexpr (NS > 0) {
                const int c = code16 ? (int)sc16[i * kLaneStep] : sc32[i * kLaneStep];
                const bool change = valid && (c != scur);
                if constexpr (FASTP && MID) {
                    // Two groups in the tile: inside either one the codes ascend, so among the runs that
                    // end in the first group (or at the split) no two lanes hold the same code, and
                    // likewise in the second group -- but a code may end a run on both sides in the same
                    // step, and a lane that holds the split may meet a code twice.  Hence two ordered
                    // read-modify-writes per row, one per side (the run that ends here lies where the
                    // previous row lies).
                    const bool in_second = li + i > split;
#pragma unroll
                    for (int side = 0; side < 2; ++side) {
#pragma unroll
                        for (int t = 0; t < NS; ++t)
#pragma unroll
                            for (int b = 0; b < BT; ++b) {
                                const uint32_t slot = gradS_addr + (uint32_t)s_at(p, t, scur, b) * (uint32_t)sizeof(T);
                                const T updated = lds(slot, T(0)) + accS[t][b];
                                if (change && in_second == (side == 1)) sts(slot, updated);
                            }
                        __syncwarp();
                    }
                } else if constexpr (FASTP) {
                    // Inside one segment equal codes are contiguous, so the run that ends here ends in
                    // this lane only: its sum goes to the table with a plain read-modify-write (lanes
                    // that hold the same code in their *last* run add theirs after the loop).  Written
                    // without branches -- load and add always, store under the predicate.
#pragma unroll
                    for (int t = 0; t < NS; ++t)
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            const uint32_t slot = gradS_addr + (uint32_t)s_at(p, t, scur, b) * (uint32_t)sizeof(T);
                            const T updated = lds_free_if(change, slot, T(0)) + accS[t][b];
                            sts_free_if(change, slot, updated, table_token);
                        }
                } else {
                    // the tile straddles primary segments: a code may occur in several runs of the
                    // tile, so lanes flushing the same entry at the same step are combined first
                    const bool flush = change && s_changed;
                    double v[NSa * BT];
#pragma unroll
                    for (int t = 0; t < NS; ++t)
#pragma unroll
                        for (int b = 0; b < BT; ++b) v[t * BT + b] = (double)accS[t][b];
                    if (warp_match_sum<NSa * BT>(flush, scur, v)) {
#pragma unroll
                        for (int t = 0; t < NS; ++t)
#pragma unroll
                            for (int b = 0; b < BT; ++b)
                                gradS[s_at(p, t, scur, b)] += (T)v[t * BT + b];
                    }
                    __syncwarp();
                    if (change && !s_changed) {
                        s_head_key = scur;
#pragma unroll
                        for (int t = 0; t < NS; ++t)
#pragma unroll
                            for (int b = 0; b < BT; ++b) s_head[t * BT + b] = (double)accS[t][b];
                    }
                }
                if (change) {
                    s_changed = true;
                    scur = c;
#pragma unroll
                    for (int t = 0; t < NS; ++t)
#pragma unroll
                        for (int b = 0; b < BT; ++b) accS[t][b] = T(0);
                }
                // the value of the row's own code: independent of the run bookkeeping above
#pragma unroll
                for (int t = 0; t < NS; ++t)
#pragma unroll
                    for (int b = 0; b < BT; ++b) vs[t][b] = valS[s_at(p, t, c, b)];
            }

            // data of the row
            const T yv = y_u8 ? (T)sy8[i * kLaneStep] : sy[i * kLaneStep];
            T xg[NGa], xp[NPa], xs[NSa];
#pragma unroll
            for (int t = 0; t < NG; ++t) xg[t] = Spec::xg(p, t) ? pxg[t][i * kLaneStep] : T(1);
#pragma unroll
            for (int t = 0; t < NP; ++t) xp[t] = Spec::xp(p, t) ? pxp[t][i * kLaneStep] : T(1);
#pragma unroll
            for (int t = 0; t < NS; ++t) xs[t] = Spec::xs(p, t) ? pxs[t][i * kLaneStep] : T(1);
            const T off = has_off ? soff[i * kLaneStep] : T(0);

#pragma unroll
            for (int b = 0; b < BT; ++b) {
                T vgb[NGa], vpb[NPa], vsb[NSa];
#pragma unroll
                for (int t = 0; t < NGa; ++t) vgb[t] = vg[t][b];
#pragma unroll
                for (int t = 0; t < NPa; ++t) vpb[t] = (MID && li + i >= split) ? vp_second[t][b] : vp[t][b];
#pragma unroll
                for (int t = 0; t < NSa; ++t) vsb[t] = vs[t][b];
                if constexpr (kLogSigma) {
                    T lp, dg[NGa], dp[NPa], ds[NSa];
                    row_terms(p, fc, yv, off, vgb, xg, vpb, xp, vsb, xs, lp, dg, dp, ds);
                    accL[b] += valid ? lp : T(0);
#pragma unroll
                    for (int t = 0; t < NG; ++t) accG[t][b] += valid ? dg[t] : T(0);
#pragma unroll
                    for (int t = 0; t < NP; ++t) {
                        const T v = valid ? dp[t] : T(0);
                        if constexpr (MID) {
                            const bool second = li + i >= split;
                            accP[t][b] += second ? T(0) : v;
                            accP2[t][b] += second ? v : T(0);
                        } else {
                            accP[t][b] += v;
                        }
                    }
#pragma unroll
                    for (int t = 0; t < NS; ++t) accS[t][b] += valid ? ds[t] : T(0);
                } else {
                const T eta = eta_of(p, off, vgb, xg, vpb, xp, vsb, xs);
                T lp, d;
                family_row<T, FAM>(yv, eta, lp, d, fc);
                if (!valid) {
                    lp = T(0);
                    d = T(0);
                }
                accL[b] += lp;
#pragma unroll
                for (int t = 0; t < NG; ++t) accG[t][b] += d * xg[t];
#pragma unroll
                for (int t = 0; t < NP; ++t) {
                    if constexpr (MID) {
                        const bool second = li + i >= split;
                        accP[t][b] += second ? T(0) : d * xp[t];
                        accP2[t][b] += second ? d * xp[t] : T(0);
                    } else {
                        accP[t][b] += d * xp[t];
                    }
                }
#pragma unroll
                for (int t = 0; t < NS; ++t) accS[t][b] += d * xs[t];
                }
            }
        }

        if constexpr (NS > 0 && FASTP) table_stores_done(table_token);

        // ---- global terms and logp: lane-private fp64 accumulators over the CTA's whole range ------
#pragma unroll
        for (int b = 0; b < BT; ++b) {
#pragma unroll
            for (int t = 0; t < NG; ++t) accG64[t][b] += (double)accG[t][b];
            accG64[NG][b] += (double)accL[b];
        }

        // ---- primary level ------------------------------------------------------------------------
        if constexpr (NP > 0) {
            if constexpr (FASTP) {
#pragma unroll
                for (int t = 0; t < NP; ++t)
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const double s = warp_sum((double)accP[t][b]);
                        if (lane == 0) p.tile_head[((int64_t)tl.tile * NP + t) * BT + b] = s;
                        if constexpr (MID) {
                            const double s2 = warp_sum((double)accP2[t][b]);
                            if (lane == 0) p.tile_tail[((int64_t)tl.tile * NP + t) * BT + b] = s2;
                        }
                    }
            } else {
                double *head = p.tile_head + (int64_t)tl.tile * NP * BT;
                double *tail = p.tile_tail + (int64_t)tl.tile * NP * BT;
                if (!staged) {
                    // more groups than the scratch holds: accumulate in global memory; the groups of the
                    // tile that no single lane stored are cleared first (this set never changes)
                    if (lane < NP * BT) {
                        head[lane] = 0.0;
                        tail[lane] = 0.0;
                    }
                    __syncwarp();
                }
                // round A: the first run of every lane (or its only run); round B: the last run
                double v[NPa * BT];
                int key;
                bool has;
                for (int round = 0; round < 2; ++round) {
                    if (round == 0) {
                        // first run of the lane unless it was a complete group stored directly
                        has = nvalid > 0 && (p_have_head || !p_changed);
                        key = p_have_head ? p_head_key : kcur;
#pragma unroll
                        for (int k = 0; k < NP * BT; ++k)
                            v[k] = p_have_head ? p_head[k] : (double)accP[k / BT][k % BT];
                    } else {
                        has = p_changed && nvalid > 0;
                        key = kcur;
#pragma unroll
                        for (int k = 0; k < NP * BT; ++k) v[k] = (double)accP[k / BT][k % BT];
                    }
                    // first-run keys ascend with the lane and equal keys sit in consecutive lanes; a last
                    // run that began inside its lane is the only pending piece of its group in this round
                    const bool leader = round == 0 ? warp_sorted_run_sum<double, NPa * BT>(has, key, v) : has;
                    if (leader) {
                        if (staged) {
#pragma unroll
                            for (int k = 0; k < NP * BT; ++k) s_acc[(key - tl.kf) * NP * BT + k] += v[k];
                            s_touch[key - tl.kf] = 1;
                        } else {
                            double *sink = key == tl.kf ? head : key == tl.kl ? tail : nullptr;
#pragma unroll
                            for (int k = 0; k < NP * BT; ++k) {
                                double *dst =
                                    sink ? sink + k
                                         : p.direct_p + ((int64_t)(k / BT) * p.n_primary + key) * BT + (k % BT);
                                *dst = *reinterpret_cast<volatile double *>(dst) + v[k];
                            }
                        }
                    }
                    __syncwarp();
                }
                if (staged && lane < ng) {
                    // one plain store per group: first -> tile_head, last -> tile_tail, groups in between
                    // (not stored by a single lane already) -> direct_p
                    const int gkey = tl.kf + lane;
                    if (lane == 0 || lane == ng - 1 || s_touch[lane]) {
#pragma unroll
                        for (int k = 0; k < NP * BT; ++k) {
                            const double tot = s_acc[lane * NP * BT + k];
                            if (lane == 0)
                                head[k] = tot;
                            else if (lane == ng - 1)
                                tail[k] = tot;
                            else
                                p.direct_p[((int64_t)(k / BT) * p.n_primary + gkey) * BT + (k % BT)] = tot;
                        }
                    }
                }
                __syncwarp();
            }
        }

        // ---- secondary level: the runs still open at the end of the lanes' spans ------------------------
        if constexpr (NS > 0) {
            double v[NSa * BT];
            if constexpr (FASTP) {
                // codes ascend with the lane: one sorted-run sum (in the table's own type), the first
                // lane of each run updates the table.  (Two groups in the tile: the lanes that end in the
                // first group, then those that end in the second -- the codes ascend within either set.)
                __syncwarp();
#pragma unroll
                for (int side = 0; side < (MID ? 2 : 1); ++side) {
                    const bool mine = !MID || ((li + R - 1 >= split) == (side == 1));
                    T w[NSa * BT];
#pragma unroll
                    for (int k = 0; k < NS * BT; ++k) w[k] = accS[k / BT][k % BT];
                    if (warp_sorted_run_sum<T, NSa * BT>(nvalid > 0 && mine, scur, w)) {
#pragma unroll
                        for (int k = 0; k < NS * BT; ++k)
                            gradS[((int64_t)(k / BT) * p.n_secondary + scur) * BT + (k % BT)] += w[k];
                    }
                    __syncwarp();
                }
            } else {
                // a tile that straddles segments may repeat a code anywhere: general grouping, first
                // runs (or only runs) in round A, last runs in round B
                for (int round = 0; round < 2; ++round) {
                    bool has;
                    int key;
                    if (round == 0) {
                        has = nvalid > 0;
                        key = s_changed ? s_head_key : scur;
#pragma unroll
                        for (int k = 0; k < NS * BT; ++k) v[k] = s_changed ? s_head[k] : (double)accS[k / BT][k % BT];
                    } else {
                        has = s_changed;
                        key = scur;
#pragma unroll
                        for (int k = 0; k < NS * BT; ++k) v[k] = (double)accS[k / BT][k % BT];
                    }
                    if (warp_match_sum<NSa * BT>(has, key, v)) {
#pragma unroll
                        for (int k = 0; k < NS * BT; ++k)
                            gradS[((int64_t)(k / BT) * p.n_secondary + key) * BT + (k % BT)] += (T)v[k];
                    }
                    __syncwarp();
                }
            }
        }
    }

    // -----------------------------------------------------------------------------------------
    static __device__ __forceinline__ void issue_tile(const RowParams<T> &p, int tile, unsigned char *stage,
                                                      const StageLayout &sl, uint64_t *bar) {
        const int64_t row = (int64_t)tile * WT;
        const bool y_u8 = Spec::y_u8(p), has_off = Spec::has_off(p);
        const uint32_t xbytes = WT * (uint32_t)sizeof(T);
        const uint32_t ybytes = y_u8 ? WT : xbytes;
        const uint32_t cbytes = NS > 0 ? WT * (Spec::code16(p) ? 2u : 4u) : 0u;
        const int n_x = Spec::n_x(p);
        mbar_arrive_expect_tx(bar, (uint32_t)n_x * xbytes + ybytes + (has_off ? xbytes : 0u) + cbytes);
#pragma unroll
        for (int c = 0; c < kMaxXCols; ++c)
            if (c < n_x) tma_load_1d(stage + c * sl.x_stride, p.x[c] + row, xbytes, bar);
        if (y_u8)
            tma_load_1d(stage + sl.y_off, p.y8 + row, ybytes, bar);
        else
            tma_load_1d(stage + sl.y_off, p.y + row, ybytes, bar);
        if (has_off) tma_load_1d(stage + sl.off_off, p.eta_offset + row, xbytes, bar);
        if (NS > 0)
            tma_load_1d(stage + sl.code_off,
                        reinterpret_cast<const unsigned char *>(p.sec_codes) + row * (Spec::code16(p) ? 2 : 4), cbytes,
                        bar);
    }

    // values of the primary terms at a tile's first slot (their theta elements ride in the tile meta)
    static __device__ __forceinline__ void load_vp(const RowParams<T> &p, const int4 &meta, T (&vp)[NPa][BT]) {
#pragma unroll
        for (int t = 0; t < NPa; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) vp[t][b] = (NP > 0) ? theta_at(p, b, t == 0 ? meta.z : meta.w) : T(0);
    }

    // -----------------------------------------------------------------------------------------
    // Prior terms, evaluated next to the first tiles.  theta is cut into groups of 32 elements and the
    // links into their chunks of 32; group / chunk q belongs to CTA q % grid and there to the warps
    // from the top (warp W-1 first): the last warps of a CTA are the ones that hold one tile less when
    // the tiles do not divide evenly, and no warp of the grid gets more than ceil(groups / grid) of them.
    // Own gradient term -> prior_grad, logp -> one partial per warp; then, in link order, what the
    // children contribute to the gradient of their priors' parameters -> one partial per chunk of 32
    // links (summed per hyper-parameter by the finish kernel).  Same arithmetic as skk_prior_kernel.
    static __device__ __forceinline__ void priors(const RowParams<T> &p, int warp, int lane) {
        const int warps = blockDim.x >> 5;
        double lp[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) lp[b] = 0.0;
        for (int64_t q = (int64_t)(warps - 1 - warp) * gridDim.x + blockIdx.x; q * 32 < p.n_params;
             q += (int64_t)warps * gridDim.x) {
            const int64_t j = q * 32 + lane;
            if (j >= p.n_params) continue;
            const int family = p.prior_family[j];
            const int a_src = p.prior_a_src[j], b_src = p.prior_b_src[j];
            const double a_c = p.prior_a_const[j], b_c = p.prior_b_const[j];
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                if (b >= p.b_valid) continue;
                const T *th = p.theta + (int64_t)b * p.n_params;
                const double a = a_src >= 0 ? prior_mu_of(family, (double)th[a_src]) : a_c;
                const double sd = b_src >= 0 ? exp((double)th[b_src]) : b_c;
                double grad, d_mu, d_log_sd;
                lp[b] += prior_term<false>(family & kPriorFamilyMask, (double)th[j], a, sd, grad, d_mu, d_log_sd);
                p.prior_grad[(int64_t)b * p.n_params + j] = grad;
            }
        }
        const int64_t parts = (int64_t)gridDim.x * warps;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const double s = warp_sum(lp[b]);
            if (lane == 0 && b < p.b_valid) p.prior_lp_part[b * parts + blockIdx.x * warps + warp] = s;
        }
        // links: a warp holds one whole chunk
        for (int64_t c = (int64_t)(warps - 1 - warp) * gridDim.x + blockIdx.x; c < p.n_chunks;
             c += (int64_t)warps * gridDim.x) {
            const int64_t w = c * 32 + lane;
            const int j = p.link_child[w];
            const int kind = p.link_kind[w];
            const int a_src = p.link_a_src[w], b_src = p.link_b_src[w];
            const double a_c = p.link_a_const[w], b_c = p.link_b_const[w];
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                double v = 0.0;
                if (j >= 0 && b < p.b_valid) {
                    const T *th = p.theta + (int64_t)b * p.n_params;
                    // kind bits 0-1: 0 link to the prior's mu, 1 to its sigma; bit 2: the child's mu is a
                    // positive (log-transformed) variable, mu = exp(u), d/du = d/dmu * mu
                    const bool mu_is_log = (kind & 4) != 0;
                    const double a = a_src >= 0 ? (mu_is_log ? exp((double)th[a_src]) : (double)th[a_src]) : a_c;
                    const double sd = b_src >= 0 ? exp((double)th[b_src]) : b_c;
                    double grad, d_mu, d_log_sd;
                    prior_term<false>(SKK_PRIOR_NORMAL, (double)th[j], a, sd, grad, d_mu, d_log_sd);
                    v = (kind & 3) == 0 ? (mu_is_log ? d_mu * a : d_mu) : d_log_sd;
                }
                v = warp_sum(v);
                if (lane == 0 && b < p.b_valid) p.link_part[(int64_t)b * p.n_chunks + c] = v;
            }
        }
    }

    // CTA-wide: the warps' accumulator tables -> the CTA's fp64 partial, tables cleared (all warps call it at the
    // same point of their tile lists)
    static __device__ __forceinline__ void fold_tables(const RowParams<T> &p, T *gradS_all, int warps, bool first) {
        const int n = NS * p.n_secondary * BT;
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            double s = 0.0;
            for (int w = 0; w < warps; ++w) {
                s += (double)gradS_all[(int64_t)w * n + k];
                gradS_all[(int64_t)w * n + k] = T(0);
            }
            double *dst = p.cta_sec + (int64_t)blockIdx.x * n + k;
            *dst = first ? s : *dst + s;
        }
        __syncthreads();
    }

    static __device__ void run(const RowParams<T> &p) {
        extern __shared__ __align__(128) unsigned char smem[];
        const int warps = blockDim.x >> 5;
        // broadcast from lane 0: provably warp-uniform, so everything derived from it (the warp's
        // barriers, staging ring, scratch and accumulator table) can live in uniform registers
        const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
        const int lane = threadIdx.x & 31;
        const int stages = p.stages;

        const StageLayout sl =
            make_stage_layout<T>(WT, p.n_x, p.y8 != nullptr, p.eta_offset != nullptr, NS > 0 ? p.sec_code_bytes : 0);
        const CtaLayout cl = make_cta_layout<T>(warps, stages, sl.bytes, NS, p.n_secondary, BT, NP);
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + cl.bars_off) + warp * stages;
        T *valS = reinterpret_cast<T *>(smem + cl.vals_off);
        T *gradS_all = reinterpret_cast<T *>(smem + cl.grads_off);
        T *gradS = gradS_all + (int64_t)warp * NS * p.n_secondary * BT;
        // the table base as a 32-bit shared address pinned in a register: without the opaque move the
        // compiler re-derives warp * n_secondary in every row of the unrolled loop
        uint32_t gradS_addr = smem_addr(gradS);
        asm volatile("mov.u32 %0, %0;" : "+r"(gradS_addr));
        unsigned char *stage0 = smem + cl.stage_off + (int64_t)warp * stages * sl.bytes;
        unsigned char *scratch = smem + cl.scratch_off + warp * cl.scratch_bytes;

        // The tiles of this warp, in order, from the plan's static schedule (cost-balanced over all warps of the
        // grid at plan creation: tiles with two or more groups cost more than the common ones, and so does the
        // prior work some warps carry).  A fixed list per warp keeps every sum in a fixed order.
        // [sched_len + 4] entries per warp, -1 behind the last tile; sched_len is a multiple of 4.
        const int32_t *my_tiles = p.sched + ((int64_t)blockIdx.x * warps + warp) * (p.sched_len + 4);
        // optional timeline of the warp (skk_trace_enable); nothing of it is carried through the tile loop
        auto stamp = [&](int word) {
            if (p.trace != nullptr && lane == 0)
                p.trace[((int64_t)blockIdx.x * warps + warp) * kTraceWords + word] = global_timer_ns();
        };
        stamp(0);
        const int4 first4 = *reinterpret_cast<const int4 *>(my_tiles);

        // The small loads of the prologue are requested BEFORE the pipeline starts: all warps of the grid start
        // their rings at once (c4: 296 x 8 x 3 tiles = 27 MB), and a load issued behind that burst waits for it in the
        // memory system -- the value table used to be complete 4.8 us after the kernel's entry, most of it queueing.
        // Everything the warp needs before its first tile is two levels of dependent global loads (slot ->
        // theta element -> value): what needs no earlier result goes first.
        // level 1: theta elements of the global terms, elements of the value table (none to load when the planner
        // numbered the slots so that slot -> theta is `base + slot`), and the table's values themselves
        int eg[NGa];
#pragma unroll
        for (int t = 0; t < NGa; ++t) eg[t] = (NG > 0 && t < NG) ? p.ti_g[t][0] : 0;
        constexpr int kFillUnroll = 4;
        const int n_sec = NS * p.n_secondary * BT;
        int element[kFillUnroll];
        T vfill[kFillUnroll];
        if constexpr (NS > 0) {
#pragma unroll
            for (int u = 0; u < kFillUnroll; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                element[u] = k >= n_sec ? -1
                             : (NS == 1 && p.sec_theta_base >= 0) ? p.sec_theta_base + (k / BT) % p.n_secondary
                                                                  : p.ti_s[k / (BT * p.n_secondary)][(k / BT) % p.n_secondary];
            }
#pragma unroll
            for (int u = 0; u < kFillUnroll; ++u)
                vfill[u] = element[u] >= 0 ? theta_at(p, (threadIdx.x + u * blockDim.x) % BT, element[u]) : T(0);
        }

        int t0 = first4.x, t1 = first4.y, t2 = first4.z, t3 = first4.w;
        // keys of the first two tiles (their theta elements ride in the tile meta)
        int4 k0 = make_int4(0, 0, 0, 0), k1 = make_int4(0, 0, 0, 0);
        int sp0 = -1, sp1 = -1;
        if constexpr (NP > 0) {
            if (t0 >= 0) k0 = p.tile_meta[t0];
            if (t1 >= 0) k1 = p.tile_meta[t1];
            if constexpr (NS == 0) {
                if (t0 >= 0) sp0 = p.tile_split[t0];
                if (t1 >= 0) sp1 = p.tile_split[t1];
            }
        }

        if (lane == 0) {
            for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
            mbar_fence_init();
        }
        __syncwarp();
        // start the pipeline
        if (elect_one()) {
            for (int s = 0; s < stages; ++s) {   // stages <= 3
                const int t = s == 0 ? t0 : (s == 1 ? t1 : t2);
                if (t >= 0) issue_tile(p, t, stage0 + (int64_t)s * sl.bytes, sl, bars + s);
            }
        }
        if (p.peer_seq != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *p.peer_seq = *p.peer_seq + 1;
        // the kernel that follows in the stream may start its own prologue under this kernel's tail (PDL)
        if (p.pdl_trigger == 1) grid_dep_launch_dependents();

        // lane j of a straddling tile stages group kf + j: its theta elements are fetched one tile ahead
        auto staged_elements = [&](const int4 &meta, int (&out)[NPa]) {
#pragma unroll
            for (int t = 0; t < NPa; ++t)
                out[t] = (NP > 0 && meta.x != meta.y && lane <= meta.y - meta.x && meta.y - meta.x < kStagedGroups)
                             ? p.ti_p[t][meta.x + lane] : 0;
        };
        // level 2.  A warp issues in order: the accumulator tables are cleared under the latency of the value-table
        // loads requested above, then the values are stored; the values of the global terms and of the first tile's
        // group -- whose theta elements are level-1 results, back by then -- are requested last and arrive under the
        // barrier and the first wait for a tile.
        if constexpr (NS > 0) {
            for (int k = threadIdx.x; k < n_sec * warps; k += blockDim.x) gradS_all[k] = T(0);
            stamp(13);
#pragma unroll
            for (int u = 0; u < kFillUnroll; ++u) {
                const int k = threadIdx.x + u * blockDim.x;
                if (element[u] >= 0) valS[k] = vfill[u];
            }
            // (tables with more entries than one round of the CTA: the rest, four at a time as above)
            for (int k0f = threadIdx.x + kFillUnroll * blockDim.x; k0f < n_sec; k0f += kFillUnroll * blockDim.x) {
                int el[kFillUnroll];
#pragma unroll
                for (int u = 0; u < kFillUnroll; ++u) {
                    const int k = k0f + u * blockDim.x;
                    el[u] = k >= n_sec ? -1
                            : (NS == 1 && p.sec_theta_base >= 0) ? p.sec_theta_base + (k / BT) % p.n_secondary
                                                                 : p.ti_s[k / (BT * p.n_secondary)][(k / BT) % p.n_secondary];
                }
#pragma unroll
                for (int u = 0; u < kFillUnroll; ++u) {
                    const int k = k0f + u * blockDim.x;
                    if (el[u] >= 0) valS[k] = theta_at(p, k % BT, el[u]);
                }
            }
            stamp(12);  // (value table written)
        }
        T vg[NGa][BT];
#pragma unroll
        for (int t = 0; t < NGa; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) vg[t][b] = (NG > 0) ? theta_at(p, b, eg[t]) : T(0);
        T vp0[NPa][BT];
        load_vp(p, k0, vp0);
        int se0[NPa], se1[NPa];
        staged_elements(k0, se0);
        if constexpr (NS > 0) __syncthreads();
        stamp(8);

        // (after the barrier: the warps without prior work start on their tiles at once)
        if (p.prior_family != nullptr) priors(p, warp, lane);
        stamp(9);

        double accG64[NG + 1][BT];
#pragma unroll
        for (int t = 0; t <= NG; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) accG64[t][b] = 0.0;

        // tiles with exactly two groups (and no secondary level) take the two-group path: the split row
        // and the second group's values are fetched ahead like the first group's
        T vq0[NPa][BT], vq1[NPa][BT];
        auto load_second = [&](const int4 &meta, const int (&se)[NPa], T (&vq)[NPa][BT]) {
#pragma unroll
            for (int t = 0; t < NPa; ++t) {
                const int element = __shfl_sync(kFull, se[t], 1);  // lane 1 holds the element of group kf + 1
#pragma unroll
                for (int b = 0; b < BT; ++b)
                    vq[t][b] = (NP > 0 && NS == 0 && meta.y == meta.x + 1) ? theta_at(p, b, element) : T(0);
            }
        };
        load_second(k0, se0, vq0);

        stamp(1);
        if (p.trace != nullptr && t0 >= 0) {
            mbar_wait(bars, 0);  // (the loop's own wait on this phase then returns at once)
            stamp(2);
        }
        int s = 0;             // stage of the ring that holds the current tile, and the parity of its barrier
        uint32_t parity = 0;   // (counted, not derived from the iteration number: no integer division per tile)
        // the accumulator tables of the secondary level are fp32 for fp32 plans: every kFoldTiles tiles the CTA
        // adds them into its fp64 partial (cta_sec) and clears them, so an entry never collects more than
        // 2 * kFoldTiles + 1 rounded additions however long the shard is (n_folds is the same for all warps)
        int pos = 0;
        for (; t0 >= 0; ++pos) {
            const int tile = t0;
            int4 k2 = make_int4(0, 0, 0, 0);
            T vp1[NPa][BT];
            if constexpr (NP > 0) {
                if (t2 >= 0) k2 = p.tile_meta[t2];
            }
            load_vp(p, k1, vp1);
            staged_elements(k1, se1);
            load_second(k1, se1, vq1);
            int sp2 = -1;
            if constexpr (NP > 0 && NS == 0) {
                if (t2 >= 0) sp2 = p.tile_split[t2];
            }
            const int t4 = my_tiles[pos + 4];
            mbar_wait(bars + s, parity);

            Tile tl{stage0 + (int64_t)s * sl.bytes, sl, tile, k0.x, k0.y};
            const bool full = (int64_t)(tile + 1) * WT <= p.n_rows;
            if (NP == 0 || k0.x == k0.y) {
                if (full)
                    process<true, true>(p, tl, vp0, vq0, -1, se0, vg, accG64, valS, gradS, gradS_addr, scratch, lane);
                else
                    process<true, false>(p, tl, vp0, vq0, -1, se0, vg, accG64, valS, gradS, gradS_addr, scratch, lane);
            } else if (NS == 0 && full && sp0 >= 0) {
                process<true, true, true>(p, tl, vp0, vq0, sp0, se0, vg, accG64, valS, gradS, gradS_addr, scratch, lane);
            } else if (NS > 0 && full && k0.y == k0.x + 1) {
                // two groups and a secondary level: the split row and the second group's values are
                // fetched here (two dependent loads, a few times per warp) rather than carried through
                // the loop next to the common path, which runs at the register limit
                const int split = p.tile_split[tile];
                T vsecond[NPa][BT];
#pragma unroll
                for (int t = 0; t < NPa; ++t)
#pragma unroll
                    for (int b = 0; b < BT; ++b)
                        vsecond[t][b] = (NP > 0 && t < NP) ? theta_at(p, b, p.ti_p[t][k0.y]) : T(0);
                process<true, true, true>(p, tl, vp0, vsecond, split, se0, vg, accG64, valS, gradS, gradS_addr, scratch,
                                          lane);
            } else {
                process<false, false>(p, tl, vp0, vq0, -1, se0, vg, accG64, valS, gradS, gradS_addr, scratch, lane);
            }

            __syncwarp();
            const int next = stages == 2 ? t2 : t3;   // the tile `stages` positions ahead takes over this slot
            if (next >= 0) {
                if (elect_one()) {
                    fence_proxy_async();
                    issue_tile(p, next, stage0 + (int64_t)s * sl.bytes, sl, bars + s);
                }
            }
            if (++s == stages) {
                s = 0;
                parity ^= 1u;
            }
            t0 = t1;
            t1 = t2;
            t2 = t3;
            t3 = t4;
            if constexpr (NS > 0) {
                if (((pos + 1) & (kFoldTiles - 1)) == 0 && (pos + 1) / kFoldTiles <= p.sched_folds)
                    fold_tables(p, gradS_all, warps, pos + 1 == kFoldTiles);
            }
            k0 = k1;
            k1 = k2;
            sp0 = sp1;
            sp1 = sp2;
#pragma unroll
            for (int t = 0; t < NPa; ++t) {
                se0[t] = se1[t];
#pragma unroll
                for (int b = 0; b < BT; ++b) vq0[t][b] = vq1[t][b];
            }
#pragma unroll
            for (int t = 0; t < NPa; ++t)
#pragma unroll
                for (int b = 0; b < BT; ++b) vp0[t][b] = vp1[t][b];
        }

        stamp(3);
        if (p.pdl_trigger == 2) grid_dep_launch_dependents();
        if (p.trace != nullptr && lane == 0) {
            // tiles of this warp by class (one group / two groups / more), for tools/trace_rows.py
            unsigned long long n_class[3] = {0, 0, 0};
            for (int i = 0; my_tiles[i] >= 0; ++i) {
                const int4 meta = p.tile_meta[my_tiles[i]];
                const int groups = NP > 0 ? meta.y - meta.x : 0;
                ++n_class[groups < 2 ? groups : 2];
            }
            for (int c = 0; c < 3; ++c) p.trace[((int64_t)blockIdx.x * warps + warp) * kTraceWords + 5 + c] = n_class[c];
        }
        // ---- CTA epilogue: fixed-order reductions over lanes, then warps ---------------------------
        __syncthreads();  // all warps are done with their staging buffers; reuse the first one
        stamp(10);
        double *red = reinterpret_cast<double *>(smem + cl.stage_off);  // [warps][(NG+1)*BT]
        constexpr int NV = (NG + 1) * BT;
#pragma unroll
        for (int t = 0; t <= NG; ++t)
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const double s = warp_sum(accG64[t][b]);
                if (lane == 0) red[warp * NV + t * BT + b] = s;
            }
        __syncthreads();
        if (threadIdx.x < NV) {
            double s = 0.0;
            for (int w = 0; w < warps; ++w) s += red[w * NV + threadIdx.x];
            p.cta_glob[(int64_t)blockIdx.x * NV + threadIdx.x] = s;
        }
        stamp(11);
        if constexpr (NS > 0) {
            const int n = NS * p.n_secondary * BT;
            for (int k = threadIdx.x; k < n; k += blockDim.x) {
                double s = 0.0;
                for (int w = 0; w < warps; ++w) s += (double)gradS_all[(int64_t)w * n + k];
                double *dst = p.cta_sec + (int64_t)blockIdx.x * n + k;
                *dst = p.sched_folds == 0 ? s : *dst + s;
            }
        }
        stamp(4);
    }
};

template <typename T, int FAM, int BT, int NG, int NP, int NS, int R, int SPEC = -1>
__global__ void __launch_bounds__(256, (BT == 1 ? 2 : 1)) skk_row_kernel(const __grid_constant__ RowParams<T> p) {
    RowKernel<T, FAM, BT, NG, NP, NS, R, SPEC>::run(p);
}

}  // namespace skk
