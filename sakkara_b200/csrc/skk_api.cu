// Host side of libskk_b200.so: the C ABI of include/skk_b200.h.
//   plan creation : validate the descriptor, upload the row columns (padded to whole tiles), derive
//                   the tile keys, the reduce CSR and the prior link CSR, pick the kernel
//                   specialisation and its launch geometry, allocate every work buffer once.
//   evaluation    : single GPU, every theta element fed by at most one slot:
//                     H2D theta -> per batch tile {K1 rows (priors in its prologue), finish} -> D2H;
//                   otherwise: H2D theta -> {K3a, K3b priors on a parallel branch} || per batch tile
//                     {K1 rows, K4a slot totals, K4f theta} -> [peer exchange | NCCL all-reduce -> K4b final
//                     when row-sharded] -> D2H;  batches >= 32 without a secondary level: K2 chain kernels.
//                   No allocation, no host synchronisation before the final copy; the sequence is
//                   captured once per (batch, buffers) and replayed as a CUDA graph.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <queue>
#include <utility>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "skk_aux_kernels.cuh"
#include "skk_chain_kernel.cuh"
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f32_normal_bt1(int, int, int, int);
const void *row_kernel_f32_normal_bt4(int, int, int, int);
const void *row_kernel_f32_bernoulli_bt1(int, int, int, int);
const void *row_kernel_f32_bernoulli_bt4(int, int, int, int);
const void *row_kernel_f32_poisson_bt1(int, int, int, int);
const void *row_kernel_f32_poisson_bt4(int, int, int, int);
const void *row_kernel_f64_normal_bt1(int, int, int, int);
const void *row_kernel_f64_normal_bt4(int, int, int, int);
const void *row_kernel_f64_bernoulli_bt1(int, int, int, int);
const void *row_kernel_f64_bernoulli_bt4(int, int, int, int);
const void *row_kernel_f64_poisson_bt1(int, int, int, int);
const void *row_kernel_f64_poisson_bt4(int, int, int, int);
const void *row_kernel_f32_negbinomial_bt1(int, int, int, int);
const void *row_kernel_f32_negbinomial_bt4(int, int, int, int);
const void *row_kernel_f64_negbinomial_bt1(int, int, int, int);
const void *row_kernel_f64_negbinomial_bt4(int, int, int, int);
const void *row_kernel_f32_gamma_bt1(int, int, int, int);
const void *row_kernel_f32_gamma_bt4(int, int, int, int);
const void *row_kernel_f64_gamma_bt1(int, int, int, int);
const void *row_kernel_f64_gamma_bt4(int, int, int, int);
const void *row_kernel_f32_logsigma_bt1(int, int, int, int);
const void *row_kernel_f32_logsigma_bt4(int, int, int, int);
const void *row_kernel_f64_logsigma_bt1(int, int, int, int);
const void *row_kernel_f64_logsigma_bt4(int, int, int, int);
const void *row_kernel_f32_studentt_bt1(int, int, int, int);
const void *row_kernel_f32_studentt_bt4(int, int, int, int);
const void *row_kernel_f64_studentt_bt1(int, int, int, int);
const void *row_kernel_f64_studentt_bt4(int, int, int, int);
const void *chain_kernel_f32(int, int, int);
const void *chain_kernel_f64(int, int, int);
}  // namespace skk

using namespace skk;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_error;

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t err__ = (expr);                                                                    \
        if (err__ != cudaSuccess)                                                                      \
            return fail(SKK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                        __LINE__);                                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: no link-time dependency, the library loads on machines without NCCL and
// shares the libnccl.so.2 that torch already mapped into the process.
// ---------------------------------------------------------------------------------------------
struct NcclId {
    char bytes[128];
};
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.handle) return SKK_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(SKK_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(SKK_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.handle = h;
    return SKK_OK;
}

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        int r__ = (expr);                                                                           \
        if (r__ != 0)                                                                               \
            return fail(SKK_ERR_NCCL, "%s failed: %s", #expr,                                       \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "unknown NCCL error"); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct LaunchCfg {
    const void *fn = nullptr;
    int grid = 0;
    int smem = 0;
    int32_t *sched = nullptr;  // static tile schedule of this grid (device): [grid * warps][sched_len + 4]
    int sched_len = 0, sched_folds = 0;
};

struct GraphEntry {
    int batch = 0;
    unsigned flags = 0;
    const void *theta = nullptr, *logp = nullptr, *grad = nullptr;
    cudaGraphExec_t exec = nullptr;
};

struct skk_plan {
    int dtype = 0, family = 0, device = 0;
    int64_t n_rows = 0, n_params = 0, n_rows_total = 0;
    int max_batch = 1;
    int n_x = 0, sec_code_bytes = 4;
    bool y_u8 = false, has_off = false;
    int64_t n_primary = 0, n_secondary = 0;
    int ng = 0, np = 0, ns = 0;
    int xcol_g[kMaxTermsPerLevel], xcol_p[kMaxTermsPerLevel], xcol_s[kMaxTermsPerLevel];
    double sigma_const = 1.0, logp_const = 0.0;
    int64_t sigma_idx = -1;
    int include_priors = 1;

    int rows_per_tile = 0, n_tiles = 0, warps = 8, stages = 3, n_sms = 0;
    LaunchCfg cfg1, cfg4;  // batch tile 1 and 4

    cudaStream_t stream = nullptr, side = nullptr;  // side: the prior kernels run next to the row kernel
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    std::vector<void *> allocs;
    unsigned char *arena = nullptr;   // current arena of the small allocations
    size_t arena_used = 0, arena_size = 0;
    void *first_arena = nullptr;      // (the one with the persisting L2 window)
    bool l2_window = false;
    size_t device_bytes = 0;
    bool use_graphs = true;
    std::vector<GraphEntry> graphs;

    void *x[kMaxXCols] = {nullptr};
    void *y = nullptr, *off = nullptr, *codes = nullptr;
    // the same columns in row order for the chain kernel (the row kernel's copies are stored tile-transposed,
    // skk_row_kernel.cuh: kStageTransposed); null when the plan has no chain path
    void *x_rows[kMaxXCols] = {nullptr};
    void *y_rows = nullptr, *off_rows = nullptr;
    int4 *tile_meta = nullptr;
    int32_t *tile_split = nullptr;
    int64_t *seg_off = nullptr;
    int n_terms = 0;

    void *theta_dev = nullptr;
    double *totals = nullptr;
    double *direct_p = nullptr, *direct_p1 = nullptr;  // per batch-tile width (4 / 1): the entries the row kernel
                                                       // writes are fixed per width, the rest must stay zero
    double *tile_head = nullptr, *tile_tail = nullptr, *cta_glob = nullptr, *cta_sec = nullptr;
    double *lik_vec = nullptr, *prior_grad = nullptr, *prior_lp = nullptr, *link_val = nullptr;
    // links in chunk order (fused single-GPU path)
    int32_t *lk_child = nullptr, *lk_a_src = nullptr, *lk_b_src = nullptr;
    int8_t *lk_kind = nullptr;
    double *lk_a_const = nullptr, *lk_b_const = nullptr, *link_part = nullptr;
    int64_t n_chunks = 0;
    void *out_dev = nullptr;  // [B*P grad | B logp]

    PriorParams prior{};
    ReduceParams reduce{};
    int prior_ctas = 0;
    bool fused_ok = false;  // single-GPU evaluation = row kernel (with priors) + finish kernel

    void *h_in = nullptr, *h_out = nullptr;  // pinned staging

    void *comm = nullptr;
    int nranks = 1, rank = 0;
    // chain-per-lane path (K2) for large batches
    bool chain_ready = false;
    const void *chain_fn = nullptr;
    int chain_ranges = 0, chain_tiles = 0, chain_smem = 0, chain_bp = 0;
    int2 *chain_keys = nullptr;
    double *chain_totals = nullptr, *chain_head = nullptr, *chain_tail = nullptr, *chain_glob = nullptr;
    ChainReduceParams chain_reduce{};
    int32_t *theta_index_dev[kMaxTerms] = {nullptr};
    int32_t sec_theta_base = -1;  // >= 0: the secondary term reads theta[base + slot]
    int term_level[kMaxTerms] = {0}, term_slot_in_level[kMaxTerms] = {0};
    // one-shot all-reduce over peer memory
    PeerParams peer{};
    bool peer_ready = false;
    void *peer_local = nullptr;            // this rank's receive buffer
    void *peer_opened[kMaxPeers] = {nullptr};
    int *peer_status = nullptr;            // mapped host memory
    int peer_nranks = 0;
    std::vector<uint8_t> fed;              // [P + 2] entries of the exchanged vector this rank feeds
    bool peer_failed = false;              // an exchange timed out: the ranks' counters may disagree from then on
    int exchange_ctas_max = 0;             // CTAs of skk_finish_exchange_kernel that are resident at once

    unsigned long long *trace = nullptr;   // per-warp timeline of the row kernel (skk_trace_enable)
    size_t trace_words = 0;

    skk_stats_t stats{};
};

static const int kChainMinBatch = 32;     // batches from this size on use the chain-per-lane kernels
static const int64_t kHeavyLinks = 512;  // hyper-parameters with more children get a whole CTA

static size_t elt_size(int dtype) { return dtype == SKK_F64 ? 8 : 4; }

// Device memory of a plan.  Row columns and other large buffers are allocations of their own; everything small
// (descriptors, slot tables, partial sums, theta and the outputs) is carved out of shared 4 MiB arenas, so that
// what an evaluation touches besides the rows lies in a handful of pages -- after the caches and TLBs have
// been swept (the bench flushes L2 between steps) every separate allocation costs its own page walk.
// The first arena is larger and, unless SKK_L2_PERSIST=0, is given a persisting L2 access-policy window on the plan's
// streams: the row data of one evaluation sweeps the whole L2 (c4: 700 MB per pass), and without the window every
// evaluation would start by fetching its descriptors, schedules, theta and the previous partials' lines from HBM again
// -- each a full memory latency on the dependent-load chains that the small kernels consist of.
static const size_t kArenaBytes = 4u << 20, kArenaMaxItem = 4u << 20, kFirstArenaBytes = 24u << 20;

template <typename U>
static int dev_alloc(skk_plan *pl, U **out, size_t count, bool zero = false) {
    void *p = nullptr;
    const size_t bytes = (std::max<size_t>(count * sizeof(U), 16) + 255) & ~(size_t)255;
    if (bytes <= kArenaMaxItem) {
        if (!pl->arena || pl->arena_used + bytes > pl->arena_size) {
            const size_t size = pl->arena ? kArenaBytes : kFirstArenaBytes;
            CUDA_TRY(cudaMalloc(&p, size));
            pl->allocs.push_back(p);
            pl->device_bytes += size;
            if (!pl->arena) pl->first_arena = p;
            pl->arena = static_cast<unsigned char *>(p);
            pl->arena_size = size;
            pl->arena_used = 0;
        }
        p = pl->arena + pl->arena_used;
        pl->arena_used += bytes;
    } else {
        CUDA_TRY(cudaMalloc(&p, bytes));
        pl->allocs.push_back(p);
        pl->device_bytes += bytes;
    }
    if (zero) CUDA_TRY(cudaMemset(p, 0, bytes));
    *out = static_cast<U *>(p);
    return SKK_OK;
}

template <typename U>
static int dev_upload(skk_plan *pl, U **out, const U *host, size_t count, size_t padded_count = 0) {
    const size_t n = std::max(count, padded_count);
    int rc = dev_alloc(pl, out, n, padded_count > count);
    if (rc) return rc;
    if (count) CUDA_TRY(cudaMemcpy(*out, host, count * sizeof(U), cudaMemcpyHostToDevice));
    return SKK_OK;
}

static int upload_bytes(skk_plan *pl, void **out, const void *host, size_t bytes, size_t padded_bytes) {
    unsigned char *p = nullptr;
    int rc = dev_upload<unsigned char>(pl, &p, static_cast<const unsigned char *>(host), bytes, padded_bytes);
    *out = p;
    return rc;
}

// A column of the row kernel: tile by tile (WT = 32 * R rows, lane l owns rows l * R ... l * R + R - 1 of the tile) the
// lane's row i is stored at position i * 32 + l, so that the 32 lanes read consecutive elements at every step
// (skk_row_kernel.cuh: kStageTransposed).  Rows behind the last one are zero.
template <typename E>
static int upload_tiled(skk_plan *pl, void **out, const void *host, size_t n, size_t padded, int R) {
    E *dev = nullptr;
    if (!kStageTransposed) return upload_bytes(pl, out, host, n * sizeof(E), padded * sizeof(E));
    const E *src = static_cast<const E *>(host);
    const size_t WT = 32 * (size_t)R;
    std::vector<E> tmp(padded, E(0));
    const size_t full_tiles = n / WT;
    for (size_t t = 0; t < full_tiles; ++t) {
        const E *in = src + t * WT;
        E *o = tmp.data() + t * WT;
        for (int l = 0; l < 32; ++l)
            for (int i = 0; i < R; ++i) o[(size_t)i * 32 + l] = in[(size_t)l * R + i];
    }
    for (size_t q = full_tiles * WT; q < n; ++q) {
        const size_t w = q - full_tiles * WT, l = w / R, i = w % R;
        tmp[full_tiles * WT + i * 32 + l] = src[q];
    }
    int rc = dev_upload<E>(pl, &dev, tmp.data(), padded);
    *out = dev;
    return rc;
}

static const void *lookup_kernel(int dtype, int family, int bt, int ng, int np, int ns, int spec) {
    typedef const void *(*Lookup)(int, int, int, int);
    static const Lookup table[2][7][2] = {
        {{row_kernel_f32_normal_bt1, row_kernel_f32_normal_bt4},
         {row_kernel_f32_bernoulli_bt1, row_kernel_f32_bernoulli_bt4},
         {row_kernel_f32_poisson_bt1, row_kernel_f32_poisson_bt4},
         {row_kernel_f32_negbinomial_bt1, row_kernel_f32_negbinomial_bt4},
         {row_kernel_f32_logsigma_bt1, row_kernel_f32_logsigma_bt4},
         {row_kernel_f32_studentt_bt1, row_kernel_f32_studentt_bt4},
         {row_kernel_f32_gamma_bt1, row_kernel_f32_gamma_bt4}},
        {{row_kernel_f64_normal_bt1, row_kernel_f64_normal_bt4},
         {row_kernel_f64_bernoulli_bt1, row_kernel_f64_bernoulli_bt4},
         {row_kernel_f64_poisson_bt1, row_kernel_f64_poisson_bt4},
         {row_kernel_f64_negbinomial_bt1, row_kernel_f64_negbinomial_bt4},
         {row_kernel_f64_logsigma_bt1, row_kernel_f64_logsigma_bt4},
         {row_kernel_f64_studentt_bt1, row_kernel_f64_studentt_bt4},
         {row_kernel_f64_gamma_bt1, row_kernel_f64_gamma_bt4}}};
    if (family < 0 || family > 6 || (bt != 1 && bt != 4)) return nullptr;
    return table[dtype == SKK_F64 ? 1 : 0][family][bt == 4 ? 1 : 0](ng, np, ns, spec);
}

static int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

template <typename T>
static int cta_smem(const skk_plan *pl, int warps, int stages, int bt) {
    const StageLayout sl = make_stage_layout<T>(pl->rows_per_tile, pl->n_x, pl->y_u8, pl->has_off,
                                                pl->ns > 0 ? pl->sec_code_bytes : 0);
    return make_cta_layout<T>(warps, stages, sl.bytes, pl->ns, (int)pl->n_secondary, bt, pl->np).total;
}

template <typename T>
static int configure_launch(skk_plan *pl, int bt, LaunchCfg *cfg) {
    // which terms carry a data column, how y and the codes are stored (RowSpec bits)
    int spec = 0;
    for (int t = 0; t < pl->ng; ++t) spec |= (pl->xcol_g[t] >= 0) << t;
    for (int t = 0; t < pl->np; ++t) spec |= (pl->xcol_p[t] >= 0) << (2 + t);
    for (int t = 0; t < pl->ns; ++t) spec |= (pl->xcol_s[t] >= 0) << (4 + t);
    if (pl->y_u8) spec |= 64;
    if (pl->has_off) spec |= 128;
    if (pl->ns > 0 && pl->sec_code_bytes == 2) spec |= 256;
    // the fixed kernels count the staged x columns from the spec: one column per term that has one
    if (__builtin_popcount(spec & 63) != pl->n_x) spec = -2;
    if (env_int("SKK_GENERIC", 0)) spec = -2;  // force the run-time kernels (testing)
    cfg->fn = lookup_kernel(pl->dtype, pl->family, bt, pl->ng, pl->np, pl->ns, spec);
    if (!cfg->fn)
        return fail(SKK_ERR_UNSUPPORTED, "no row kernel for %d global / %d primary / %d secondary terms", pl->ng,
                    pl->np, pl->ns);
    cfg->smem = cta_smem<T>(pl, pl->warps, pl->stages, bt);
    CUDA_TRY(cudaFuncSetAttribute(cfg->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg->smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfg->fn, pl->warps * 32, cfg->smem));
    if (per_sm < 1) return fail(SKK_ERR_UNSUPPORTED, "row kernel does not fit on an SM (smem %d B)", cfg->smem);
    per_sm = std::min(per_sm, env_int("SKK_CTAS_PER_SM", per_sm));
    const int want = (pl->n_tiles + pl->warps - 1) / pl->warps;
    cfg->grid = std::max(1, std::min(pl->n_sms * per_sm, want));
    return SKK_OK;
}

// Static schedule of the row kernel: which tiles every warp of the grid processes, and in which order.  Tiles are
// handed out in index order to the warp with the least work so far -- measured per tile class (tools/trace_rows.py:
// a tile with two groups costs ~3x a common one on models with a secondary level (two ordered table updates
// per row; prefetching its split row and second group's slots a tile ahead changed nothing),, the rolled generic path
// more), and the warps that carry prior work in the prologue start with that much on their account.  A fixed
// list per warp keeps every sum of the evaluation in a fixed order; the balance removes the tail in which a few
// warps with several expensive tiles kept their SMs busy while all others had finished.
static int build_schedule(skk_plan *pl, const std::vector<int2> &keys, int64_t n_rows, int64_t n_prior_groups,
                          int64_t n_chunks, LaunchCfg *cfg) {
    const int W = pl->warps, grid = cfg->grid, n_warps = grid * W, WT = pl->rows_per_tile;
    // (SKK_COST_TWO / SKK_COST_PRIOR, in hundredths: calibration runs)
    const double cost_two = env_int("SKK_COST_TWO", pl->ns > 0 ? 300 : 110) / 100.0, cost_generic = pl->ns > 0 ? 4.0 : 3.0,
                 cost_prior = env_int("SKK_COST_PRIOR", 100) / 100.0;
    std::vector<double> load(n_warps, 0.0);
    if (pl->fused_ok && pl->include_priors) {
        // (skk_row_kernel.cuh: priors(): group / chunk q -> CTA q % grid, warp W - 1 - (q / grid) % W)
        for (int64_t q = 0; q < n_prior_groups; ++q) load[(q % grid) * W + (W - 1 - (q / grid) % W)] += cost_prior;
        for (int64_t q = 0; q < n_chunks; ++q) load[(q % grid) * W + (W - 1 - (q / grid) % W)] += cost_prior;
    }
    typedef std::pair<double, int> Item;
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
    for (int w = 0; w < n_warps; ++w) heap.push(Item(load[w], w));
    std::vector<std::vector<int32_t>> lists(n_warps);
    // longest processing time first: the expensive tiles are placed before the common ones (in index order within
    // a class), so the granularity left at the end is one common tile; every warp then walks its tiles in index order
    std::vector<double> cost_of(pl->n_tiles);
    std::vector<int32_t> order(pl->n_tiles);
    for (int t = 0; t < pl->n_tiles; ++t) {
        const bool full = (int64_t)(t + 1) * WT <= n_rows;
        const int groups = pl->np > 0 ? keys[t].y - keys[t].x + 1 : 1;
        cost_of[t] = groups <= 1 ? (full ? 1.0 : 1.5) : (groups == 2 && full ? cost_two : cost_generic);
        order[t] = t;
    }
    if (env_int("SKK_SCHED_INDEX_ORDER", 0) == 0)
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cost_of[a] > cost_of[b]; });
    for (int t : order) {
        Item it = heap.top();
        heap.pop();
        lists[it.second].push_back(t);
        heap.push(Item(it.first + cost_of[t], it.second));
    }
    for (auto &l : lists) std::sort(l.begin(), l.end());
    size_t longest = 0, shortest = SIZE_MAX;
    for (const auto &l : lists) {
        longest = std::max(longest, l.size());
        shortest = std::min(shortest, l.size());
    }
    cfg->sched_len = (int)((longest + 3) & ~(size_t)3);
    cfg->sched_folds = pl->ns > 0 ? (int)(shortest / kFoldTiles) : 0;
    std::vector<int32_t> flat((size_t)n_warps * (cfg->sched_len + 4), -1);
    for (int w = 0; w < n_warps; ++w) std::copy(lists[w].begin(), lists[w].end(), flat.begin() + (size_t)w * (cfg->sched_len + 4));
    return dev_upload(pl, &cfg->sched, flat.data(), flat.size());
}

template <typename T>
static int create_impl(const skk_plan_desc *d, skk_plan *pl) {
    const int WT = 32 * kRowsPerLane;
    pl->rows_per_tile = WT;
    pl->n_tiles = (int)((d->n_rows + WT - 1) / WT);
    if (pl->n_tiles < 1) pl->n_tiles = 1;
    // whole tiles of the row kernel (544 rows) and of the chain kernel (512 rows)
    const size_t padded = std::max((size_t)pl->n_tiles * WT,
                                   (size_t)((d->n_rows + kChainTileRows - 1) / kChainTileRows) * kChainTileRows);
    const size_t n = (size_t)d->n_rows;

    // ---- row columns, padded to whole tiles ------------------------------------------------------
    // (a plan with a chain path keeps a second copy in row order for the chain kernel)
    const bool row_copy = kStageTransposed && pl->ns == 0 && pl->max_batch >= kChainMinBatch && env_int("SKK_NO_CHAIN", 0) == 0 &&
                          pl->family != SKK_LIK_NORMAL_LOGSIGMA && pl->family != SKK_LIK_STUDENTT_LOGSIGMA;
    const int R = kRowsPerLane;
    for (int c = 0; c < d->n_xcols; ++c) {
        int rc = upload_tiled<T>(pl, &pl->x[c], d->xcols[c], n, padded, R);
        if (rc) return rc;
        if (row_copy && (rc = upload_bytes(pl, &pl->x_rows[c], d->xcols[c], n * sizeof(T), padded * sizeof(T)))) return rc;
    }
    {
        const size_t ye = pl->y_u8 ? 1 : sizeof(T);
        int rc = pl->y_u8 ? upload_tiled<uint8_t>(pl, &pl->y, d->y, n, padded, R) : upload_tiled<T>(pl, &pl->y, d->y, n, padded, R);
        if (rc) return rc;
        if (row_copy && (rc = upload_bytes(pl, &pl->y_rows, d->y, n * ye, padded * ye))) return rc;
    }
    if (d->eta_offset) {
        int rc = upload_tiled<T>(pl, &pl->off, d->eta_offset, n, padded, R);
        if (rc) return rc;
        if (row_copy && (rc = upload_bytes(pl, &pl->off_rows, d->eta_offset, n * sizeof(T), padded * sizeof(T)))) return rc;
    }
    if (pl->ns > 0) {
        int rc = pl->sec_code_bytes == 2 ? upload_tiled<uint16_t>(pl, &pl->codes, d->sec_codes, n, padded, R)
                                         : upload_tiled<int32_t>(pl, &pl->codes, d->sec_codes, n, padded, R);
        if (rc) return rc;
    }

    // ---- primary segments and tile keys ------------------------------------------------------------
    std::vector<int64_t> seg(d->n_primary + 1, 0);
    if (d->n_primary > 0) {
        std::memcpy(seg.data(), d->seg_offsets, seg.size() * sizeof(int64_t));
        if (seg.front() != 0 || seg.back() != d->n_rows) return fail(SKK_ERR_INVALID, "seg_offsets must span [0, n_rows]");
        for (size_t k = 1; k < seg.size(); ++k)
            if (seg[k] < seg[k - 1]) return fail(SKK_ERR_INVALID, "seg_offsets must be non-decreasing");
    } else {
        seg.back() = d->n_rows;
    }
    {
        int rc = dev_upload(pl, &pl->seg_off, seg.data(), seg.size());
        if (rc) return rc;
    }
    std::vector<int2> keys(pl->n_tiles + 1, make_int2(0, 0));
    if (d->n_primary > 0) {
        for (int t = 0; t < pl->n_tiles; ++t) {
            const int64_t r0 = (int64_t)t * WT, r1 = std::min<int64_t>(d->n_rows, r0 + WT) - 1;
            const int kf = (int)(std::upper_bound(seg.begin(), seg.end(), r0) - seg.begin()) - 1;
            const int kl = (int)(std::upper_bound(seg.begin(), seg.end(), std::max(r0, r1)) - seg.begin()) - 1;
            keys[t] = make_int2(std::min<int>(kf, (int)d->n_primary - 1), std::min<int>(kl, (int)d->n_primary - 1));
        }
    }
    // ---- terms: value-table layout and the reduce CSR ------------------------------------------------
    struct Entry {
        int64_t theta;
        int8_t level, t;
        int32_t slot;
    };
    std::vector<Entry> entries;
    int cnt[3] = {0, 0, 0};
    pl->n_terms = d->n_terms;
    const int32_t *host_ti_p[kMaxTermsPerLevel] = {nullptr, nullptr};
    for (int i = 0; i < d->n_terms; ++i) {
        const skk_term_desc &td = d->terms[i];
        const int t = cnt[td.level]++;
        int32_t *ti = nullptr;
        int rc = dev_upload(pl, &ti, td.theta_index, (size_t)td.n_slots);
        if (rc) return rc;
        const int64_t slots = td.level == SKK_LEVEL_GLOBAL ? 1 : td.level == SKK_LEVEL_PRIMARY ? d->n_primary : d->n_secondary;
        (void)slots;
        if (td.level == SKK_LEVEL_PRIMARY) host_ti_p[t] = td.theta_index;
        if (td.level == SKK_LEVEL_SECONDARY && t == 0) {
            // slot -> theta is `base + slot`? (then the row kernel fills its value table straight from theta)
            bool affine = td.n_slots > 0 && env_int("SKK_NO_AFFINE_TABLE", 0) == 0;
            for (int64_t k = 0; affine && k < td.n_slots; ++k) affine = td.theta_index[k] == td.theta_index[0] + k;
            pl->sec_theta_base = affine ? td.theta_index[0] : -1;
        }
        pl->theta_index_dev[i] = ti;
        pl->term_level[i] = td.level;
        pl->term_slot_in_level[i] = t;
        (td.level == SKK_LEVEL_GLOBAL ? pl->xcol_g : td.level == SKK_LEVEL_PRIMARY ? pl->xcol_p : pl->xcol_s)[t] = td.xcol;
        for (int64_t s = 0; s < td.n_slots; ++s) {
            if (td.theta_index[s] < 0 || td.theta_index[s] >= d->n_params)
                return fail(SKK_ERR_INVALID, "term %d: theta_index out of range", i);
            entries.push_back(Entry{td.theta_index[s], (int8_t)td.level, (int8_t)t, (int32_t)s});
        }
    }
    // tile meta: first / last primary slot and the theta elements of the primary terms at the first slot
    {
        std::vector<int4> meta(keys.size());
        for (size_t t = 0; t < keys.size(); ++t)
            meta[t] = make_int4(keys[t].x, keys[t].y, host_ti_p[0] ? host_ti_p[0][keys[t].x] : 0,
                                host_ti_p[1] ? host_ti_p[1][keys[t].x] : 0);
        if (int rc = dev_upload(pl, &pl->tile_meta, meta.data(), meta.size())) return rc;
        // tiles that hold exactly two groups: where the second one starts (row inside the tile)
        std::vector<int32_t> split(keys.size(), -1);
        if (d->n_primary > 0)
            for (int t = 0; t < pl->n_tiles; ++t)
                if (keys[t].y == keys[t].x + 1) split[t] = (int32_t)(seg[keys[t].y] - (int64_t)t * WT);
        if (int rc = dev_upload(pl, &pl->tile_split, split.data(), split.size())) return rc;
    }

    // theta element -> units of the `totals` buffer ([NG+1 global | NP*GP primary | NS*GS secondary])
    std::stable_sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.theta < b.theta; });
    auto unit_of = [&](const Entry &e) -> int64_t {
        if (e.level == SKK_LEVEL_GLOBAL) return e.t;
        if (e.level == SKK_LEVEL_PRIMARY) return (pl->ng + 1) + (int64_t)e.t * d->n_primary + e.slot;
        return (pl->ng + 1) + (int64_t)pl->np * d->n_primary + (int64_t)e.t * d->n_secondary + e.slot;
    };
    if ((pl->ng + 1) + (int64_t)pl->np * d->n_primary + (int64_t)pl->ns * d->n_secondary >= INT32_MAX)
        return fail(SKK_ERR_UNSUPPORTED, "too many group slots");
    std::vector<int32_t> single_unit(d->n_params, -1), multi_elem, multi_unit;
    std::vector<int64_t> multi_ptr(1, 0);
    for (size_t k = 0; k < entries.size();) {
        size_t e = k;
        while (e < entries.size() && entries[e].theta == entries[k].theta) ++e;
        if (e - k == 1) {
            single_unit[entries[k].theta] = (int32_t)unit_of(entries[k]);
        } else {
            single_unit[entries[k].theta] = -2;
            multi_elem.push_back((int32_t)entries[k].theta);
            for (size_t q = k; q < e; ++q) multi_unit.push_back((int32_t)unit_of(entries[q]));
            multi_ptr.push_back((int64_t)multi_unit.size());
        }
        k = e;
    }
    std::vector<int32_t> multi_of(d->n_params, -1);
    for (size_t m = 0; m < multi_elem.size(); ++m) multi_of[multi_elem[m]] = (int32_t)m;
    int32_t *multi_of_dev = nullptr;
    if (int rc = dev_upload(pl, &multi_of_dev, multi_of.data(), multi_of.size())) return rc;
    int32_t *single_dev = nullptr, *multi_elem_dev = nullptr, *multi_unit_dev = nullptr;
    int64_t *multi_ptr_dev = nullptr;
    if (int rc = dev_upload(pl, &single_dev, single_unit.data(), single_unit.size())) return rc;
    if (int rc = dev_upload(pl, &multi_elem_dev, multi_elem.data(), multi_elem.size())) return rc;
    if (int rc = dev_upload(pl, &multi_unit_dev, multi_unit.data(), multi_unit.size())) return rc;
    if (int rc = dev_upload(pl, &multi_ptr_dev, multi_ptr.data(), multi_ptr.size())) return rc;

    // primary slots: tile range, and which partial of the first / last tile belongs to the slot
    std::vector<int4> slot_desc(std::max<int64_t>(1, d->n_primary), make_int4(0, -1, 0, 0));
    for (int64_t sl = 0; sl < d->n_primary; ++sl) {
        const int64_t r0 = seg[sl], r1 = seg[sl + 1];
        if (r1 <= r0) continue;
        const int t_lo = (int)(r0 / WT), t_hi = (int)((r1 - 1) / WT);
        auto kind = [&](int tile) { return keys[tile].x == sl ? 1 : (keys[tile].y == sl ? 2 : 0); };
        slot_desc[sl] = make_int4(t_lo, t_hi, kind(t_lo), kind(t_hi));
    }
    int4 *slot_desc_dev = nullptr;
    if (int rc = dev_upload(pl, &slot_desc_dev, slot_desc.data(), slot_desc.size())) return rc;

    // elements fed by several slots: per term whether none / all / some of its elements are such, and the raw
    // partials of their primary slots as one flat list per element (the finish kernel's warp adds them)
    int8_t term_multi[3][kMaxTermsPerLevel] = {{0, 0}, {0, 0}, {0, 0}};
    {
        int64_t n_of[3][kMaxTermsPerLevel] = {{0, 0}, {0, 0}, {0, 0}}, multi_of_term[3][kMaxTermsPerLevel] = {{0, 0}, {0, 0}, {0, 0}};
        for (const Entry &e : entries) {
            ++n_of[e.level][e.t];
            if (single_unit[e.theta] == -2) ++multi_of_term[e.level][e.t];
        }
        for (int lv = 0; lv < 3; ++lv)
            for (int t = 0; t < kMaxTermsPerLevel; ++t)
                term_multi[lv][t] = multi_of_term[lv][t] == 0 ? 0 : (multi_of_term[lv][t] == n_of[lv][t] ? 1 : 2);
    }
    std::vector<int64_t> msrc_ptr(1, 0);
    std::vector<uint32_t> msrc;
    {
        const int64_t first_p = pl->ng + 1, first_s = first_p + (int64_t)pl->np * d->n_primary;
        if ((int64_t)pl->n_tiles * std::max(1, pl->np) >= (1 << 30) || (int64_t)pl->np * d->n_primary >= (1 << 30))
            return fail(SKK_ERR_UNSUPPORTED, "too many tiles for the flattened partial lists");
        for (size_t m = 0; m < multi_elem.size(); ++m) {
            for (int64_t q = multi_ptr[m]; q < multi_ptr[m + 1]; ++q) {
                const int64_t unit = multi_unit[q];
                if (unit < first_p || unit >= first_s) continue;
                const int t = (int)((unit - first_p) / d->n_primary);
                const int64_t sl = (unit - first_p) % d->n_primary;
                const int4 sd = slot_desc[sl];
                if (sd.y < sd.x) continue;  // empty group
                // (direct_p holds something only for a group that lies inside one tile and is either an inner
                // group of it, or begins on the tile's first row -- such a group may start and end inside lane 0,
                // which then stores it directly and leaves tile_head zero; everywhere else it is zero for good)
                if (sd.x == sd.y && (sd.z == 0 || seg[sl] % WT == 0)) msrc.push_back((uint32_t)(t * d->n_primary + sl));
                for (int tile = sd.x; tile <= sd.y; ++tile) {
                    const int kind = tile == sd.x ? sd.z : (tile == sd.y ? sd.w : 1);
                    if (kind) msrc.push_back(((uint32_t)kind << 30) | (uint32_t)(tile * pl->np + t));
                }
            }
            msrc_ptr.push_back((int64_t)msrc.size());
        }
    }
    int64_t *msrc_ptr_dev = nullptr;
    uint32_t *msrc_dev = nullptr;
    if (int rc = dev_upload(pl, &msrc_ptr_dev, msrc_ptr.data(), msrc_ptr.size())) return rc;
    if (int rc = dev_upload(pl, &msrc_dev, msrc.data(), msrc.size())) return rc;

    // ---- priors, flattened per theta element; hyper-parameter links as contiguous ranges -----------------
    const size_t Pn = (size_t)d->n_params;
    std::vector<int8_t> fam(Pn, -1);
    std::vector<int32_t> a_src(Pn, -1), b_src(Pn, -1), e_mu(Pn, -1), e_sigma(Pn, -1);
    std::vector<double> a_const(Pn, 0.0), b_const(Pn, 1.0);
    struct Link {
        int64_t parent;
        int32_t child;
        int8_t kind;
    };
    std::vector<Link> links;
    // family of the block every theta element belongs to (a linked parameter may refer to any other block)
    std::vector<int8_t> elem_family(Pn, -1);
    bool any_extended = false;  // a family the row kernel's prologue does not evaluate
    for (int k = 0; k < d->n_blocks; ++k) {
        const skk_block_desc &bd = d->blocks[k];
        if (bd.offset < 0 || bd.size < 0 || bd.offset + bd.size > d->n_params)
            return fail(SKK_ERR_INVALID, "block %d outside theta", k);
        if (bd.family != SKK_PRIOR_NORMAL && bd.family != SKK_PRIOR_HALFNORMAL && bd.family != SKK_PRIOR_EXPONENTIAL &&
            !(bd.family >= SKK_PRIOR_HALFCAUCHY && bd.family <= SKK_PRIOR_HALFSTUDENTT))
            return fail(SKK_ERR_UNSUPPORTED, "block %d: unknown prior family %d", k, bd.family);
        for (int64_t g = 0; g < bd.size; ++g) {
            if (elem_family[bd.offset + g] != -1)
                return fail(SKK_ERR_INVALID, "blocks overlap at theta[%lld]", (long long)(bd.offset + g));
            elem_family[bd.offset + g] = (int8_t)bd.family;
        }
    }
    for (int k = 0; k < d->n_blocks; ++k) {
        const skk_block_desc &bd = d->blocks[k];
        const bool normal = bd.family == SKK_PRIOR_NORMAL;
        // families with a second (constant) parameter
        const bool two = normal || (prior_is_extended(bd.family) && bd.family != SKK_PRIOR_HALFCAUCHY);
        if (prior_is_extended(bd.family)) any_extended = true;
        if (!normal && (bd.a_index || bd.b_index))
            return fail(SKK_ERR_UNSUPPORTED, "block %d: only Normal priors take random parameters", k);
        // (index and constants together: per element the constant where the index is -1 -- the member-wise
        // parameters of a GroupComponent, /root/reference/sakkara/model/composable/group.py:69-103)
        if (bd.a_index ? (bd.a_const != nullptr && bd.a_const_len != bd.size) : (bd.a_const_len != 1 && bd.a_const_len != bd.size))
            return fail(SKK_ERR_INVALID, "block %d: bad a_const_len", k);
        if (two && (bd.b_index ? (bd.b_const != nullptr && bd.b_const_len != bd.size) : (bd.b_const_len != 1 && bd.b_const_len != bd.size)))
            return fail(SKK_ERR_INVALID, "block %d: bad b_const_len", k);
        for (int64_t g = 0; g < bd.size; ++g) {
            const int64_t j = bd.offset + g;
            fam[j] = (int8_t)bd.family;
            if (bd.a_index && (bd.a_index[g] >= 0 || bd.a_const == nullptr)) {
                const int64_t src = bd.a_index[g];
                if (src < 0 || src >= d->n_params) return fail(SKK_ERR_INVALID, "prior link outside theta");
                // a mu that is a positive variable is sampled as its log: mu = exp(theta[src])
                const bool mu_is_log = prior_is_log(elem_family[src]);
                if (mu_is_log) fam[j] = (int8_t)(fam[j] | kPriorMuIsLog);
                a_src[j] = (int32_t)src;
                links.push_back(Link{src, (int32_t)j, 0});
            } else {
                a_const[j] = bd.a_const[bd.a_const_len == 1 ? 0 : g];
            }
            if (two) {
                if (bd.b_index && (bd.b_index[g] >= 0 || bd.b_const == nullptr)) {
                    const int64_t src = bd.b_index[g];
                    if (src < 0 || src >= d->n_params) return fail(SKK_ERR_INVALID, "prior link outside theta");
                    if (!prior_is_log(elem_family[src]))
                        return fail(SKK_ERR_INVALID, "block %d: a linked sigma must be a positive (log-transformed) variable", k);
                    b_src[j] = (int32_t)src;
                    links.push_back(Link{src, (int32_t)j, 1});
                } else {
                    b_const[j] = bd.b_const[bd.b_const_len == 1 ? 0 : g];
                }
            }
        }
    }
    for (size_t j = 0; j < Pn; ++j)
        if (fam[j] < 0) return fail(SKK_ERR_INVALID, "theta[%lld] belongs to no block", (long long)j);
    for (const Link &l : links)
        if (l.parent < 0 || l.parent >= d->n_params) return fail(SKK_ERR_INVALID, "prior link outside theta");
    std::stable_sort(links.begin(), links.end(), [](const Link &a, const Link &b) { return a.parent < b.parent; });
    // parents = elements with children; those with long lists first (a whole CTA each)
    struct Parent {
        int32_t element;
        int64_t lo, hi;
    };
    std::vector<Parent> parents;
    for (size_t k = 0; k < links.size();) {
        size_t e = k;
        while (e < links.size() && links[e].parent == links[k].parent) ++e;
        parents.push_back(Parent{(int32_t)links[k].parent, (int64_t)k, (int64_t)e});
        k = e;
    }
    std::stable_sort(parents.begin(), parents.end(), [](const Parent &a, const Parent &b) {
        return (a.hi - a.lo > kHeavyLinks) > (b.hi - b.lo > kHeavyLinks);
    });
    int n_heavy = 0;
    std::vector<int32_t> parent_el(parents.size());
    std::vector<int64_t> link_ptr(parents.size() + 1, 0);
    {
        int64_t w = 0;
        for (size_t q = 0; q < parents.size(); ++q) {
            if (parents[q].hi - parents[q].lo > kHeavyLinks) ++n_heavy;
            parent_el[q] = parents[q].element;
            link_ptr[q] = w;
            for (int64_t e = parents[q].lo; e < parents[q].hi; ++e, ++w)
                (links[e].kind == 0 ? e_mu : e_sigma)[links[e].child] = (int32_t)w;
        }
        link_ptr[parents.size()] = w;
    }
    int8_t *fam_dev = nullptr;
    int32_t *a_src_dev = nullptr, *b_src_dev = nullptr, *e_mu_dev = nullptr, *e_sigma_dev = nullptr, *parent_dev = nullptr;
    double *a_const_dev = nullptr, *b_const_dev = nullptr;
    int64_t *link_ptr_dev = nullptr;
    if (int rc = dev_upload(pl, &fam_dev, fam.data(), fam.size())) return rc;
    if (int rc = dev_upload(pl, &a_src_dev, a_src.data(), a_src.size())) return rc;
    if (int rc = dev_upload(pl, &b_src_dev, b_src.data(), b_src.size())) return rc;
    if (int rc = dev_upload(pl, &a_const_dev, a_const.data(), a_const.size())) return rc;
    if (int rc = dev_upload(pl, &b_const_dev, b_const.data(), b_const.size())) return rc;
    if (int rc = dev_upload(pl, &e_mu_dev, e_mu.data(), e_mu.size())) return rc;
    if (int rc = dev_upload(pl, &e_sigma_dev, e_sigma.data(), e_sigma.size())) return rc;
    if (int rc = dev_upload(pl, &parent_dev, parent_el.data(), parent_el.size())) return rc;
    if (int rc = dev_upload(pl, &link_ptr_dev, link_ptr.data(), link_ptr.size())) return rc;
    pl->prior = PriorParams{fam_dev,    a_src_dev,    b_src_dev, a_const_dev, b_const_dev, e_mu_dev, e_sigma_dev,
                            parent_dev, link_ptr_dev, (int32_t)parents.size(), n_heavy, d->n_params,
                            (int64_t)links.size(), 0};
    pl->prior_ctas = (int)((Pn + kPriorThreads - 1) / kPriorThreads);
    // Single GPU, batches below the chain kernel's: the row kernel evaluates the prior terms in its
    // prologue and one finish kernel does the rest.  That kernel finishes a hyper-parameter where it
    // sums over its children, so no term may feed one (otherwise the separate kernels run).
    std::vector<int32_t> none_elem;
    {
        std::vector<char> is_parent(Pn, 0);
        pl->fused_ok = !any_extended;
        for (const Parent &q : parents) {
            is_parent[q.element] = 1;
            if (single_unit[q.element] != -1 || (d->family == SKK_LIK_NORMAL && q.element == d->sigma_theta_index))
                pl->fused_ok = false;
        }
        for (size_t j = 0; j < Pn; ++j)
            if (single_unit[j] == -1 && !is_parent[j]) none_elem.push_back((int32_t)j);
    }
    int32_t *none_dev = nullptr;
    if (int rc = dev_upload(pl, &none_dev, none_elem.data(), none_elem.size())) return rc;
    // which entries of the exchanged vector this rank's rows feed (the logp partial and the constant always)
    pl->fed.assign(Pn + 2, 0);
    for (size_t j = 0; j < Pn; ++j) pl->fed[j] = single_unit[j] != -1 ? 1 : 0;
    pl->fed[Pn] = pl->fed[Pn + 1] = 1;
    std::vector<uint8_t> is_parent_host(std::max<size_t>(1, Pn), 0);
    for (const Parent &q : parents) is_parent_host[q.element] = 1;
    uint8_t *is_parent_dev = nullptr;
    if (int rc = dev_upload(pl, &is_parent_dev, is_parent_host.data(), is_parent_host.size())) return rc;
    // The same links once more in link order, every hyper-parameter's range padded to whole chunks of 32
    // and each link carrying its child's prior parameters: the row kernel's prologue reduces a chunk per
    // warp, the finish kernel adds the chunk partials of a hyper-parameter.
    std::vector<int32_t> lk_child, lk_a_src, lk_b_src;
    std::vector<int8_t> lk_kind;
    std::vector<double> lk_a_const, lk_b_const;
    std::vector<int64_t> chunk_ptr(parents.size() + 1, 0);
    for (size_t q = 0; q < parents.size(); ++q) {
        chunk_ptr[q] = (int64_t)(lk_child.size() / 32);
        for (int64_t e = parents[q].lo; e < parents[q].hi; ++e) {
            const int32_t j = links[e].child;
            lk_child.push_back(j);
            lk_kind.push_back((int8_t)(links[e].kind | ((fam[j] & kPriorMuIsLog) ? 4 : 0)));  // bit 2: the child's mu is exp(u)
            lk_a_src.push_back(a_src[j]);
            lk_b_src.push_back(b_src[j]);
            lk_a_const.push_back(a_const[j]);
            lk_b_const.push_back(b_const[j]);
        }
        while (lk_child.size() % 32) {
            lk_child.push_back(-1);
            lk_kind.push_back(0);
            lk_a_src.push_back(-1);
            lk_b_src.push_back(-1);
            lk_a_const.push_back(0.0);
            lk_b_const.push_back(1.0);
        }
    }
    chunk_ptr[parents.size()] = (int64_t)(lk_child.size() / 32);
    pl->n_chunks = (int64_t)(lk_child.size() / 32);
    int64_t *chunk_ptr_dev = nullptr;
    if (int rc = dev_upload(pl, &pl->lk_child, lk_child.data(), lk_child.size())) return rc;
    if (int rc = dev_upload(pl, &pl->lk_kind, lk_kind.data(), lk_kind.size())) return rc;
    if (int rc = dev_upload(pl, &pl->lk_a_src, lk_a_src.data(), lk_a_src.size())) return rc;
    if (int rc = dev_upload(pl, &pl->lk_b_src, lk_b_src.data(), lk_b_src.size())) return rc;
    if (int rc = dev_upload(pl, &pl->lk_a_const, lk_a_const.data(), lk_a_const.size())) return rc;
    if (int rc = dev_upload(pl, &pl->lk_b_const, lk_b_const.data(), lk_b_const.size())) return rc;
    if (int rc = dev_upload(pl, &chunk_ptr_dev, chunk_ptr.data(), chunk_ptr.size())) return rc;

    // ---- launch geometry -------------------------------------------------------------------------------------
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, pl->device));
    pl->n_sms = prop.multiProcessorCount;
    const int smem_limit = (int)prop.sharedMemPerBlockOptin;
    pl->warps = std::max(1, std::min(8, env_int("SKK_WARPS", 8)));  // kernels are compiled for <= 256 threads
    pl->stages = std::max(2, std::min(3, env_int("SKK_STAGES", 3)));   // (the kernel looks three tiles ahead in its schedule)
    const int bt_max = d->max_batch > 1 ? 4 : 1;
    while (cta_smem<T>(pl, pl->warps, pl->stages, bt_max) > smem_limit) {
        if (pl->stages > 2) --pl->stages;
        else if (pl->warps > 1) pl->warps /= 2;
        else
            return fail(SKK_ERR_UNSUPPORTED, "secondary level with %lld members does not fit in shared memory",
                        (long long)pl->n_secondary);
    }
    // resident CTAs per SM hide more latency than a third pipeline stage: prefer the depth that fits more CTAs
    if (!getenv("SKK_STAGES") && pl->stages > 2) {
        const int per_sm = (int)prop.sharedMemPerMultiprocessor;
        // (the kernels are compiled for two CTAs per SM: 128 registers x 256 threads)
        auto fit = [&](int stages) { return std::min(2, per_sm / (cta_smem<T>(pl, pl->warps, stages, 1) + 1024)); };
        if (fit(2) > fit(pl->stages)) pl->stages = 2;
    }
    if (int rc = configure_launch<T>(pl, 1, &pl->cfg1)) return rc;
    if (bt_max == 4)
        if (int rc = configure_launch<T>(pl, 4, &pl->cfg4)) return rc;
    const int grid_max = std::max(pl->cfg1.grid, pl->cfg4.grid);
    {
        const int64_t n_prior_groups = ((int64_t)d->n_params + 31) / 32;
        if (int rc = build_schedule(pl, keys, d->n_rows, n_prior_groups, pl->n_chunks, &pl->cfg1)) return rc;
        if (bt_max == 4)
            if (int rc = build_schedule(pl, keys, d->n_rows, n_prior_groups, pl->n_chunks, &pl->cfg4)) return rc;
    }

    // ---- work buffers --------------------------------------------------------------------------------------------
    const size_t B = (size_t)pl->max_batch, P = (size_t)d->n_params, bt = (size_t)bt_max;
    T *tmp = nullptr;
    if (int rc = dev_alloc(pl, &tmp, B * P)) return rc;
    pl->theta_dev = tmp;
    const size_t units = (size_t)(pl->ng + 1) + (size_t)pl->np * d->n_primary + (size_t)pl->ns * d->n_secondary;
    if (int rc = dev_alloc(pl, &pl->totals, units * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->direct_p, std::max<size_t>(1, pl->np) * std::max<int64_t>(1, d->n_primary) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->direct_p1, std::max<size_t>(1, pl->np) * std::max<int64_t>(1, d->n_primary), true)) return rc;
    if (int rc = dev_alloc(pl, &pl->tile_head, (size_t)pl->n_tiles * std::max(1, pl->np) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->tile_tail, (size_t)pl->n_tiles * std::max(1, pl->np) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->cta_glob, (size_t)grid_max * (pl->ng + 1) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->cta_sec, (size_t)grid_max * std::max(1, pl->ns) * std::max<int64_t>(1, d->n_secondary) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->lik_vec, B * (P + 2), true)) return rc;
    if (int rc = dev_alloc(pl, &pl->prior_grad, B * P, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->prior_lp, B * std::max<size_t>(pl->prior_ctas + 1, (size_t)grid_max * 8), true)) return rc;
    if (int rc = dev_alloc(pl, &pl->link_val, B * std::max<size_t>(1, links.size()), true)) return rc;
    if (int rc = dev_alloc(pl, &pl->link_part, B * std::max<size_t>(1, (size_t)pl->n_chunks), true)) return rc;
    if (int rc = dev_alloc(pl, &tmp, B * P + B, true)) return rc;
    pl->out_dev = tmp;
    CUDA_TRY(cudaMallocHost(&pl->h_in, std::max<size_t>(16, B * P * sizeof(T))));
    CUDA_TRY(cudaMallocHost(&pl->h_out, std::max<size_t>(16, (B * P + B) * sizeof(T))));

    ReduceParams &rp = pl->reduce;
    rp = ReduceParams{};
    rp.slot_desc = slot_desc_dev;
    rp.tile_head = pl->tile_head;
    rp.tile_tail = pl->tile_tail;
    rp.cta_glob = pl->cta_glob;
    rp.cta_sec = pl->cta_sec;
    rp.totals = pl->totals;
    rp.direct_p = pl->direct_p;
    rp.single_unit = single_dev;
    rp.multi_elem = multi_elem_dev;
    rp.multi_ptr = multi_ptr_dev;
    rp.multi_unit = multi_unit_dev;
    rp.n_multi = (int32_t)multi_elem.size();
    rp.none_elem = none_dev;
    rp.n_none = (int32_t)none_elem.size();
    rp.none_blocks = (int32_t)((none_elem.size() + 255) / 256);
    rp.multi_blocks = (int32_t)((multi_elem.size() + 7) / 8);
    for (int t = 0; t < kMaxTermsPerLevel; ++t) {
        rp.multi_g[t] = term_multi[SKK_LEVEL_GLOBAL][t];
        rp.multi_p[t] = term_multi[SKK_LEVEL_PRIMARY][t];
        rp.multi_s[t] = term_multi[SKK_LEVEL_SECONDARY][t];
    }
    rp.multi_src_ptr = msrc_ptr_dev;
    rp.multi_src = msrc_dev;
    rp.is_parent = is_parent_dev;
    {
        rp.parent = pl->prior.parent;
        rp.chunk_ptr = chunk_ptr_dev;
        rp.link_part = pl->link_part;
        rp.n_chunks = pl->n_chunks;
        rp.n_parents = pl->prior.n_parents;
        rp.n_heavy = pl->prior.n_heavy;
        rp.parent_blocks = rp.n_heavy + (rp.n_parents - rp.n_heavy + 7) / 8;
    }
    for (int i = 0; i < d->n_terms; ++i) {
        const int t = pl->term_slot_in_level[i];
        (pl->term_level[i] == SKK_LEVEL_GLOBAL ? rp.ti_g : pl->term_level[i] == SKK_LEVEL_PRIMARY ? rp.ti_p : rp.ti_s)[t] =
            pl->theta_index_dev[i];
    }
    rp.theta_blocks = (int32_t)std::max<size_t>(1, (P + 255) / 256);
    rp.lik_vec = pl->lik_vec;
    rp.n_params = d->n_params;
    rp.n_primary = d->n_primary;
    rp.n_secondary = d->n_secondary;
    rp.ng = pl->ng;
    rp.np = pl->np;
    rp.ns = pl->ns;
    rp.blocks_p = pl->np > 0 ? (int)((d->n_primary + 255) / 256) : 0;
    rp.blocks_s = pl->ns > 0 ? (int)((d->n_secondary + 31) / 32) : 0;
    rp.logp_const = d->logp_const;
    rp.prior_grad = pl->prior_grad;
    rp.prior_lp_part = pl->prior_lp;
    rp.prior_ctas = pl->prior_ctas;
    rp.include_priors = pl->include_priors;
    rp.n_rows_total = pl->n_rows_total;
    rp.sigma_theta_index = pl->sigma_idx;
    rp.sigma_const = pl->sigma_const;
    rp.family = pl->family;
    pl->use_graphs = env_int("SKK_NO_GRAPH", 0) == 0;

    // ---- chain-per-lane path: batches of >= kChainMinBatch vectors, models without a secondary level ----------
    if (pl->ns == 0 && pl->max_batch >= kChainMinBatch && env_int("SKK_NO_CHAIN", 0) == 0 && pl->family != SKK_LIK_NORMAL_LOGSIGMA &&
        pl->family != SKK_LIK_STUDENTT_LOGSIGMA) {
        pl->chain_fn = pl->dtype == SKK_F64 ? chain_kernel_f64(pl->family, pl->ng, pl->np)
                                            : chain_kernel_f32(pl->family, pl->ng, pl->np);
        if (!pl->chain_fn) return fail(SKK_ERR_UNSUPPORTED, "no chain kernel for this shape");
        const int TR = kChainTileRows;
        pl->chain_tiles = (int)std::max<int64_t>(1, (d->n_rows + TR - 1) / TR);
        pl->chain_bp = (pl->max_batch + 31) & ~31;
        const int gy = (pl->chain_bp / 32 + kChainWarps - 1) / kChainWarps;
        pl->chain_ranges = std::max(1, std::min(pl->chain_tiles, std::max(1, pl->n_sms * 4 / gy)));
        const int RR = pl->chain_ranges;
        const size_t es = sizeof(T);
        const int stage = (int)(((size_t)pl->n_x * TR * es + (pl->y_u8 ? TR : TR * es) + (pl->has_off ? TR * es : 0) + 127) & ~(size_t)127);
        pl->chain_smem = 128 + 2 * stage;
        CUDA_TRY(cudaFuncSetAttribute(pl->chain_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->chain_smem));
        std::vector<int2> rk(RR, make_int2(0, 0));
        std::vector<int64_t> rstart(RR + 1, 0);
        for (int c = 0; c <= RR; ++c) rstart[c] = std::min<int64_t>(d->n_rows, ((int64_t)pl->chain_tiles * c / RR) * TR);
        if (d->n_primary > 0)
            for (int c = 0; c < RR; ++c) {
                if (rstart[c] >= rstart[c + 1]) continue;
                const int kf = (int)(std::upper_bound(seg.begin(), seg.end(), rstart[c]) - seg.begin()) - 1;
                const int kl = (int)(std::upper_bound(seg.begin(), seg.end(), rstart[c + 1] - 1) - seg.begin()) - 1;
                rk[c] = make_int2(std::min<int>(kf, (int)d->n_primary - 1), std::min<int>(kl, (int)d->n_primary - 1));
            }
        std::vector<int4> cdesc(std::max<int64_t>(1, d->n_primary), make_int4(0, -1, 0, 0));
        for (int64_t sl = 0; sl < d->n_primary; ++sl) {
            const int64_t q0 = seg[sl], q1 = seg[sl + 1];
            if (q1 <= q0) continue;
            const int c_lo = (int)(std::upper_bound(rstart.begin(), rstart.end(), q0) - rstart.begin()) - 1;
            const int c_hi = (int)(std::upper_bound(rstart.begin(), rstart.end(), q1 - 1) - rstart.begin()) - 1;
            auto kind = [&](int c) { return rk[c].x == sl ? 1 : (rk[c].y == sl ? 2 : 0); };
            cdesc[sl] = make_int4(c_lo, c_hi, kind(c_lo), kind(c_hi));
        }
        int4 *cdesc_dev = nullptr;
        if (int rc = dev_upload(pl, &pl->chain_keys, rk.data(), rk.size())) return rc;
        if (int rc = dev_upload(pl, &cdesc_dev, cdesc.data(), cdesc.size())) return rc;
        const size_t bp = (size_t)pl->chain_bp;
        const size_t cunits = (size_t)(pl->ng + 1) + (size_t)pl->np * d->n_primary;
        if (int rc = dev_alloc(pl, &pl->chain_totals, cunits * bp, true)) return rc;
        if (int rc = dev_alloc(pl, &pl->chain_head, (size_t)RR * std::max(1, pl->np) * bp, true)) return rc;
        if (int rc = dev_alloc(pl, &pl->chain_tail, (size_t)RR * std::max(1, pl->np) * bp, true)) return rc;
        if (int rc = dev_alloc(pl, &pl->chain_glob, (size_t)RR * (pl->ng + 1) * bp, true)) return rc;
        ChainReduceParams &cr = pl->chain_reduce;
        cr = ChainReduceParams{};
        cr.slot_desc = cdesc_dev;
        cr.totals = pl->chain_totals;
        cr.head = pl->chain_head;
        cr.tail = pl->chain_tail;
        cr.glob = pl->chain_glob;
        cr.single_unit = single_dev;
        cr.multi_elem = multi_elem_dev;
        cr.multi_ptr = multi_ptr_dev;
        cr.multi_unit = multi_unit_dev;
        cr.multi_of = multi_of_dev;
        cr.ranges = RR;
        cr.ng = pl->ng;
        cr.np = pl->np;
        cr.bp = pl->chain_bp;
        cr.n_primary = d->n_primary;
        cr.n_params = d->n_params;
        cr.prior_grad = pl->prior_grad;
        cr.prior_lp_part = pl->prior_lp;
        cr.prior_ctas = pl->prior_ctas;
        cr.include_priors = pl->include_priors;
        cr.family = pl->family;
        cr.n_rows_total = pl->n_rows_total;
        cr.sigma_theta_index = pl->sigma_idx;
        cr.sigma_const = pl->sigma_const;
        cr.logp_const = d->logp_const;
        pl->chain_ready = true;
    }
    return SKK_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluation
// ---------------------------------------------------------------------------------------------
template <typename T>
static int launch_rows(skk_plan *pl, const LaunchCfg &cfg, bool timed, const T *th, int b_valid, int bt, int b0,
                       bool with_priors) {
    RowParams<T> rp{};
    for (int c = 0; c < pl->n_x; ++c) rp.x[c] = static_cast<const T *>(pl->x[c]);
    rp.y = pl->y_u8 ? nullptr : static_cast<const T *>(pl->y);
    rp.y8 = pl->y_u8 ? static_cast<const uint8_t *>(pl->y) : nullptr;
    rp.eta_offset = static_cast<const T *>(pl->off);
    rp.sec_codes = pl->codes;
    rp.n_x = pl->n_x;
    rp.sec_code_bytes = pl->sec_code_bytes;
    rp.n_rows = pl->n_rows;
    rp.n_tiles = pl->n_tiles;
    rp.stages = pl->stages;
    rp.sched = cfg.sched;
    rp.sched_len = cfg.sched_len;
    rp.sched_folds = cfg.sched_folds;
    rp.tile_meta = pl->tile_meta;
    rp.tile_split = pl->tile_split;
    rp.seg_offsets = pl->seg_off;
    rp.n_primary = (int32_t)pl->n_primary;
    rp.n_secondary = (int32_t)pl->n_secondary;
    for (int t = 0; t < kMaxTermsPerLevel; ++t) {
        rp.xcol_g[t] = pl->xcol_g[t];
        rp.xcol_p[t] = pl->xcol_p[t];
        rp.xcol_s[t] = pl->xcol_s[t];
    }
    rp.theta = th;
    rp.fam_a = (T)((pl->family == SKK_LIK_NEGBINOMIAL_LOG || pl->family == SKK_LIK_STUDENTT_LOGSIGMA || pl->family == SKK_LIK_GAMMA_LOG)
                       ? pl->sigma_const : 0.0);
    rp.fam_b = (T)(pl->family == SKK_LIK_NEGBINOMIAL_LOG ? log(pl->sigma_const) : 0.0);
    rp.sec_theta_base = pl->sec_theta_base;
    rp.n_params = pl->n_params;
    rp.b_valid = b_valid;
    for (int i = 0; i < pl->n_terms; ++i) {
        const int t = pl->term_slot_in_level[i];
        (pl->term_level[i] == SKK_LEVEL_GLOBAL ? rp.ti_g : pl->term_level[i] == SKK_LEVEL_PRIMARY ? rp.ti_p : rp.ti_s)[t] =
            pl->theta_index_dev[i];
    }
    rp.direct_p = bt == 4 ? pl->direct_p : pl->direct_p1;
    rp.tile_head = pl->tile_head;
    rp.tile_tail = pl->tile_tail;
    rp.cta_glob = pl->cta_glob;
    rp.cta_sec = pl->cta_sec;
    if (with_priors) {
        const PriorParams &pp = pl->prior;
        rp.prior_family = pp.family;
        rp.prior_a_src = pp.a_src;
        rp.prior_b_src = pp.b_src;
        rp.prior_a_const = pp.a_const;
        rp.prior_b_const = pp.b_const;
        rp.prior_grad = pl->prior_grad + (size_t)b0 * pl->n_params;
        rp.prior_lp_part = pl->prior_lp + (size_t)b0 * cfg.grid * pl->warps;
        rp.link_child = pl->lk_child;
        rp.link_kind = pl->lk_kind;
        rp.link_a_src = pl->lk_a_src;
        rp.link_b_src = pl->lk_b_src;
        rp.link_a_const = pl->lk_a_const;
        rp.link_b_const = pl->lk_b_const;
        rp.link_part = pl->link_part + (size_t)b0 * pl->n_chunks;
        rp.n_chunks = pl->n_chunks;
    }
    rp.trace = pl->trace;
    rp.peer_seq = (pl->peer_ready && b0 == 0) ? pl->peer.seq : nullptr;  // the evaluation's first row kernel counts it
    rp.pdl_trigger = env_int("SKK_NO_PDL", 0) ? 0 : env_int("SKK_PDL_TRIGGER", 2);
    void *args[] = {&rp};
    if (timed) CUDA_TRY(cudaEventRecord(pl->evk0, pl->stream));
    CUDA_TRY(cudaLaunchKernel(cfg.fn, dim3(cfg.grid), dim3(pl->warps * 32), args, cfg.smem, pl->stream));
    if (timed) CUDA_TRY(cudaEventRecord(pl->evk1, pl->stream));
    return SKK_OK;
}

// Launch of a kernel that follows the row kernel in the stream.  With `pdl` it may start while the row kernel's last
// CTAs are still running (programmatic dependent launch): its CTAs load their descriptors and then wait for the
// row kernel's completion inside the kernel (grid_dep_wait), which hides the launch gap and the first level of
// dependent loads.  Captured into the evaluation's CUDA graph like any other launch.
template <typename... KArgs, typename... Args>
static cudaError_t launch_after_rows(void (*kernel)(KArgs...), unsigned blocks, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Enqueues one evaluation on pl->stream (+ the prior kernels on pl->side, forked and joined with
// events).  Used directly, or recorded once into a CUDA graph and replayed.
//   theta_in/out pointers are device pointers here; returns the number of kernels enqueued.
template <typename T>
static int enqueue_eval(skk_plan *pl, const T *th, int batch, T *out_logp, T *out_grad, bool timed, int *n_launch) {
    const size_t P = (size_t)pl->n_params;
    cudaStream_t st = pl->stream;
    int launches = 0;
    const bool sharded = pl->comm != nullptr || pl->peer_ready;

    // K3: priors for the whole batch, next to the row kernel
    const bool priors = pl->include_priors && P > 0;
    const bool chain = pl->chain_ready && batch >= kChainMinBatch && !sharded;
    // single GPU, row kernel: the priors ride in the row kernel's prologue and one finish kernel follows
    const bool fused = !sharded && !chain && pl->fused_ok && env_int("SKK_NO_FUSED_FINISH", 0) == 0;
    // a row-sharded run likewise: priors in the row kernel's prologue, then one kernel that hands the likelihood
    // parts to the exchange -- over peer memory that kernel also waits for the other ranks and writes the
    // outputs (skk_finish_exchange_kernel: two launches per evaluation); SKK_SHARDED_FINISH=0 forces the
    // separate kernels
    const bool fused_sharded = sharded && !chain && pl->fused_ok && batch <= 4 && env_int("SKK_SHARDED_FINISH", 1) == 1;
    const bool in_row = fused || fused_sharded;  // the prior terms ride in the row kernel
    const bool pdl = env_int("SKK_NO_PDL", 0) == 0;  // the finish kernel starts under the row kernel's tail
    // (for large batches they are big enough to fill the GPU themselves: run them in line -- on a parallel branch they
    // take SMs from the chain kernel: c5 509 -> 1012 us, 128 chains 119 -> 138 us, profiles/r02/q_all.log)
    const bool fork = batch < kChainMinBatch;
    cudaStream_t ps = fork ? pl->side : st;
    if (priors && !in_row) {
        if (fork) {
            CUDA_TRY(cudaEventRecord(pl->ev_fork, st));
            CUDA_TRY(cudaStreamWaitEvent(pl->side, pl->ev_fork, 0));
        }
        PriorParams pp = pl->prior;
        pp.batch = batch;
        skk_prior_kernel<T><<<dim3((unsigned)pl->prior_ctas, (unsigned)batch), kPriorThreads, 0, ps>>>(
            pp, th, pl->prior_grad, pl->prior_lp, pl->link_val);
        ++launches;
        if (pp.n_parents > 0) {
            const int per_cta = pp.n_heavy > 0 ? 32 : 8;  // light parents: one warp each
            const unsigned blocks = (unsigned)(pp.n_heavy + (pp.n_parents - pp.n_heavy + per_cta - 1) / per_cta);
            skk_prior_link_kernel<<<dim3(blocks, (unsigned)batch), pp.n_heavy > 0 ? 1024 : 256, 0, ps>>>(
                pp, pl->prior_grad, pl->link_val);
            ++launches;
        }
        if (fork) CUDA_TRY(cudaEventRecord(pl->ev_join, pl->side));
    }

    if (chain) {
        // K2: lane = chain; one pass over the rows for the whole batch
        ChainParams<T> cp{};
        // (the chain kernel reads the columns in row order)
        const bool rows = pl->y_rows != nullptr;
        for (int c = 0; c < pl->n_x; ++c) cp.x[c] = static_cast<const T *>(rows ? pl->x_rows[c] : pl->x[c]);
        const void *yy = rows ? pl->y_rows : pl->y;
        cp.y = pl->y_u8 ? nullptr : static_cast<const T *>(yy);
        cp.y8 = pl->y_u8 ? static_cast<const uint8_t *>(yy) : nullptr;
        cp.eta_offset = static_cast<const T *>(rows ? pl->off_rows : pl->off);
        cp.n_x = pl->n_x;
        cp.n_rows = pl->n_rows;
        cp.n_tiles = pl->chain_tiles;
        cp.range_keys = pl->chain_keys;
        cp.seg_offsets = pl->seg_off;
        cp.n_primary = (int32_t)pl->n_primary;
        for (int i = 0; i < pl->n_terms; ++i) {
            const int t = pl->term_slot_in_level[i];
            if (pl->term_level[i] == SKK_LEVEL_GLOBAL) {
                cp.xcol_g[t] = pl->xcol_g[t];
                cp.theta_index_g[t] = pl->theta_index_dev[i];
            } else {
                cp.xcol_p[t] = pl->xcol_p[t];
                cp.theta_index_p[t] = pl->theta_index_dev[i];
            }
        }
        cp.theta = th;
        cp.fam_a = (T)((pl->family == SKK_LIK_NEGBINOMIAL_LOG || pl->family == SKK_LIK_GAMMA_LOG) ? pl->sigma_const : 0.0);
        cp.fam_b = (T)(pl->family == SKK_LIK_NEGBINOMIAL_LOG ? log(pl->sigma_const) : 0.0);
        cp.n_params = pl->n_params;
        cp.batch = batch;
        cp.bp = pl->chain_bp;
        cp.direct = pl->chain_totals + (size_t)(pl->ng + 1) * pl->chain_bp;
        cp.head = pl->chain_head;
        cp.tail = pl->chain_tail;
        cp.glob = pl->chain_glob;
        if (pl->np > 0)
            CUDA_TRY(cudaMemsetAsync(cp.direct, 0, (size_t)pl->np * pl->n_primary * pl->chain_bp * sizeof(double), st));
        const int gy = ((batch + 31) / 32 + kChainWarps - 1) / kChainWarps;
        void *args[] = {&cp};
        if (timed) CUDA_TRY(cudaEventRecord(pl->evk0, st));
        CUDA_TRY(cudaLaunchKernel(pl->chain_fn, dim3(pl->chain_ranges, gy), dim3(kChainWarps * 32), args, pl->chain_smem, st));
        if (timed) CUDA_TRY(cudaEventRecord(pl->evk1, st));
        ++launches;
        pl->stats.row_kernel_launches++;
        ChainReduceParams cr = pl->chain_reduce;
        cr.batch = batch;
        const int threads = std::max(128, std::min(256, (batch + 31) & ~31));  // (>= 4 warps share the ranges of a global term)
        skk_chain_slot_kernel<<<dim3((unsigned)std::max<int64_t>(pl->ng + 1, std::min<int64_t>(std::max<int64_t>(1, pl->n_primary), 2048)), 2),
                                256, 0, st>>>(cr);
        ++launches;
        if (priors && fork) CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));  // the final kernel adds the prior terms
        skk_chain_final_kernel<T><<<(unsigned)std::min<size_t>(std::max<size_t>(1, P), 2048), threads, 0, st>>>(cr, th, out_logp, out_grad);
        ++launches;
        CUDA_TRY(cudaGetLastError());
        *n_launch = launches;
        return SKK_OK;
    }

    bool joined = false, one_launch = false;
    int sharded_lp_stride = 0;
    for (int b0 = 0; b0 < batch;) {
        const int remaining = batch - b0;
        const bool wide = remaining >= 2 && pl->cfg4.fn != nullptr;
        const int bt = wide ? 4 : 1;
        const int b_valid = std::min(bt, remaining);
        const LaunchCfg &cfg = wide ? pl->cfg4 : pl->cfg1;
        // K1: the rows
        if (int rc = launch_rows<T>(pl, cfg, timed, th + (size_t)b0 * P, b_valid, bt, b0, in_row && priors)) return rc;
        ++launches;
        pl->stats.row_kernel_launches++;

        // K4: totals per slot, then per theta element
        ReduceParams rp = pl->reduce;
        rp.bt = bt;
        rp.direct_p = bt == 4 ? pl->direct_p : pl->direct_p1;
        rp.b0 = b0;
        rp.b_valid = b_valid;
        rp.grid = cfg.grid;
        rp.batch_total = batch;
        rp.trace = pl->trace ? pl->trace + (size_t)std::max(pl->cfg1.grid, pl->cfg4.grid) * pl->warps * kTraceWords : nullptr;
        const unsigned slot_blocks = (unsigned)(rp.blocks_p + rp.blocks_s + 1);
        if (fused) {
            if (priors) {
                rp.prior_ctas = cfg.grid * pl->warps;  // one logp partial per warp of the row kernel
                rp.n_lp_parts = rp.prior_ctas;
            }
            const unsigned blocks = slot_blocks + (unsigned)(rp.parent_blocks + rp.none_blocks + rp.multi_blocks);
            if (bt == 4)
                CUDA_TRY(launch_after_rows(skk_finish_kernel<T, 4>, blocks, st, pdl, rp, pl->peer, th, out_logp, out_grad));
            else
                CUDA_TRY(launch_after_rows(skk_finish_kernel<T, 1>, blocks, st, pdl, rp, pl->peer, th, out_logp, out_grad));
            ++launches;
            b0 += b_valid;
            continue;
        }
        if (fused_sharded) {
            const unsigned blocks = slot_blocks + (unsigned)(rp.parent_blocks + rp.none_blocks + rp.multi_blocks);
            rp.prior_ctas = cfg.grid * pl->warps;  // the prologue's warp partials of the prior logp
            rp.n_lp_parts = rp.prior_ctas;
            sharded_lp_stride = rp.prior_ctas;
            // every CTA of the one-launch exchange waits inside the kernel: all of them must be resident
            one_launch = pl->peer_ready && (int)blocks <= pl->exchange_ctas_max && env_int("SKK_SHARDED_FINISH_SPLIT", 0) == 0;
            if (one_launch) {
                if (bt == 4)
                    CUDA_TRY(launch_after_rows(skk_finish_exchange_kernel<T, 4>, blocks, st, pdl, rp, pl->peer, th, out_logp, out_grad));
                else
                    CUDA_TRY(launch_after_rows(skk_finish_exchange_kernel<T, 1>, blocks, st, pdl, rp, pl->peer, th, out_logp, out_grad));
            } else {
#define SKK_PUSH(BTV, MODEV) CUDA_TRY(launch_after_rows(skk_finish_push_kernel<T, BTV, MODEV>, blocks, st, pdl, rp, pl->peer))
                if (bt == 4) {
                    if (pl->peer_ready) SKK_PUSH(4, 2); else SKK_PUSH(4, 1);
                } else {
                    if (pl->peer_ready) SKK_PUSH(1, 2); else SKK_PUSH(1, 1);
                }
#undef SKK_PUSH
            }
            ++launches;
            b0 += b_valid;
            continue;
        }
        const unsigned theta_blocks = (unsigned)(rp.theta_blocks + (rp.n_multi + 7) / 8);
        if (bt == 4)
            skk_slot_total_kernel<4><<<slot_blocks, 256, 0, st>>>(rp);
        else
            skk_slot_total_kernel<1><<<slot_blocks, 256, 0, st>>>(rp);
        ++launches;

        // K4f: per theta element; single GPU: straight to the outputs (needs the priors)
        if (!sharded && priors && fork && !joined) {
            CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));
            joined = true;
        }
        const int mode = !sharded ? 0 : (pl->peer_ready ? 2 : 1);
#define SKK_THETA(BTV, MODEV) \
    skk_theta_kernel<T, BTV, MODEV><<<theta_blocks, 256, 0, st>>>(rp, pl->peer, th, out_logp, out_grad)
        if (bt == 4) {
            if (mode == 0) SKK_THETA(4, 0); else if (mode == 1) SKK_THETA(4, 1); else SKK_THETA(4, 2);
        } else {
            if (mode == 0) SKK_THETA(1, 0); else if (mode == 1) SKK_THETA(1, 1); else SKK_THETA(1, 2);
        }
#undef SKK_THETA
        ++launches;
        b0 += b_valid;
    }

    if (sharded && !one_launch) {
        // one all-reduce of [dlogp, logp partial, constant] per evaluation, then the final combination
        FinalParams fp{pl->lik_vec, pl->prior_grad, pl->prior_lp, fused_sharded ? sharded_lp_stride : pl->prior_ctas,
                       pl->n_params, pl->n_rows_total, pl->sigma_idx, pl->sigma_const, pl->family, pl->include_priors};
        dim3 grid((unsigned)std::max<size_t>(1, (P + 255) / 256), (unsigned)batch);
        if (pl->peer_ready) {
            if (priors && fork && !in_row) CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));
            skk_final_kernel<T, true><<<grid, 256, 0, st>>>(fp, pl->peer, th, out_logp, out_grad);
        } else {
            NCCL_TRY(g_nccl.AllReduce(pl->lik_vec, pl->lik_vec, (size_t)batch * (P + 2), /*ncclDouble*/ 8,
                                      /*ncclSum*/ 0, pl->comm, st));
            ++launches;
            if (priors && fork && !in_row) CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));
            skk_final_kernel<T, false><<<grid, 256, 0, st>>>(fp, pl->peer, th, out_logp, out_grad);
        }
        ++launches;
    } else if (priors && !in_row && fork && !joined) {
        CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));
    }
    CUDA_TRY(cudaGetLastError());
    *n_launch = launches;
    return SKK_OK;
}

template <typename T>
static int eval_impl(skk_plan *pl, const void *theta, int batch, void *logp_out, void *grad_out, unsigned flags) {
    const size_t P = (size_t)pl->n_params;
    const bool dev_io = flags & SKK_EVAL_DEVICE_IO;
    const bool timed = flags & SKK_EVAL_TIME_KERNELS;
    cudaStream_t st = pl->stream;
    CUDA_TRY(cudaSetDevice(pl->device));
    // a timed-out exchange leaves the ranks' sequence counters out of step: the plan refuses further work
    // (the status word is host-mapped, so enqueue-only evaluations notice it at the next call)
    if (pl->peer_status && *pl->peer_status) pl->peer_failed = true;
    if (pl->peer_failed)
        return fail(SKK_ERR_NCCL, "peer all-reduce timed out waiting for another rank; destroy the plans of all ranks "
                                  "and connect again");

    // (which kernels this evaluation takes: the same test as in enqueue_eval)
    pl->stats.kernel_variant = (pl->chain_ready && batch >= kChainMinBatch && !(pl->comm != nullptr || pl->peer_ready)) ? 2 : 1;

    const T *th = dev_io ? static_cast<const T *>(theta) : static_cast<const T *>(pl->theta_dev);
    T *out_grad = dev_io ? static_cast<T *>(grad_out) : static_cast<T *>(pl->out_dev);
    T *out_logp = dev_io ? static_cast<T *>(logp_out) : static_cast<T *>(pl->out_dev) + (size_t)batch * P;
    if (!dev_io) std::memcpy(pl->h_in, theta, (size_t)batch * P * sizeof(T));

    // Graph replay: the sequence (including the NCCL all-reduce of sharded runs) is fixed for
    // (batch, buffers); evaluations with event timing of the row kernel stay on the direct path.
    const bool graphable = pl->use_graphs && !timed;
    GraphEntry *hit = nullptr;
    if (graphable)
        for (GraphEntry &g : pl->graphs)
            if (g.batch == batch && (g.flags & 0xffu) == (flags & SKK_EVAL_DEVICE_IO) && g.theta == th && g.logp == out_logp &&
                g.grad == out_grad)
                hit = &g;

    int launches = 0;
    auto enqueue_all = [&]() -> int {
        if (!dev_io)
            CUDA_TRY(cudaMemcpyAsync(pl->theta_dev, pl->h_in, (size_t)batch * P * sizeof(T), cudaMemcpyHostToDevice, st));
        if (int rc = enqueue_eval<T>(pl, th, batch, out_logp, out_grad, timed, &launches)) return rc;
        if (!dev_io)  // gradient rows and logp are contiguous in out_dev for this batch size
            CUDA_TRY(cudaMemcpyAsync(pl->h_out, pl->out_dev, ((size_t)batch * P + batch) * sizeof(T),
                                     cudaMemcpyDeviceToHost, st));
        return SKK_OK;
    };

    CUDA_TRY(cudaEventRecord(pl->ev0, st));
    if (graphable && !hit && pl->graphs.size() < 16) {
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_all();
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != SKK_OK || ce != cudaSuccess) {
            // not capturable in this environment (e.g. an NCCL build without graph support): run directly
            if (graph) cudaGraphDestroy(graph);
            (void)cudaGetLastError();
            pl->use_graphs = false;
            if (int rc2 = enqueue_all()) return rc2;
            goto enqueued;
        }
        GraphEntry g;
        g.batch = batch;
        g.flags = flags & SKK_EVAL_DEVICE_IO;
        g.theta = th;
        g.logp = out_logp;
        g.grad = out_grad;
        ce = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail(SKK_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
        pl->graphs.push_back(g);
        hit = &pl->graphs.back();
        hit->flags |= (unsigned)launches << 8;  // remember the kernel count of this graph
    }
    if (hit) {
        CUDA_TRY(cudaGraphLaunch(hit->exec, st));
        launches = (int)(hit->flags >> 8);
    } else {
        if (int rc = enqueue_all()) return rc;
    }
enqueued:
    pl->stats.launches += launches;
    CUDA_TRY(cudaEventRecord(pl->ev1, st));

    if (dev_io && (flags & SKK_EVAL_NO_SYNC)) {
        pl->stats.evals++;
        return SKK_OK;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (pl->peer_status && *pl->peer_status) {
        pl->peer_failed = true;
        return fail(SKK_ERR_NCCL, "peer all-reduce timed out waiting for another rank");
    }
    if (!dev_io) {
        std::memcpy(grad_out, pl->h_out, (size_t)batch * P * sizeof(T));
        std::memcpy(logp_out, static_cast<T *>(pl->h_out) + (size_t)batch * P, (size_t)batch * sizeof(T));
    }
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, pl->ev0, pl->ev1));
    pl->stats.last_eval_ms = ms;
    if (timed) {
        CUDA_TRY(cudaEventElapsedTime(&ms, pl->evk0, pl->evk1));
        pl->stats.last_row_kernel_ms = ms;
    }
    pl->stats.evals++;
    return SKK_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *skk_last_error(void) { return g_error.c_str(); }

int skk_abi_version(void) { return SKK_ABI_VERSION; }

static void free_plan(skk_plan *pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    if (pl->stream) cudaStreamSynchronize(pl->stream);
    // graphs that captured the all-reduce hold the communicator: release them before destroying it
    for (GraphEntry &g : pl->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    pl->graphs.clear();
    if (pl->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(pl->comm);
    for (int r = 0; r < kMaxPeers; ++r)
        if (pl->peer_opened[r]) cudaIpcCloseMemHandle(pl->peer_opened[r]);
    if (pl->peer_status) cudaFreeHost(pl->peer_status);
    for (void *p : pl->allocs) cudaFree(p);
    if (pl->h_in) cudaFreeHost(pl->h_in);
    if (pl->h_out) cudaFreeHost(pl->h_out);
    for (GraphEntry &g : pl->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (cudaEvent_t e : {pl->ev0, pl->ev1, pl->evk0, pl->evk1, pl->ev_fork, pl->ev_join})
        if (e) cudaEventDestroy(e);
    if (pl->side) cudaStreamDestroy(pl->side);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    delete pl;
}

int skk_plan_create(const skk_plan_desc *d, skk_plan **out) {
    if (!d || !out) return fail(SKK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (d->abi_version != SKK_ABI_VERSION)
        return fail(SKK_ERR_INVALID, "descriptor ABI %d, library ABI %d", d->abi_version, SKK_ABI_VERSION);
    if (d->dtype != SKK_F32 && d->dtype != SKK_F64) return fail(SKK_ERR_INVALID, "bad dtype %d", d->dtype);
    if (d->family < SKK_LIK_NORMAL || d->family > SKK_LIK_GAMMA_LOG) return fail(SKK_ERR_UNSUPPORTED, "bad family %d", d->family);
    if (d->family == SKK_LIK_STUDENTT_LOGSIGMA && !(d->sigma_const > 0.0)) return fail(SKK_ERR_INVALID, "StudentT nu must be positive");
    if (d->family == SKK_LIK_NEGBINOMIAL_LOG && !(d->sigma_const > 0.0)) return fail(SKK_ERR_INVALID, "NegativeBinomial alpha must be positive");
    if (d->family == SKK_LIK_GAMMA_LOG && !(d->sigma_const > 0.0)) return fail(SKK_ERR_INVALID, "Gamma alpha must be positive");
    if (d->n_rows < 0 || d->n_rows >= (int64_t)1 << 40) return fail(SKK_ERR_INVALID, "bad n_rows");
    if ((d->n_rows + 32 * kRowsPerLane - 1) / (32 * kRowsPerLane) >= INT32_MAX) return fail(SKK_ERR_UNSUPPORTED, "too many rows");
    if (d->n_params < 0 || d->n_params > INT32_MAX - 2) return fail(SKK_ERR_INVALID, "bad n_params");
    if (d->max_batch < 1) return fail(SKK_ERR_INVALID, "max_batch must be >= 1");
    if (d->n_xcols < 0 || d->n_xcols > kMaxXCols) return fail(SKK_ERR_UNSUPPORTED, "at most %d data columns", kMaxXCols);
    if (d->n_terms < 0 || d->n_terms > kMaxTerms) return fail(SKK_ERR_UNSUPPORTED, "at most %d terms", kMaxTerms);
    if (d->n_primary >= INT32_MAX || d->n_secondary >= INT32_MAX) return fail(SKK_ERR_UNSUPPORTED, "too many groups");
    if (d->y_kind != SKK_Y_FLOAT && d->y_kind != SKK_Y_U8) return fail(SKK_ERR_INVALID, "bad y_kind");
    if (d->n_rows > 0 && !d->y) return fail(SKK_ERR_INVALID, "y is null");

    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0)
        return fail(SKK_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    if (d->device < 0 || d->device >= n_dev) return fail(SKK_ERR_INVALID, "device %d of %d", d->device, n_dev);

    skk_plan *pl = new skk_plan();
    pl->dtype = d->dtype;
    pl->family = d->family;
    pl->device = d->device;
    pl->n_rows = d->n_rows;
    pl->n_params = d->n_params;
    pl->n_rows_total = d->n_rows_total > 0 ? d->n_rows_total : d->n_rows;
    pl->max_batch = d->max_batch;
    pl->n_x = d->n_xcols;
    pl->y_u8 = d->y_kind == SKK_Y_U8;
    pl->has_off = d->eta_offset != nullptr;
    pl->sec_code_bytes = d->sec_code_bytes == 2 ? 2 : 4;
    pl->n_primary = d->n_primary;
    pl->n_secondary = d->n_secondary;
    pl->sigma_const = d->sigma_const;
    pl->sigma_idx = d->sigma_theta_index;
    pl->logp_const = d->logp_const;
    pl->include_priors = d->include_priors;
    for (int t = 0; t < kMaxTermsPerLevel; ++t) pl->xcol_g[t] = pl->xcol_p[t] = pl->xcol_s[t] = -1;

    int rc = SKK_OK;
    for (int i = 0; i < d->n_terms && rc == SKK_OK; ++i) {
        const skk_term_desc &td = d->terms[i];
        if (td.level < SKK_LEVEL_GLOBAL || td.level > SKK_LEVEL_SECONDARY) rc = fail(SKK_ERR_INVALID, "term %d: bad level", i);
        else if ((td.xcol < -1 && !(td.xcol == SKK_XCOL_LOG_SIGMA && (d->family == SKK_LIK_NORMAL_LOGSIGMA || d->family == SKK_LIK_STUDENTT_LOGSIGMA))) || td.xcol >= d->n_xcols)
            rc = fail(SKK_ERR_INVALID, "term %d: bad xcol", i);
        else if (!td.theta_index) rc = fail(SKK_ERR_INVALID, "term %d: null theta_index", i);
        else if (td.level == SKK_LEVEL_GLOBAL) { if (td.n_slots != 1) rc = fail(SKK_ERR_INVALID, "term %d: global terms have one slot", i); ++pl->ng; }
        else if (td.level == SKK_LEVEL_PRIMARY) { if (td.n_slots != d->n_primary) rc = fail(SKK_ERR_INVALID, "term %d: n_slots != n_primary", i); ++pl->np; }
        else { if (td.n_slots != d->n_secondary) rc = fail(SKK_ERR_INVALID, "term %d: n_slots != n_secondary", i); ++pl->ns; }
    }
    if (rc == SKK_OK && (pl->ng > kMaxTermsPerLevel || pl->np > kMaxTermsPerLevel || pl->ns > 1))
        rc = fail(SKK_ERR_UNSUPPORTED, "term counts %d/%d/%d exceed the compiled kernels (2/2/1)", pl->ng, pl->np, pl->ns);
    if (rc == SKK_OK && pl->np > 0 && (d->n_primary < 1 || !d->seg_offsets)) rc = fail(SKK_ERR_INVALID, "primary terms without segments");
    if (rc == SKK_OK && pl->ns > 0 && (d->n_secondary < 1 || !d->sec_codes)) rc = fail(SKK_ERR_INVALID, "secondary terms without codes");
    if (rc == SKK_OK && d->sigma_theta_index >= d->n_params) rc = fail(SKK_ERR_INVALID, "sigma_theta_index outside theta");
    if (rc != SKK_OK) {
        delete pl;
        return rc;
    }

    auto guard = [&](int code) {
        if (code != SKK_OK) free_plan(pl);
        return code;
    };
    ce = cudaSetDevice(pl->device);
    if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(ce)));
    ce = cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)));
    ce = cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking);
    if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)));
    for (cudaEvent_t *e : {&pl->ev0, &pl->ev1, &pl->evk0, &pl->evk1}) {
        ce = cudaEventCreate(e);
        if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
    }
    for (cudaEvent_t *e : {&pl->ev_fork, &pl->ev_join}) {
        ce = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
        if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
    }
    rc = d->dtype == SKK_F64 ? create_impl<double>(d, pl) : create_impl<float>(d, pl);
    if (rc != SKK_OK) return guard(rc);
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "plan upload: %s", cudaGetErrorString(ce)));
    if (pl->first_arena && env_int("SKK_L2_PERSIST", 1)) {
        // best effort: a device without the feature (or with the set-aside taken) runs without the window
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, pl->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, 32u << 20);
            size_t have = 0;
            cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
            if (have < want) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = pl->first_arena;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(kFirstArenaBytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            pl->l2_window = cudaStreamSetAttribute(pl->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess &&
                            cudaStreamSetAttribute(pl->side, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
        }
        (void)cudaGetLastError();
    }

    const size_t e = elt_size(pl->dtype);
    pl->stats.stored_bytes_per_row = (uint64_t)(pl->n_x * e + (pl->y_u8 ? 1 : e) + (pl->has_off ? e : 0) +
                                                (pl->ns > 0 ? pl->sec_code_bytes : 0));
    pl->stats.device_bytes = pl->device_bytes;
    pl->stats.grid = pl->cfg1.grid;
    pl->stats.block = pl->warps * 32;
    pl->stats.smem_bytes = pl->cfg1.smem;
    pl->stats.n_tiles = pl->n_tiles;
    pl->stats.rows_per_tile = pl->rows_per_tile;
    pl->stats.kernel_variant = 1;
    *out = pl;
    return SKK_OK;
}

int skk_eval(skk_plan *pl, const void *theta, int32_t batch, void *logp_out, void *grad_out, uint32_t flags) {
    if (!pl) return fail(SKK_ERR_INVALID, "null plan");
    if (batch < 1 || batch > pl->max_batch) return fail(SKK_ERR_INVALID, "batch %d outside [1, %d]", batch, pl->max_batch);
    if (!theta || !logp_out || !grad_out) return fail(SKK_ERR_INVALID, "null buffer");
    return pl->dtype == SKK_F64 ? eval_impl<double>(pl, theta, batch, logp_out, grad_out, flags)
                                : eval_impl<float>(pl, theta, batch, logp_out, grad_out, flags);
}

int skk_plan_destroy(skk_plan *pl) {
    free_plan(pl);
    return SKK_OK;
}

int skk_comm_unique_id(void *out128) {
    if (!out128) return fail(SKK_ERR_INVALID, "null argument");
    if (int rc = load_nccl()) return rc;
    NCCL_TRY(g_nccl.GetUniqueId(out128));
    return SKK_OK;
}

int skk_comm_init(skk_plan *pl, const void *unique_id128, int32_t rank, int32_t nranks) {
    if (!pl || !unique_id128) return fail(SKK_ERR_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(SKK_ERR_INVALID, "rank %d of %d", rank, nranks);
    if (pl->comm) return fail(SKK_ERR_INVALID, "communicator already initialised");
    if (int rc = load_nccl()) return rc;
    CUDA_TRY(cudaSetDevice(pl->device));
    NcclId id;
    std::memcpy(id.bytes, unique_id128, 128);
    NCCL_TRY(g_nccl.CommInitRank(&pl->comm, nranks, id, rank));
    for (GraphEntry &g : pl->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    pl->graphs.clear();
    pl->nranks = nranks;
    pl->rank = rank;
    return SKK_OK;
}

int skk_comm_destroy(skk_plan *pl) {
    if (!pl) return fail(SKK_ERR_INVALID, "null plan");
    if (pl->comm) {
        CUDA_TRY(cudaStreamSynchronize(pl->stream));
        for (GraphEntry &g : pl->graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        pl->graphs.clear();
        NCCL_TRY(g_nccl.CommDestroy(pl->comm));
        pl->comm = nullptr;
    }
    return SKK_OK;
}

static int64_t peer_stride(const skk_plan *pl) {
    return (int64_t)pl->max_batch * (pl->n_params + 2);  // 16-byte lines per slot
}

int skk_peer_alloc(skk_plan *pl, int32_t nranks, void *handle_out64) {
    if (!pl || !handle_out64) return fail(SKK_ERR_INVALID, "null argument");
    if (nranks < 1 || nranks > kMaxPeers) return fail(SKK_ERR_UNSUPPORTED, "peer all-reduce handles up to %d ranks", kMaxPeers);
    if (pl->peer_local) return fail(SKK_ERR_INVALID, "peer buffer already allocated");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CUDA_TRY(cudaSetDevice(pl->device));
    // [2 parities][nranks][stride] lines of 16 bytes | fed8 [n_params + 2] words of 8 bytes
    const size_t bytes = (size_t)2 * nranks * peer_stride(pl) * sizeof(uint4) + (size_t)(pl->n_params + 2) * 8;
    CUDA_TRY(cudaMalloc(&pl->peer_local, bytes));   // (its own allocation: the IPC handle exports the whole allocation)
    pl->allocs.push_back(pl->peer_local);
    pl->device_bytes += bytes;
    CUDA_TRY(cudaMemset(pl->peer_local, 0, bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, pl->peer_local));
    std::memcpy(handle_out64, &h, 64);
    pl->peer_nranks = nranks;
    return SKK_OK;
}

int skk_peer_init(skk_plan *pl, const void *handles, int32_t rank, int32_t nranks) {
    if (!pl || !handles) return fail(SKK_ERR_INVALID, "null argument");
    if (!pl->peer_local || nranks != pl->peer_nranks) return fail(SKK_ERR_INVALID, "call skk_peer_alloc with the same nranks first");
    if (rank < 0 || rank >= nranks) return fail(SKK_ERR_INVALID, "rank %d of %d", rank, nranks);
    if (pl->peer_ready) return fail(SKK_ERR_INVALID, "peer all-reduce already initialised");
    CUDA_TRY(cudaSetDevice(pl->device));
    const int64_t stride = peer_stride(pl);
    const int64_t fed8_line_offset = (int64_t)2 * nranks * stride;
    PeerParams pp{};
    pp.stride = stride;
    pp.rank = rank;
    pp.nranks = nranks;
    for (int r = 0; r < nranks; ++r) {
        void *base = pl->peer_local;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const char *>(handles) + (size_t)r * 64, 64);
            CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            pl->peer_opened[r] = base;
        }
        pp.peer_lines[r] = static_cast<uint4 *>(base);
    }
    pp.fed8 = reinterpret_cast<const unsigned long long *>(static_cast<uint4 *>(pl->peer_local) + fed8_line_offset);
    unsigned long long *words = nullptr;
    if (int rc = dev_alloc(pl, &words, 2, true)) return rc;
    pp.seq = words;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&pl->peer_status), sizeof(int), cudaHostAllocMapped));
    *pl->peer_status = 0;
    CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&pp.status), pl->peer_status, 0));
    pl->peer = pp;
    {
        // this rank's byte of fed8 on every peer.  The caller synchronises the ranks (a barrier) between
        // skk_peer_init and the first evaluation, as sakkara_b200/distributed.py does.
        uint8_t *fed_dev = nullptr;
        if (int rc = dev_upload(pl, &fed_dev, pl->fed.data(), pl->fed.size())) return rc;
        const int64_t n = (int64_t)pl->fed.size();
        skk_peer_fed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pl->stream>>>(pp, fed_dev, n, fed8_line_offset);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(pl->stream));
    }
    {
        // CTAs of the one-launch exchange kernel that fit on the device at once (they wait for each other)
        int per_sm1 = 0, per_sm4 = 0;
        if (pl->dtype == SKK_F64) {
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, skk_finish_exchange_kernel<double, 1>, 256, 0));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm4, skk_finish_exchange_kernel<double, 4>, 256, 0));
        } else {
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, skk_finish_exchange_kernel<float, 1>, 256, 0));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm4, skk_finish_exchange_kernel<float, 4>, 256, 0));
        }
        pl->exchange_ctas_max = pl->n_sms * std::min(per_sm1, per_sm4);
    }
    for (GraphEntry &g : pl->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    pl->graphs.clear();
    pl->peer_ready = true;
    pl->nranks = nranks;
    pl->rank = rank;
    return SKK_OK;
}

int skk_stats(skk_plan *pl, skk_stats_t *out) {
    if (!pl || !out) return fail(SKK_ERR_INVALID, "null argument");
    *out = pl->stats;
    return SKK_OK;
}

void *skk_stream(skk_plan *pl) { return pl ? pl->stream : nullptr; }

int skk_leapfrog(skk_plan *pl, void *q, void *r, const void *grad, const void *eps, const void *inv_mass, int32_t chains,
                 double a, double b) {
    if (!pl) return fail(SKK_ERR_INVALID, "null plan");
    if (!q || !r || !grad || !eps || !inv_mass) return fail(SKK_ERR_INVALID, "null buffer");
    if (chains < 1 || chains > 65535) return fail(SKK_ERR_INVALID, "chains %d outside [1, 65535]", chains);
    CUDA_TRY(cudaSetDevice(pl->device));
    const int n = (int)pl->n_params;
    const dim3 grid((unsigned)std::min((n + 255) / 256, 64), (unsigned)chains);
    if (pl->dtype == SKK_F64)
        skk_leapfrog_kernel<double><<<grid, 256, 0, pl->stream>>>(
            static_cast<double *>(q), static_cast<double *>(r), static_cast<const double *>(grad),
            static_cast<const double *>(eps), static_cast<const double *>(inv_mass), n, a, b);
    else
        skk_leapfrog_kernel<float><<<grid, 256, 0, pl->stream>>>(
            static_cast<float *>(q), static_cast<float *>(r), static_cast<const float *>(grad),
            static_cast<const float *>(eps), static_cast<const float *>(inv_mass), n, (float)a, (float)b);
    CUDA_TRY(cudaGetLastError());
    return SKK_OK;
}

int skk_trace_enable(skk_plan *pl, int32_t on) {
    if (!pl) return fail(SKK_ERR_INVALID, "null plan");
    CUDA_TRY(cudaSetDevice(pl->device));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    for (GraphEntry &g : pl->graphs)   // the pointer is a kernel argument: recorded graphs are stale
        if (g.exec) cudaGraphExecDestroy(g.exec);
    pl->graphs.clear();
    if (on && !pl->trace) {
        pl->trace_words = ((size_t)std::max(pl->cfg1.grid, pl->cfg4.grid) * pl->warps + kFinishTraceCtas) * kTraceWords;
        if (int rc = dev_alloc(pl, &pl->trace, pl->trace_words, true)) return rc;
    }
    if (!on) pl->trace = nullptr;  // (the buffer stays with the plan's allocations)
    return SKK_OK;
}

int64_t skk_trace_read(skk_plan *pl, uint64_t *out, int64_t capacity) {
    if (!pl || !out) return fail(SKK_ERR_INVALID, "null argument");
    if (!pl->trace) return fail(SKK_ERR_INVALID, "tracing is not enabled");
    CUDA_TRY(cudaSetDevice(pl->device));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    const int64_t n = std::min<int64_t>(capacity, (int64_t)pl->trace_words);
    CUDA_TRY(cudaMemcpy(out, pl->trace, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return n;
}

}  // extern "C"
