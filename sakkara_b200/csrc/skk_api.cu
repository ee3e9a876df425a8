// Host side of libskk_b200.so: the C ABI of include/skk_b200.h.
//   plan creation : validate the descriptor, upload the row columns (padded to whole tiles), derive
//                   the tile keys, the reduce CSR and the prior link CSR, pick the kernel
//                   specialisation and its launch geometry, allocate every work buffer once.
//   evaluation    : H2D theta -> K3 priors -> per batch tile {K0 prepare, K1 rows, K4a reduce}
//                   -> [NCCL all-reduce when row-sharded] -> K4b final -> D2H.  One stream, no
//                   allocation, no host synchronisation before the final copy.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "skk_aux_kernels.cuh"
#include "skk_row_inst.cuh"

namespace skk {
const void *row_kernel_f32_normal_bt1(int, int, int);
const void *row_kernel_f32_normal_bt4(int, int, int);
const void *row_kernel_f32_bernoulli_bt1(int, int, int);
const void *row_kernel_f32_bernoulli_bt4(int, int, int);
const void *row_kernel_f32_poisson_bt1(int, int, int);
const void *row_kernel_f32_poisson_bt4(int, int, int);
const void *row_kernel_f64_normal_bt1(int, int, int);
const void *row_kernel_f64_normal_bt4(int, int, int);
const void *row_kernel_f64_bernoulli_bt1(int, int, int);
const void *row_kernel_f64_bernoulli_bt4(int, int, int);
const void *row_kernel_f64_poisson_bt1(int, int, int);
const void *row_kernel_f64_poisson_bt4(int, int, int);
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

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    std::vector<void *> allocs;
    size_t device_bytes = 0;

    void *x[kMaxXCols] = {nullptr};
    void *y = nullptr, *off = nullptr, *codes = nullptr;
    int2 *tile_keys = nullptr;
    int64_t *seg_off = nullptr;
    PrepParams prep{};

    void *theta_dev = nullptr, *val_g = nullptr, *val_p = nullptr, *val_s = nullptr;
    double *direct_p = nullptr, *tile_head = nullptr, *tile_tail = nullptr, *cta_glob = nullptr, *cta_sec = nullptr;
    double *lik_vec = nullptr, *prior_grad = nullptr, *prior_lp = nullptr, *to_mu = nullptr, *to_sigma = nullptr;
    double *glob_tot = nullptr, *sec_tot = nullptr;
    void *out_dev = nullptr;  // [B*P grad | B logp]

    PriorParams prior{};
    ReduceParams reduce{};

    void *h_in = nullptr, *h_out = nullptr;  // pinned staging

    void *comm = nullptr;
    int nranks = 1, rank = 0;

    skk_stats_t stats{};
};

static const int64_t kHeavyLinks = 512;  // hyper-parameters with more children get a whole CTA

static size_t elt_size(int dtype) { return dtype == SKK_F64 ? 8 : 4; }

template <typename U>
static int dev_alloc(skk_plan *pl, U **out, size_t count, bool zero = false) {
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(U), 16);
    CUDA_TRY(cudaMalloc(&p, bytes));
    pl->allocs.push_back(p);
    pl->device_bytes += bytes;
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

static const void *lookup_kernel(int dtype, int family, int bt, int ng, int np, int ns) {
    typedef const void *(*Lookup)(int, int, int);
    static const Lookup table[2][3][2] = {
        {{row_kernel_f32_normal_bt1, row_kernel_f32_normal_bt4},
         {row_kernel_f32_bernoulli_bt1, row_kernel_f32_bernoulli_bt4},
         {row_kernel_f32_poisson_bt1, row_kernel_f32_poisson_bt4}},
        {{row_kernel_f64_normal_bt1, row_kernel_f64_normal_bt4},
         {row_kernel_f64_bernoulli_bt1, row_kernel_f64_bernoulli_bt4},
         {row_kernel_f64_poisson_bt1, row_kernel_f64_poisson_bt4}}};
    if (family < 0 || family > 2 || (bt != 1 && bt != 4)) return nullptr;
    return table[dtype == SKK_F64 ? 1 : 0][family][bt == 4 ? 1 : 0](ng, np, ns);
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
    cfg->fn = lookup_kernel(pl->dtype, pl->family, bt, pl->ng, pl->np, pl->ns);
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

template <typename T>
static int create_impl(const skk_plan_desc *d, skk_plan *pl) {
    const int WT = 32 * kRowsPerLane;
    pl->rows_per_tile = WT;
    pl->n_tiles = (int)((d->n_rows + WT - 1) / WT);
    if (pl->n_tiles < 1) pl->n_tiles = 1;
    const size_t padded = (size_t)pl->n_tiles * WT;
    const size_t n = (size_t)d->n_rows;

    // ---- row columns, padded to whole tiles ------------------------------------------------------
    for (int c = 0; c < d->n_xcols; ++c) {
        int rc = upload_bytes(pl, &pl->x[c], d->xcols[c], n * sizeof(T), padded * sizeof(T));
        if (rc) return rc;
    }
    {
        const size_t ye = pl->y_u8 ? 1 : sizeof(T);
        int rc = upload_bytes(pl, &pl->y, d->y, n * ye, padded * ye);
        if (rc) return rc;
    }
    if (d->eta_offset) {
        int rc = upload_bytes(pl, &pl->off, d->eta_offset, n * sizeof(T), padded * sizeof(T));
        if (rc) return rc;
    }
    if (pl->ns > 0) {
        int rc = upload_bytes(pl, &pl->codes, d->sec_codes, n * pl->sec_code_bytes, padded * pl->sec_code_bytes);
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
    {
        int rc = dev_upload(pl, &pl->tile_keys, keys.data(), keys.size());
        if (rc) return rc;
    }

    // ---- terms: value-table layout and the reduce CSR ------------------------------------------------
    struct Entry {
        int64_t theta;
        int8_t level, t;
        int32_t slot;
    };
    std::vector<Entry> entries;
    int cnt[3] = {0, 0, 0};
    pl->prep.n_terms = d->n_terms;
    pl->prep.n_params = d->n_params;
    for (int i = 0; i < d->n_terms; ++i) {
        const skk_term_desc &td = d->terms[i];
        const int t = cnt[td.level]++;
        int32_t *ti = nullptr;
        int rc = dev_upload(pl, &ti, td.theta_index, (size_t)td.n_slots);
        if (rc) return rc;
        const int64_t slots = td.level == SKK_LEVEL_GLOBAL ? 1 : td.level == SKK_LEVEL_PRIMARY ? d->n_primary : d->n_secondary;
        pl->prep.term[i] = PrepTerm{ti, td.n_slots, (int64_t)t * slots, td.level};
        (td.level == SKK_LEVEL_GLOBAL ? pl->xcol_g : td.level == SKK_LEVEL_PRIMARY ? pl->xcol_p : pl->xcol_s)[t] = td.xcol;
        for (int64_t s = 0; s < td.n_slots; ++s) {
            if (td.theta_index[s] < 0 || td.theta_index[s] >= d->n_params)
                return fail(SKK_ERR_INVALID, "term %d: theta_index out of range", i);
            entries.push_back(Entry{td.theta_index[s], (int8_t)td.level, (int8_t)t, (int32_t)s});
        }
    }
    std::stable_sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.theta < b.theta; });
    std::vector<int64_t> csr_ptr(d->n_params + 2, 0);
    std::vector<int8_t> e_level(entries.size()), e_t(entries.size());
    std::vector<int32_t> e_slot(entries.size());
    for (size_t k = 0; k < entries.size(); ++k) {
        csr_ptr[entries[k].theta + 1]++;
        e_level[k] = entries[k].level;
        e_t[k] = entries[k].t;
        e_slot[k] = entries[k].slot;
    }
    for (int64_t j = 0; j < d->n_params; ++j) csr_ptr[j + 1] += csr_ptr[j];
    csr_ptr[d->n_params + 1] = csr_ptr[d->n_params];
    int64_t *csr_dev = nullptr;
    int8_t *lvl_dev = nullptr, *t_dev = nullptr;
    int32_t *slot_dev = nullptr;
    if (int rc = dev_upload(pl, &csr_dev, csr_ptr.data(), csr_ptr.size())) return rc;
    if (int rc = dev_upload(pl, &lvl_dev, e_level.data(), e_level.size())) return rc;
    if (int rc = dev_upload(pl, &t_dev, e_t.data(), e_t.size())) return rc;
    if (int rc = dev_upload(pl, &slot_dev, e_slot.data(), e_slot.size())) return rc;

    // ---- priors: block table and hyper-parameter link CSR -----------------------------------------------
    std::vector<DevBlock> blocks(d->n_blocks);
    std::vector<int32_t> block_of(d->n_params, -1);
    struct Link {
        int64_t parent;
        int32_t child;
        int8_t kind;
    };
    std::vector<Link> links;
    for (int k = 0; k < d->n_blocks; ++k) {
        const skk_block_desc &bd = d->blocks[k];
        if (bd.offset < 0 || bd.size < 0 || bd.offset + bd.size > d->n_params)
            return fail(SKK_ERR_INVALID, "block %d outside theta", k);
        if (bd.family != SKK_PRIOR_NORMAL && bd.family != SKK_PRIOR_HALFNORMAL && bd.family != SKK_PRIOR_EXPONENTIAL)
            return fail(SKK_ERR_UNSUPPORTED, "block %d: unknown prior family %d", k, bd.family);
        DevBlock &db = blocks[k];
        db.family = bd.family;
        db.offset = bd.offset;
        db.size = bd.size;
        db.a_const_len = bd.a_const_len;
        db.b_const_len = bd.b_const_len;
        db.a_const = db.b_const = nullptr;
        db.a_index = db.b_index = nullptr;
        if (bd.a_index) {
            if (bd.family != SKK_PRIOR_NORMAL) return fail(SKK_ERR_UNSUPPORTED, "block %d: only Normal priors take random parameters", k);
            int32_t *p = nullptr;
            if (int rc = dev_upload(pl, &p, bd.a_index, (size_t)bd.size)) return rc;
            db.a_index = p;
        } else {
            if (bd.a_const_len != 1 && bd.a_const_len != bd.size) return fail(SKK_ERR_INVALID, "block %d: bad a_const_len", k);
            double *p = nullptr;
            if (int rc = dev_upload(pl, &p, bd.a_const, (size_t)bd.a_const_len)) return rc;
            db.a_const = p;
        }
        if (bd.family == SKK_PRIOR_NORMAL) {
            if (bd.b_index) {
                int32_t *p = nullptr;
                if (int rc = dev_upload(pl, &p, bd.b_index, (size_t)bd.size)) return rc;
                db.b_index = p;
            } else {
                if (bd.b_const_len != 1 && bd.b_const_len != bd.size) return fail(SKK_ERR_INVALID, "block %d: bad b_const_len", k);
                double *p = nullptr;
                if (int rc = dev_upload(pl, &p, bd.b_const, (size_t)bd.b_const_len)) return rc;
                db.b_const = p;
            }
        }
        for (int64_t g = 0; g < bd.size; ++g) {
            if (block_of[bd.offset + g] != -1) return fail(SKK_ERR_INVALID, "blocks overlap at theta[%lld]", (long long)(bd.offset + g));
            block_of[bd.offset + g] = k;
            if (bd.a_index) links.push_back(Link{bd.a_index[g], (int32_t)(bd.offset + g), 0});
            if (bd.family == SKK_PRIOR_NORMAL && bd.b_index) links.push_back(Link{bd.b_index[g], (int32_t)(bd.offset + g), 1});
        }
    }
    for (int64_t j = 0; j < d->n_params; ++j)
        if (block_of[j] < 0) return fail(SKK_ERR_INVALID, "theta[%lld] belongs to no block", (long long)j);
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
    std::vector<int32_t> parent_el(parents.size()), link_child(links.size());
    std::vector<int64_t> link_ptr(parents.size() + 1, 0);
    std::vector<int8_t> link_kind(links.size());
    {
        size_t w = 0;
        for (size_t q = 0; q < parents.size(); ++q) {
            if (parents[q].hi - parents[q].lo > kHeavyLinks) ++n_heavy;
            parent_el[q] = parents[q].element;
            link_ptr[q] = (int64_t)w;
            for (int64_t e = parents[q].lo; e < parents[q].hi; ++e, ++w) {
                link_child[w] = links[e].child;
                link_kind[w] = links[e].kind;
            }
        }
        link_ptr[parents.size()] = (int64_t)w;
    }
    DevBlock *blocks_dev = nullptr;
    int32_t *block_of_dev = nullptr, *link_child_dev = nullptr, *parent_dev = nullptr;
    int64_t *link_ptr_dev = nullptr;
    int8_t *link_kind_dev = nullptr;
    if (int rc = dev_upload(pl, &blocks_dev, blocks.data(), blocks.size())) return rc;
    if (int rc = dev_upload(pl, &block_of_dev, block_of.data(), block_of.size())) return rc;
    if (int rc = dev_upload(pl, &parent_dev, parent_el.data(), parent_el.size())) return rc;
    if (int rc = dev_upload(pl, &link_ptr_dev, link_ptr.data(), link_ptr.size())) return rc;
    if (int rc = dev_upload(pl, &link_child_dev, link_child.data(), link_child.size())) return rc;
    if (int rc = dev_upload(pl, &link_kind_dev, link_kind.data(), link_kind.size())) return rc;
    pl->prior = PriorParams{blocks_dev, block_of_dev, parent_dev, link_ptr_dev, link_child_dev, link_kind_dev,
                            (int32_t)parents.size(), n_heavy, d->n_params, 0};

    // ---- launch geometry -------------------------------------------------------------------------------------
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, pl->device));
    pl->n_sms = prop.multiProcessorCount;
    const int smem_limit = (int)prop.sharedMemPerBlockOptin;
    pl->warps = std::max(1, std::min(8, env_int("SKK_WARPS", 8)));  // kernels are compiled for <= 256 threads
    pl->stages = std::max(2, std::min(8, env_int("SKK_STAGES", 3)));
    const int bt_max = d->max_batch > 1 ? 4 : 1;
    while (cta_smem<T>(pl, pl->warps, pl->stages, bt_max) > smem_limit) {
        if (pl->stages > 2) --pl->stages;
        else if (pl->warps > 1) pl->warps /= 2;
        else
            return fail(SKK_ERR_UNSUPPORTED, "secondary level with %lld members does not fit in shared memory",
                        (long long)pl->n_secondary);
    }
    // two resident CTAs per SM hide more latency than a third pipeline stage
    if (!getenv("SKK_STAGES") && pl->stages > 2) {
        const int per_sm = (int)prop.sharedMemPerMultiprocessor;
        if (2 * (cta_smem<T>(pl, pl->warps, pl->stages, 1) + 1024) > per_sm &&
            2 * (cta_smem<T>(pl, pl->warps, 2, 1) + 1024) <= per_sm)
            pl->stages = 2;
    }
    if (int rc = configure_launch<T>(pl, 1, &pl->cfg1)) return rc;
    if (bt_max == 4)
        if (int rc = configure_launch<T>(pl, 4, &pl->cfg4)) return rc;
    const int grid_max = std::max(pl->cfg1.grid, pl->cfg4.grid);

    // ---- work buffers --------------------------------------------------------------------------------------------
    const size_t B = (size_t)pl->max_batch, P = (size_t)d->n_params, bt = (size_t)bt_max;
    T *tmp = nullptr;
    if (int rc = dev_alloc(pl, &tmp, B * P)) return rc;
    pl->theta_dev = tmp;
    if (int rc = dev_alloc(pl, &tmp, (size_t)kMaxTermsPerLevel * bt, true)) return rc;
    pl->val_g = tmp;
    if (int rc = dev_alloc(pl, &tmp, std::max<size_t>(1, pl->np) * std::max<int64_t>(1, d->n_primary) * bt, true)) return rc;
    pl->val_p = tmp;
    if (int rc = dev_alloc(pl, &tmp, std::max<size_t>(1, pl->ns) * std::max<int64_t>(1, d->n_secondary) * bt, true)) return rc;
    pl->val_s = tmp;
    if (int rc = dev_alloc(pl, &pl->direct_p, std::max<size_t>(1, pl->np) * std::max<int64_t>(1, d->n_primary) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->tile_head, (size_t)pl->n_tiles * std::max(1, pl->np) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->tile_tail, (size_t)pl->n_tiles * std::max(1, pl->np) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->cta_glob, (size_t)grid_max * (pl->ng + 1) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->cta_sec, (size_t)grid_max * std::max(1, pl->ns) * std::max<int64_t>(1, d->n_secondary) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->lik_vec, B * (P + 2), true)) return rc;
    if (int rc = dev_alloc(pl, &pl->prior_grad, B * P, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->prior_lp, B * ((P + kPriorThreads - 1) / kPriorThreads + 1), true)) return rc;
    if (int rc = dev_alloc(pl, &pl->to_mu, B * P, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->to_sigma, B * P, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->glob_tot, (size_t)(kMaxTermsPerLevel + 1) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &pl->sec_tot, std::max<size_t>(1, pl->ns) * std::max<int64_t>(1, d->n_secondary) * bt, true)) return rc;
    if (int rc = dev_alloc(pl, &tmp, B * P + B, true)) return rc;
    pl->out_dev = tmp;
    CUDA_TRY(cudaMallocHost(&pl->h_in, std::max<size_t>(16, B * P * sizeof(T))));
    CUDA_TRY(cudaMallocHost(&pl->h_out, std::max<size_t>(16, (B * P + B) * sizeof(T))));

    pl->reduce = ReduceParams{};
    pl->reduce.csr_ptr = csr_dev;
    pl->reduce.entry_level = lvl_dev;
    pl->reduce.entry_t = t_dev;
    pl->reduce.entry_slot = slot_dev;
    pl->reduce.tile_keys = pl->tile_keys;
    pl->reduce.seg_offsets = pl->seg_off;
    pl->reduce.direct_p = pl->direct_p;
    pl->reduce.tile_head = pl->tile_head;
    pl->reduce.tile_tail = pl->tile_tail;
    pl->reduce.cta_glob = pl->cta_glob;
    pl->reduce.cta_sec = pl->cta_sec;
    pl->reduce.glob_tot = pl->glob_tot;
    pl->reduce.sec_tot = pl->sec_tot;
    pl->reduce.lik_vec = pl->lik_vec;
    pl->reduce.n_params = d->n_params;
    pl->reduce.n_primary = d->n_primary;
    pl->reduce.n_secondary = d->n_secondary;
    pl->reduce.ng = pl->ng;
    pl->reduce.np = pl->np;
    pl->reduce.ns = pl->ns;
    pl->reduce.rows_per_tile = WT;
    pl->reduce.blocks_p = pl->np > 0 ? (int)((d->n_primary + 255) / 256) : 0;
    pl->reduce.blocks_s = pl->ns > 0 ? (int)((d->n_secondary + 31) / 32) : 0;
    pl->reduce.logp_const = d->logp_const;
    return SKK_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluation
// ---------------------------------------------------------------------------------------------
template <typename T>
static int launch_rows(skk_plan *pl, const LaunchCfg &cfg, bool timed) {
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
    rp.tile_keys = pl->tile_keys;
    rp.seg_offsets = pl->seg_off;
    rp.n_primary = (int32_t)pl->n_primary;
    rp.n_secondary = (int32_t)pl->n_secondary;
    for (int t = 0; t < kMaxTermsPerLevel; ++t) {
        rp.xcol_g[t] = pl->xcol_g[t];
        rp.xcol_p[t] = pl->xcol_p[t];
        rp.xcol_s[t] = pl->xcol_s[t];
    }
    rp.val_g = static_cast<const T *>(pl->val_g);
    rp.val_p = static_cast<const T *>(pl->val_p);
    rp.val_s = static_cast<const T *>(pl->val_s);
    rp.direct_p = pl->direct_p;
    rp.tile_head = pl->tile_head;
    rp.tile_tail = pl->tile_tail;
    rp.cta_glob = pl->cta_glob;
    rp.cta_sec = pl->cta_sec;
    void *args[] = {&rp};
    if (timed) CUDA_TRY(cudaEventRecord(pl->evk0, pl->stream));
    CUDA_TRY(cudaLaunchKernel(cfg.fn, dim3(cfg.grid), dim3(pl->warps * 32), args, cfg.smem, pl->stream));
    if (timed) CUDA_TRY(cudaEventRecord(pl->evk1, pl->stream));
    pl->stats.launches++;
    pl->stats.row_kernel_launches++;
    return SKK_OK;
}

template <typename T>
static int eval_impl(skk_plan *pl, const void *theta, int batch, void *logp_out, void *grad_out, unsigned flags) {
    const size_t P = (size_t)pl->n_params;
    const bool dev_io = flags & SKK_EVAL_DEVICE_IO;
    const bool timed = flags & SKK_EVAL_TIME_KERNELS;
    cudaStream_t st = pl->stream;
    CUDA_TRY(cudaSetDevice(pl->device));
    CUDA_TRY(cudaEventRecord(pl->ev0, st));

    const T *th = static_cast<const T *>(pl->theta_dev);
    if (dev_io) {
        th = static_cast<const T *>(theta);
    } else {
        std::memcpy(pl->h_in, theta, (size_t)batch * P * sizeof(T));
        CUDA_TRY(cudaMemcpyAsync(pl->theta_dev, pl->h_in, (size_t)batch * P * sizeof(T), cudaMemcpyHostToDevice, st));
    }

    // K3: priors for the whole batch
    if (pl->include_priors && P > 0) {
        PriorParams pp = pl->prior;
        pp.batch = batch;
        skk_prior_kernel<T><<<(unsigned)((P + kPriorThreads - 1) / kPriorThreads), kPriorThreads, 0, st>>>(
            pp, th, pl->prior_grad, pl->prior_lp, pl->to_mu, pl->to_sigma);
        pl->stats.launches++;
        if (pp.n_parents > 0) {
            const int per_cta = pp.n_heavy > 0 ? 32 : 8;  // light parents: one warp each
            const unsigned blocks = (unsigned)(pp.n_heavy + (pp.n_parents - pp.n_heavy + per_cta - 1) / per_cta);
            skk_prior_link_kernel<<<blocks, pp.n_heavy > 0 ? 1024 : 256, 0, st>>>(pp, pl->prior_grad, pl->to_mu, pl->to_sigma);
            pl->stats.launches++;
        }
    }

    for (int b0 = 0; b0 < batch;) {
        const int remaining = batch - b0;
        const bool wide = remaining >= 2 && pl->cfg4.fn != nullptr;
        const int bt = wide ? 4 : 1;
        const int b_valid = std::min(bt, remaining);
        const LaunchCfg &cfg = wide ? pl->cfg4 : pl->cfg1;

        // K0: value tables in slot order
        if (pl->prep.n_terms > 0) {
            PrepParams pp = pl->prep;
            pp.bt = bt;
            pp.b_valid = b_valid;
            int64_t most = 1;
            for (int i = 0; i < pp.n_terms; ++i) most = std::max(most, pp.term[i].n_slots);
            dim3 grid((unsigned)std::min<int64_t>((most + 255) / 256, 1024), (unsigned)pp.n_terms);
            skk_prepare_kernel<T><<<grid, 256, 0, st>>>(pp, th + (size_t)b0 * P, static_cast<T *>(pl->val_g),
                                                        static_cast<T *>(pl->val_p), static_cast<T *>(pl->val_s));
            pl->stats.launches++;
        }
        if (pl->np > 0)
            CUDA_TRY(cudaMemsetAsync(pl->direct_p, 0, (size_t)pl->np * pl->n_primary * bt * sizeof(double), st));

        // K1: the rows
        if (int rc = launch_rows<T>(pl, cfg, timed)) return rc;

        // K4a: totals per slot, then per theta element, all in fixed order
        ReduceParams rp = pl->reduce;
        rp.bt = bt;
        rp.b0 = b0;
        rp.b_valid = b_valid;
        rp.grid = cfg.grid;
        const unsigned slot_blocks = (unsigned)(rp.blocks_p + rp.blocks_s + 1);
        const unsigned theta_blocks = (unsigned)std::max<size_t>(1, (P * 32 + 255) / 256);
        if (bt == 4) {
            skk_slot_total_kernel<4><<<slot_blocks, 256, 0, st>>>(rp);
            skk_theta_reduce_kernel<4><<<theta_blocks, 256, 0, st>>>(rp);
        } else {
            skk_slot_total_kernel<1><<<slot_blocks, 256, 0, st>>>(rp);
            skk_theta_reduce_kernel<1><<<theta_blocks, 256, 0, st>>>(rp);
        }
        pl->stats.launches += 2;
        b0 += b_valid;
    }

    // row-sharded: one all-reduce of [dlogp, logp partial, constant] per evaluation
    if (pl->comm) {
        NCCL_TRY(g_nccl.AllReduce(pl->lik_vec, pl->lik_vec, (size_t)batch * (P + 2), /*ncclDouble*/ 8, /*ncclSum*/ 0,
                                  pl->comm, st));
        pl->stats.launches++;
    }

    // K4b
    T *out_grad = static_cast<T *>(pl->out_dev);
    T *out_logp = out_grad + (size_t)pl->max_batch * P;
    if (dev_io) {
        out_grad = static_cast<T *>(grad_out);
        out_logp = static_cast<T *>(logp_out);
    }
    {
        FinalParams fp{pl->lik_vec, pl->prior_grad, pl->prior_lp, (int32_t)((P + kPriorThreads - 1) / kPriorThreads),
                       pl->n_params, pl->n_rows_total, pl->sigma_idx, pl->sigma_const, pl->family, pl->include_priors};
        dim3 grid((unsigned)std::max<size_t>(1, (P + 255) / 256), (unsigned)batch);
        skk_final_kernel<T><<<grid, 256, 0, st>>>(fp, th, out_logp, out_grad);
        pl->stats.launches++;
    }
    CUDA_TRY(cudaGetLastError());

    if (!dev_io) {
        // grad rows are contiguous for `batch` chains; logp sits after max_batch rows
        CUDA_TRY(cudaMemcpyAsync(pl->h_out, out_grad, (size_t)batch * P * sizeof(T), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(static_cast<T *>(pl->h_out) + (size_t)batch * P, out_logp, (size_t)batch * sizeof(T),
                                 cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaEventRecord(pl->ev1, st));
    if (dev_io && (flags & SKK_EVAL_NO_SYNC)) {
        pl->stats.evals++;
        return SKK_OK;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
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
    if (pl->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(pl->comm);
    for (void *p : pl->allocs) cudaFree(p);
    if (pl->h_in) cudaFreeHost(pl->h_in);
    if (pl->h_out) cudaFreeHost(pl->h_out);
    for (cudaEvent_t e : {pl->ev0, pl->ev1, pl->evk0, pl->evk1})
        if (e) cudaEventDestroy(e);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    delete pl;
}

int skk_plan_create(const skk_plan_desc *d, skk_plan **out) {
    if (!d || !out) return fail(SKK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (d->abi_version != SKK_ABI_VERSION)
        return fail(SKK_ERR_INVALID, "descriptor ABI %d, library ABI %d", d->abi_version, SKK_ABI_VERSION);
    if (d->dtype != SKK_F32 && d->dtype != SKK_F64) return fail(SKK_ERR_INVALID, "bad dtype %d", d->dtype);
    if (d->family < SKK_LIK_NORMAL || d->family > SKK_LIK_POISSON_LOG) return fail(SKK_ERR_UNSUPPORTED, "bad family %d", d->family);
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
        else if (td.xcol < -1 || td.xcol >= d->n_xcols) rc = fail(SKK_ERR_INVALID, "term %d: bad xcol", i);
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
    for (cudaEvent_t *e : {&pl->ev0, &pl->ev1, &pl->evk0, &pl->evk1}) {
        ce = cudaEventCreate(e);
        if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
    }
    rc = d->dtype == SKK_F64 ? create_impl<double>(d, pl) : create_impl<float>(d, pl);
    if (rc != SKK_OK) return guard(rc);
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return guard(fail(SKK_ERR_CUDA, "plan upload: %s", cudaGetErrorString(ce)));

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
    pl->nranks = nranks;
    pl->rank = rank;
    return SKK_OK;
}

int skk_comm_destroy(skk_plan *pl) {
    if (!pl) return fail(SKK_ERR_INVALID, "null plan");
    if (pl->comm) {
        CUDA_TRY(cudaStreamSynchronize(pl->stream));
        NCCL_TRY(g_nccl.CommDestroy(pl->comm));
        pl->comm = nullptr;
    }
    return SKK_OK;
}

int skk_stats(skk_plan *pl, skk_stats_t *out) {
    if (!pl || !out) return fail(SKK_ERR_INVALID, "null argument");
    *out = pl->stats;
    return SKK_OK;
}

void *skk_stream(skk_plan *pl) { return pl ? pl->stream : nullptr; }

}  // extern "C"
