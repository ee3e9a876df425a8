/*
 * skk_b200.h -- C ABI of the B200-native logp + dlogp evaluator (libskk_b200.so).
 *
 * This is the drop-in boundary of SURVEY.md §8(b).  The reference has no FFI of its own: the call
 * that this library replaces is the compiled function behind
 *     pymc.Model.logp_dlogp_function() -> ValueGradFunction.__call__(theta) -> (logp, dlogp)
 * (third-party pymc 5.3.1, pinned in /root/reference/poetry.lock:1012-1013), which PyMC's NUTS
 * invokes once per leapfrog step on the model that sakkara.model.build() assembled
 * (/root/reference/sakkara/model/utils.py:9-33, README.md:33-34).  Each entry point below states
 * the reference-side step it stands for.  Bindings: ctypes (sakkara_b200/runtime.py); the stub a
 * maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ or torch types cross the boundary.
 *   - Every function returns 0 on success or a negative SKK_ERR_* code; skk_last_error() returns a
 *     thread-local message for the last failure.  No exception ever crosses the boundary.
 *   - Ownership: the caller keeps every host array it passes; skk_plan_create copies what it needs
 *     to the device before returning and owns all device memory until skk_plan_destroy.
 *   - A plan is bound to the creating process and to one device.  Calls on one plan must be
 *     serialised by the caller (ctypes releases the GIL for the duration of a call).  Never create a
 *     plan before fork(); create it lazily in the worker process instead.
 *   - There is no CPU fallback: without a CUDA device skk_plan_create fails with SKK_ERR_CUDA.
 */
#ifndef SKK_B200_H
#define SKK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKK_ABI_VERSION 1

/* error codes */
#define SKK_OK 0
#define SKK_ERR_INVALID (-1)     /* malformed descriptor / argument */
#define SKK_ERR_UNSUPPORTED (-2) /* model shape outside what the kernels cover */
#define SKK_ERR_CUDA (-3)        /* CUDA runtime failure, including "no device" */
#define SKK_ERR_NCCL (-4)        /* collective failure: NCCL error, libnccl not loadable, peer time-out */

/* floating type of data, theta and results (pytensor.config.floatX on the reference side) */
#define SKK_F32 0
#define SKK_F64 1

/* likelihood families: the distribution given to sakkara.model.Likelihood
 * (/root/reference/sakkara/model/composable/hierarchical/likelihood.py:24-62) */
#define SKK_LIK_NORMAL 0         /* pm.Normal(mu=eta, sigma=...)          */
#define SKK_LIK_BERNOULLI_LOGIT 1 /* pm.Bernoulli(logit_p=eta)             */
#define SKK_LIK_POISSON_LOG 2    /* pm.Poisson(mu=exp(eta))               */
#define SKK_LIK_NEGBINOMIAL_LOG 3 /* pm.NegativeBinomial(mu=exp(eta), alpha=const): alpha travels in
                                    skk_plan_desc.sigma_const; lgamma(y+alpha) - lgamma(alpha) - lgamma(y+1)
                                    + alpha log(alpha) per row belongs into logp_const */

#define SKK_LIK_NORMAL_LOGSIGMA 4 /* pm.Normal(mu=eta, sigma=exp(zeta)): zeta = sum of the terms whose xcol is
                                    SKK_XCOL_LOG_SIGMA (sigma per group as a parameter; every such term reads
                                    log-transformed elements of theta).  -N log(sqrt(2 pi)) belongs into logp_const */
#define SKK_LIK_STUDENTT_LOGSIGMA 5 /* pm.StudentT(nu=const, mu=eta, sigma=exp(zeta)): nu travels in
                                    skk_plan_desc.sigma_const; zeta = sum of the SKK_XCOL_LOG_SIGMA terms (none: sigma = 1,
                                    i.e. rows standardised by a constant sigma); the normaliser
                                    N (lgamma((nu+1)/2) - lgamma(nu/2) - log(nu pi)/2) belongs into logp_const */
#define SKK_LIK_GAMMA_LOG 6      /* pm.Gamma(alpha=const, beta=alpha / exp(eta)), the Gamma GLM with mean exp(eta): alpha
                                    travels in skk_plan_desc.sigma_const; the kernel adds -alpha (eta + y exp(-eta)) per
                                    row, N (alpha log(alpha) - lgamma(alpha)) + (alpha - 1) sum log(y) belongs into
                                    logp_const */
/* skk_term_desc.xcol of a term of the log-scale predictor zeta of the *_LOGSIGMA families (no data column) */
#define SKK_XCOL_LOG_SIGMA (-2)

/* prior families of parameter blocks: the generator given to DistributionComponent
 * (/root/reference/sakkara/model/composable/hierarchical/distribution.py:41-51) */
#define SKK_PRIOR_NORMAL 0       /* a = mu, b = sigma                       */
#define SKK_PRIOR_HALFNORMAL 3   /* a = sigma; sampled as log, +log-Jacobian */
#define SKK_PRIOR_EXPONENTIAL 4  /* a = lam;   sampled as log, +log-Jacobian */
/* families with constant parameters only; the device evaluates the theta-dependent part, the
 * normaliser (a function of a, b alone) belongs into skk_plan_desc.logp_const */
#define SKK_PRIOR_HALFCAUCHY 5   /* a = beta;              log, +log-Jacobian */
#define SKK_PRIOR_GAMMA 6        /* a = alpha, b = beta;   log, +log-Jacobian */
#define SKK_PRIOR_INVGAMMA 7     /* a = alpha, b = beta;   log, +log-Jacobian */
#define SKK_PRIOR_CAUCHY 8       /* a = alpha, b = beta                       */
#define SKK_PRIOR_LAPLACE 9      /* a = mu,    b = b                          */
#define SKK_PRIOR_LOGNORMAL 10   /* a = mu,    b = sigma;  log, +log-Jacobian */
#define SKK_PRIOR_HALFSTUDENTT 11 /* a = nu,   b = sigma;  log, +log-Jacobian */

/* where the group index of a linear-predictor term comes from */
#define SKK_LEVEL_GLOBAL 0       /* one value for all rows                                    */
#define SKK_LEVEL_PRIMARY 1      /* rows are sorted by this level; index = segment number     */
#define SKK_LEVEL_SECONDARY 2    /* per-row code column; rows sorted by it inside a segment   */

/* storage of the observed column */
#define SKK_Y_FLOAT 0            /* same floating type as the plan                            */
#define SKK_Y_U8 1               /* uint8 (Bernoulli 0/1, small counts)                       */

/* skk_eval flags */
#define SKK_EVAL_DEVICE_IO 1u    /* theta / logp_out / grad_out are device pointers            */
#define SKK_EVAL_TIME_KERNELS 2u /* bracket the row kernel with events (see skk_stats)         */
#define SKK_EVAL_NO_SYNC 4u      /* with DEVICE_IO: enqueue only, do not synchronise the stream */

/*
 * One term of the linear predictor:  eta[row] += theta[theta_index[slot(row)]] * x[xcol][row].
 * Reference: a block gathered to the observations, element[tuple(mapping)]
 * (/root/reference/sakkara/relation/representation.py:222-253) times a data column
 * (/root/reference/sakkara/model/function/base.py:88-92).  slot(row) is 0 (GLOBAL), the number of
 * the primary segment the row lies in, or the row's secondary code.  Several slots may point at
 * the same theta element (a term defined on a parent of the sorted level).
 */
typedef struct {
    int32_t level;              /* SKK_LEVEL_*                                   */
    int32_t xcol;               /* index into skk_plan_desc.xcols, -1: x == 1, SKK_XCOL_LOG_SIGMA */
    int64_t n_slots;            /* 1, n_primary or n_secondary                   */
    const int32_t *theta_index; /* [n_slots] element of theta read by each slot  */
} skk_term_desc;

/*
 * One parameter block = one free random variable = theta[offset : offset+size]
 * (one DistributionComponent.build_variable call; creation order = layout of theta).
 * Parameter a / b of the prior is a constant (x_const_len = 1 or size) or, per element, another
 * element of theta (x_index != NULL; the prior-level gather of
 * /root/reference/sakkara/model/composable/hierarchical/base.py:46-55).  A linked sigma refers to
 * a log-transformed element u and means exp(u).  Both together (x_index != NULL, x_const != NULL,
 * x_const_len = size): element g reads theta[x_index[g]], or the constant x_const[g] where
 * x_index[g] is -1 -- the member-wise parameters of a GroupComponent
 * (/root/reference/sakkara/model/composable/group.py:69-103).
 */
typedef struct {
    int32_t family;             /* SKK_PRIOR_*                                   */
    int32_t reserved;
    int64_t offset;
    int64_t size;
    const double *a_const;
    int64_t a_const_len;        /* 0 when linked                                 */
    const int32_t *a_index;     /* [size] or NULL                                */
    const double *b_const;
    int64_t b_const_len;
    const int32_t *b_index;
} skk_block_desc;

/*
 * Flat kernel plan.  Rows are in kernel order: sorted by primary segment, and by secondary code
 * inside each segment.  Produced by sakkara_b200/plan/kernel_plan.py from the model that build()
 * lowered; stands for the constants PyTensor bakes into the compiled function.
 */
typedef struct {
    int32_t abi_version;        /* SKK_ABI_VERSION                               */
    int32_t dtype;              /* SKK_F32 / SKK_F64                             */
    int32_t family;             /* SKK_LIK_*                                     */
    int32_t device;             /* CUDA device ordinal                           */
    int64_t n_rows;             /* rows held by this plan (this rank's shard)    */
    int64_t n_params;           /* P = length of theta                           */
    int32_t max_batch;          /* largest batch skk_eval will be called with    */
    int32_t n_xcols;
    const void *const *xcols;   /* n_xcols host pointers, each [n_rows] of dtype */
    const void *y;              /* [n_rows], storage per y_kind                  */
    int32_t y_kind;             /* SKK_Y_*                                       */
    int32_t sec_code_bytes;     /* 2 (uint16) or 4 (int32)                       */
    const void *eta_offset;     /* optional [n_rows] of dtype, constant part of eta */
    int64_t n_primary;          /* 0: no primary level                           */
    const int64_t *seg_offsets; /* [n_primary+1], first row of every segment     */
    int64_t n_secondary;        /* 0: no secondary level                         */
    const void *sec_codes;      /* [n_rows] codes < n_secondary                  */
    int32_t n_terms;
    int32_t n_blocks;
    const skk_term_desc *terms;
    const skk_block_desc *blocks;
    double sigma_const;         /* Normal likelihood: fixed sigma ... (NegativeBinomial, Gamma: alpha; StudentT: nu) */
    int64_t sigma_theta_index;  /* ... or index of log(sigma) in theta; -1: const */
    double logp_const;          /* theta-independent part of logp of this shard
                                   (Poisson: -sum lgamma(y+1); the rows masked for NaN
                                   observations, likelihood.py:32-57, which the plan
                                   does not contain)                             */
    int64_t n_rows_total;       /* rows over all ranks (= n_rows on one GPU)     */
    int32_t include_priors;     /* 1: this rank adds the prior terms (after the all-reduce) */
    int32_t reserved;
} skk_plan_desc;

typedef struct skk_plan skk_plan;

typedef struct {
    uint64_t evals;             /* skk_eval calls completed                      */
    uint64_t launches;          /* kernels launched by this plan so far          */
    uint64_t row_kernel_launches;
    double last_row_kernel_ms;  /* with SKK_EVAL_TIME_KERNELS                    */
    double last_eval_ms;        /* device time of the last evaluation            */
    uint64_t stored_bytes_per_row; /* bytes of row data the row kernel reads per row */
    uint64_t device_bytes;      /* device memory owned by the plan               */
    int32_t grid;               /* CTAs of the row kernel                        */
    int32_t block;              /* threads per CTA                               */
    int32_t smem_bytes;
    int32_t n_tiles;
    int32_t rows_per_tile;
    int32_t kernel_variant;     /* 1: row-parallel (K1), 2: chain-per-lane (K2)  */
} skk_stats_t;

/* Replaces: compiling the model's logp_dlogp function (pymc.Model.logp_dlogp_function). */
int skk_plan_create(const skk_plan_desc *desc, skk_plan **out);

/*
 * Replaces: ValueGradFunction.__call__ -- one evaluation per row of theta.
 * theta [batch, P] row-major, logp_out [batch], grad_out [batch, P], all of the plan's dtype; host
 * pointers unless SKK_EVAL_DEVICE_IO.  Synchronous on return unless SKK_EVAL_NO_SYNC.
 */
int skk_eval(skk_plan *plan, const void *theta, int32_t batch, void *logp_out, void *grad_out, uint32_t flags);

int skk_plan_destroy(skk_plan *plan);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char *skk_last_error(void);

/*
 * Row-sharded evaluation over several GPUs (one process per GPU): every rank holds a plan over its
 * rows; after skk_comm_init each skk_eval ends with one all-reduce(sum) of [logp, dlogp] so that
 * all ranks return the full result.  unique_id: 128 bytes from skk_comm_unique_id on rank 0,
 * distributed by the caller (torch.distributed in sakkara_b200/distributed.py).
 * No counterpart in the reference (CPU only, no collectives).
 */
int skk_comm_unique_id(void *out128);
int skk_comm_init(skk_plan *plan, const void *unique_id128, int32_t rank, int32_t nranks);
int skk_comm_destroy(skk_plan *plan);

/*
 * One-shot all-reduce over NVLink peer memory, the default collective of row-sharded runs on one
 * box (<= 8 ranks): instead of ncclAllReduce, every rank stores its [dlogp, logp] vector into a
 * buffer of every peer (CUDA IPC mapping) and the final kernel adds the R vectors in rank order.
 *   skk_peer_alloc : allocates this rank's receive buffer for nranks peers, returns its 64-byte
 *                    cudaIpcMemHandle_t for the caller to all-gather;
 *   skk_peer_init  : handles = nranks x 64 bytes in rank order (own entry ignored).
 * After skk_peer_init every skk_eval ends with the peer exchange (takes precedence over NCCL).
 */
int skk_peer_alloc(skk_plan *plan, int32_t nranks, void *handle_out64);
int skk_peer_init(skk_plan *plan, const void *handles, int32_t rank, int32_t nranks);

int skk_stats(skk_plan *plan, skk_stats_t *out);

/* cudaStream_t the plan launches on (so that callers can record their own events on it). */
void *skk_stream(skk_plan *plan);

/*
 * Integrator update of a many-chain HMC / NUTS trajectory between two skk_eval calls, enqueued on the plan's
 * stream (never synchronises): for every chain c < chains and element j < P
 *     r[c][j] += a * eps[c] * grad[c][j];     q[c][j] += b * eps[c] * inv_mass[j] * r[c][j]    (b == 0: kick only)
 * q, r, grad [chains, P], eps [chains] (signed: the direction of the chain's trajectory), inv_mass [P]: device
 * pointers of the plan's dtype.  (a, b) = (0.5, 1) opens a leapfrog step, (0.5, 0) closes it, (1, 1) joins two.
 * Replaces, for all chains at once, the momentum / position updates of pymc's leapfrog integrator
 * (pymc/step_methods/hmc/integration.py, third-party; the reference reaches it through pm.sample(),
 * README.md:34).  Used by sakkara_b200/sampling.py.
 */
int skk_leapfrog(skk_plan *plan, void *q, void *r, const void *grad, const void *eps, const void *inv_mass,
                 int32_t chains, double a, double b);

/*
 * Diagnostics: per-warp timeline of the row kernel (%globaltimer at entry, end of the prologue, arrival of
 * the first tile, end of the tile loop, exit).  skk_trace_read copies [grid * warps][8] words of the last
 * evaluation (at most `capacity` words) and returns the number of words, or a negative SKK_ERR_* code.
 * No counterpart in the reference; used by tools/trace_rows.py to attribute the per-launch cost.
 */
int skk_trace_enable(skk_plan *plan, int32_t on);
int64_t skk_trace_read(skk_plan *plan, uint64_t *out, int64_t capacity);

/*
 * Measured ALU peak of the device, thread-instructions per second of a dependency-free loop: kind 0 = fp64 FMA,
 * 1 = fp32 FMA, 2 = MUFU.EX2 (multiply FMA rates by 2 for flop/s).  The denominator bench.py reports the
 * chain-per-lane kernel (configuration 5, fp64-ALU-bound) against; MEASURED_PEAKS.json has no such entry.
 */
int skk_alu_peak(int32_t device, int32_t kind, double *ops_per_second);

int skk_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SKK_B200_H */
