/*
 * TEST INFRASTRUCTURE -- CPU restatement (plain C + OpenMP) of the likelihood part of logp + dlogp,
 * used as the multi-core CPU baseline of bench.py and as a second CPU implementation next to
 * oracle/logp_numpy.py.  Never linked into or called from the product path (sakkara_b200/).
 *
 * PARITY UNPINNED for floating-point values (see the header of oracle/logp_numpy.py): the arithmetic
 * of this path lives in pymc 5.3.1 / pytensor 2.11.3 (/root/reference/poetry.lock:1012-1013,
 * 1050-1051), neither vendored nor installable here.  The formulas below restate what the compiled
 * PyTensor function evaluates for the graph that sakkara.model.build() assembles:
 *
 *   eta_i   = offset_i + sum_t theta[base_t + index_t[i]] * x_t[i]     gathers element[tuple(mapping)],
 *                                                                      /root/reference/sakkara/relation/representation.py:253;
 *                                                                      arithmetic /root/reference/sakkara/model/function/base.py:88-92
 *   Normal    : logp_i = -0.5 r_i^2 - log(sqrt(2 pi)) - log(sigma), r_i = (y_i - eta_i)/sigma, d_i = r_i/sigma
 *   Bernoulli : logp_i = y_i eta_i - softplus(eta_i),                 d_i = y_i - sigmoid(eta_i)
 *   Poisson   : logp_i = y_i eta_i - exp(eta_i) - lgamma(y_i + 1),    d_i = y_i - exp(eta_i)
 *   Gamma     : logp_i = -alpha (eta_i + y_i exp(-eta_i)) + const_i,                    d_i = alpha (y_i exp(-eta_i) - 1)
 *   NegBinomial: logp_i = y_i eta_i - (y_i + alpha) log(alpha + exp(eta_i)) + const_i,  d_i = y_i - (y_i + alpha) mu_i / (alpha + mu_i)
 *               (one observed distribution call, /root/reference/sakkara/model/composable/hierarchical/likelihood.py:58-62)
 *   dlogp[base_t + index_t[i]] += d_i * x_t[i]                        (scatter-add, AdvancedIncSubtensor1)
 *
 * Unlike PyTensor's C backend (one thread, one pass per op) the loop is fused and parallel over rows;
 * it is therefore a *stronger* CPU baseline than the reference's own path.  Every thread owns a
 * private gradient vector; the vectors are added in thread order, so results are reproducible for a
 * fixed thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_TERMS 8
#define LOG_SQRT_2PI 0.91893853320467274178

typedef struct {
    int32_t family;      /* 0 Normal, 1 Bernoulli-logit, 2 Poisson-log, 3 NegativeBinomial-log, 4 Gamma-log (alpha in sigma) */
    int32_t n_terms;
    int64_t n_rows;
    int64_t n_params;
    const int32_t *index[MAX_TERMS];   /* per-row element of the term's block */
    const void *x[MAX_TERMS];          /* per-row multiplier or NULL (= 1) */
    int64_t base[MAX_TERMS];           /* offset of the block in theta */
    const void *y;
    const void *offset;                /* optional constant part of eta */
    double sigma;                      /* Normal only */
    int64_t sigma_theta_index;         /* -1: constant */
} glm_desc;

#define DEFINE_LIKELIHOOD(NAME, REAL, EXP, LOG1P, FABS, LOGF)                                                  \
    double NAME(const glm_desc *d, const REAL *theta, double *grad, int threads) {                            \
        const REAL *y = (const REAL *)d->y;                                                                    \
        const REAL *off = (const REAL *)d->offset;                                                             \
        const int64_t P = d->n_params, N = d->n_rows;                                                          \
        const int T = d->n_terms;                                                                              \
        int nt = threads > 0 ? threads : 1;                                                                    \
        double *local = (double *)calloc((size_t)nt * (size_t)(P + 2), sizeof(double));                       \
        const REAL sigma = (REAL)d->sigma, inv_sigma = (REAL)1 / sigma;                                        \
        _Pragma("omp parallel num_threads(nt)") {                                                              \
            const int tid = omp_get_thread_num();                                                              \
            double *g = local + (size_t)tid * (size_t)(P + 2);                                                 \
            double lp = 0.0, sr2 = 0.0;                                                                        \
            _Pragma("omp for schedule(static)") for (int64_t i = 0; i < N; ++i) {                             \
                REAL eta = off ? off[i] : (REAL)0;                                                             \
                for (int t = 0; t < T; ++t) {                                                                  \
                    const REAL v = theta[d->base[t] + d->index[t][i]];                                         \
                    eta += d->x[t] ? v * ((const REAL *)d->x[t])[i] : v;                                       \
                }                                                                                              \
                REAL dd, l;                                                                                    \
                if (d->family == 0) {                                                                          \
                    const REAL r = (y[i] - eta) * inv_sigma;                                                   \
                    l = (REAL)-0.5 * r * r;                                                                    \
                    dd = r * inv_sigma;                                                                        \
                    sr2 += (double)(r * r);                                                                    \
                } else if (d->family == 1) {                                                                   \
                    const REAL e = EXP(-FABS(eta));                                                            \
                    const REAL inv = (REAL)1 / ((REAL)1 + e);                                                  \
                    l = y[i] * eta - ((eta > 0 ? eta : (REAL)0) + LOG1P(e));                                   \
                    dd = y[i] - (eta >= 0 ? inv : e * inv);                                                    \
                } else if (d->family == 2) {                                                                   \
                    const REAL mu = EXP(eta);                                                                  \
                    l = y[i] * eta - mu;                                                                       \
                    dd = y[i] - mu;                                                                            \
                } else if (d->family == 4) { /* Gamma(alpha = d->sigma, beta = alpha / exp(eta)), constants by the caller */ \
                    const REAL alpha = sigma, w = y[i] * EXP(-eta);                                            \
                    l = -alpha * (eta + w);                                                                    \
                    dd = alpha * (w - (REAL)1);                                                                \
                } else { /* NegativeBinomial(mu = exp(eta), alpha = d->sigma), constants added by the caller */ \
                    const REAL alpha = sigma, mu = EXP(eta);                                                   \
                    const REAL lam = eta > (REAL)30 ? eta + LOG1P(alpha * EXP(-eta)) : LOG1P(mu / alpha) + LOGF(alpha); \
                    l = y[i] * eta - (y[i] + alpha) * lam;                                                     \
                    dd = y[i] - (y[i] + alpha) * (eta > (REAL)30 ? (REAL)1 / ((REAL)1 + alpha * EXP(-eta)) : mu / (alpha + mu)); \
                }                                                                                              \
                lp += (double)l;                                                                               \
                for (int t = 0; t < T; ++t)                                                                    \
                    g[d->base[t] + d->index[t][i]] += (double)(d->x[t] ? dd * ((const REAL *)d->x[t])[i] : dd); \
            }                                                                                                  \
            g[P] = lp;                                                                                         \
            g[P + 1] = sr2;                                                                                    \
        }                                                                                                      \
        double logp = 0.0, sr2 = 0.0;                                                                          \
        for (int t = 0; t < nt; ++t) {                                                                         \
            const double *g = local + (size_t)t * (size_t)(P + 2);                                             \
            for (int64_t j = 0; j < P; ++j) grad[j] += g[j];                                                   \
            logp += g[P];                                                                                      \
            sr2 += g[P + 1];                                                                                   \
        }                                                                                                      \
        free(local);                                                                                           \
        if (d->family == 0) {                                                                                  \
            logp -= (double)N * (LOG_SQRT_2PI + log(d->sigma));                                                \
            if (d->sigma_theta_index >= 0) grad[d->sigma_theta_index] += sr2 - (double)N;                      \
        }                                                                                                      \
        return logp;                                                                                           \
    }

#ifndef _OPENMP
static int omp_get_thread_num(void) { return 0; }
#endif

DEFINE_LIKELIHOOD(skk_oracle_likelihood_f64, double, exp, log1p, fabs, log)
DEFINE_LIKELIHOOD(skk_oracle_likelihood_f32, float, expf, log1pf, fabsf, logf)

int skk_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
