"""
TEST INFRASTRUCTURE -- evaluates logp + dlogp from a *KernelPlan* (sorted rows, segment offsets,
slot -> theta tables, secondary codes) with plain NumPy.  It shares the formulas of
``oracle/logp_numpy.py`` but none of the structure decisions of the CUDA path, so comparing it
with the oracle on the ModelPlan pins ``make_kernel_plan`` (permutation, segments, slot tables,
uint8/uint16 narrowing, constants) on the CPU, without a GPU.
"""
import numpy as np

from oracle.logp_numpy import LOG_SQRT_2PI, sigmoid, softplus


def likelihood_from_kernel_plan(kp, theta):
    """Returns (logp_likelihood incl. constants, grad[P]) in float64."""
    theta = np.asarray(theta, dtype=np.float64)
    n = kp.n_rows
    eta = np.zeros(n) if kp.eta_offset is None else kp.eta_offset.astype(np.float64).copy()
    zeta = None
    slot_rows = []
    for t in kp.terms:
        if t.level == 0:
            slot = np.zeros(n, dtype=np.int64)
        elif t.level == 1:
            slot = np.repeat(np.arange(kp.n_primary), np.diff(kp.seg_offsets))
        else:
            slot = kp.sec_codes.astype(np.int64)
        slot_rows.append(slot)
        x = 1.0 if t.xcol < 0 else kp.xcols[t.xcol].astype(np.float64)
        if t.xcol == -2:            # a term of log(sigma) (SKK_XCOL_LOG_SIGMA, family 4)
            zeta = (zeta if zeta is not None else 0.0) + theta[t.theta_index.astype(np.int64)[slot]]
        else:
            eta += theta[t.theta_index.astype(np.int64)[slot]] * x
    y = kp.y.astype(np.float64)
    grad = np.zeros(kp.n_params)
    dz = None
    if kp.family == 5:              # StudentT, nu in the sigma field, sigma = exp(zeta) (1 without a zeta term)
        nu = kp.sigma_const
        zeta = np.zeros(n) if zeta is None else zeta
        sigma = np.exp(zeta)
        r = (y - eta) / sigma
        w = (nu + 1.0) / (nu + r * r)
        logp = np.sum(-0.5 * (nu + 1.0) * np.log1p(r * r / nu) - zeta)
        d = w * r / sigma
        dz = w * r * r - 1.0
    elif kp.family == 4:
        sigma = np.exp(zeta)
        r = (y - eta) / sigma
        logp = np.sum(-0.5 * r * r - zeta)
        d = r / sigma
        dz = r * r - 1.0
    elif kp.family == 0:
        sigma = np.exp(theta[kp.sigma_theta_index]) if kp.sigma_theta_index >= 0 else kp.sigma_const
        r = (y - eta) / sigma
        logp = np.sum(-0.5 * r * r) - n * (LOG_SQRT_2PI + np.log(sigma))
        d = r / sigma
        if kp.sigma_theta_index >= 0:
            grad[kp.sigma_theta_index] += np.sum(r * r - 1.0)
    elif kp.family == 1:
        logp = np.sum(y * eta - softplus(eta))
        d = y - sigmoid(eta)
    elif kp.family == 2:
        mu = np.exp(eta)
        logp = np.sum(y * eta - mu)
        d = y - mu
    elif kp.family == 6:
        alpha = kp.sigma_const                                   # Gamma(alpha, beta = alpha / exp(eta)): alpha in the sigma field
        w = y * np.exp(-eta)
        logp = np.sum(-alpha * (eta + w))
        d = alpha * (w - 1.0)
    else:
        alpha = kp.sigma_const                                   # NegativeBinomial: alpha in the sigma field
        log_sum = np.logaddexp(eta, np.log(alpha))
        logp = np.sum(y * eta - (y + alpha) * log_sum)
        d = y - (y + alpha) * np.exp(eta - log_sum)
    # -sum(lgamma(y + 1)) of a Poisson model; rows masked out for NaN observations (not the normalisers of the
    # prior families that the plan's constant also carries: this function is the likelihood alone)
    logp += kp.logp_const - kp.prior_logp_const
    for t, slot in zip(kp.terms, slot_rows):
        x = 1.0 if t.xcol < 0 else kp.xcols[t.xcol].astype(np.float64)
        np.add.at(grad, t.theta_index.astype(np.int64)[slot], dz if t.xcol == -2 else d * x)
    return logp, grad
