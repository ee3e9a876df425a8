"""
TEST INFRASTRUCTURE -- CPU oracle for logp + dlogp.  Never imported by the product path
(``sakkara_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs use it.

PARITY UNPINNED for the floating-point values: the arithmetic of this path lives in third-party
``pymc 5.3.1`` / ``pytensor 2.11.3`` (reference ``poetry.lock:1012-1013, 1050-1051``), which are not
vendored in ``/root/reference`` and cannot be installed in this image (no wheel, no network;
SURVEY.md §0, §8c), and the reference's tests never evaluate ``logp`` or its gradient.
This file therefore *restates* what the compiled PyTensor function computes for the models
``sakkara.model.build()`` assembles, op for op:

    forward : AdvancedSubtensor1 gathers ``theta_block[index]``  -> Elemwise linear predictor ->
              fused Elemwise logp of the family -> Sum
    backward: Elemwise d logp / d eta -> AdvancedIncSubtensor1 scatter-add per gathered block
              (``np.add.at``) -> chain rule through PyMC's LogTransform -> Join into dlogp[P]

and is cross-checked by an independent torch-autograd evaluation of the *recorded graph*
(``oracle/autograd_check.py``), by ``scipy.stats`` and by finite differences (tests/test_oracle.py).
The integer half (gather indices, theta layout) IS pinned: it is compared bit-for-bit with the
reference's own ``build()`` through ``oracle/ref_harness.py`` and ``tests/golden/``.

PyMC call sites in the reference that this restates: ``sakkara/model/utils.py:31`` (``pm.Model``),
``sakkara/model/composable/hierarchical/distribution.py:50`` (one logp term per distribution call),
``sakkara/model/fixed/data.py:41`` (``pm.ConstantData``), ``sakkara/relation/representation.py:253``
(advanced indexing), ``README.md:34`` (``pm.sample`` -> ``logp_dlogp_function`` per leapfrog step).
"""
from typing import Tuple

import numpy as np
from scipy.special import gammaln

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)
LOG_SQRT_2_OVER_PI = 0.5 * np.log(2.0 / np.pi)


def softplus(x: np.ndarray) -> np.ndarray:
    """log(1 + exp(x)) evaluated the way PyTensor's ``softplus`` does (piecewise, overflow-safe)."""
    return np.logaddexp(0.0, x)


def sigmoid(x: np.ndarray) -> np.ndarray:
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def _block_value(plan, theta, number):
    """Constrained value of a block: exp(u) for log-transformed blocks."""
    b = plan.blocks[number]
    raw = theta[b.offset:b.offset + b.size]
    return np.exp(raw) if b.transform == 'log' else raw


def _is_log(plan) -> np.ndarray:
    """Per element of theta: True where the element is sampled on the log scale."""
    flags = np.zeros(plan.n_params + 1, dtype=bool)         # (+1: index -1 of a member-wise parameter reads False)
    for b in plan.blocks:
        flags[b.offset:b.offset + b.size] = b.transform == 'log'
    return flags


def _all_values(plan, theta):
    """Constrained values of all of theta, block after block."""
    return np.concatenate([_block_value(plan, theta, k) for k in range(len(plan.blocks))]) if plan.blocks else theta


def _param_value(plan, theta, prm, dtype):
    """Value of a prior parameter per element: constant, gathered, or (member-wise) one or the other."""
    T = np.dtype(dtype).type
    if prm.is_const:
        return T(1) * prm.const.astype(dtype)
    if prm.is_mixed:
        values = _all_values(plan, theta)
        gathered = values[np.maximum(prm.index, 0)]
        return np.where(prm.index >= 0, gathered, np.broadcast_to(prm.const.astype(dtype), prm.index.shape))
    return _block_value(plan, theta, prm.block)[prm.index]


def logp_dlogp(plan, theta: np.ndarray, dtype=np.float64, return_mass: bool = False):
    """
    Joint log-posterior and gradient at one point.

    :param plan: :class:`sakkara_b200.plan.model_plan.ModelPlan`
    :param theta: raveled unconstrained parameter vector, length ``plan.n_params``
    :param dtype: ``float64`` (PyTensor default) or ``float32`` (``floatX=float32``: elementwise
        maths and scatter-adds in fp32, ``Sum`` accumulates in fp64; SURVEY.md §A "dtype rules")
    :param return_mass: also return, per gradient entry and for logp, the L1 mass of the reduction
        that produced it (denominator of the error metric)
    :return: ``(logp, dlogp[P])`` or ``(logp, dlogp, logp_mass, dlogp_mass)``
    """
    T = np.dtype(dtype).type
    theta = np.asarray(theta, dtype=dtype)
    P = plan.n_params
    if theta.shape != (P,):
        raise ValueError(f'theta must have shape ({P},)')
    grad = np.zeros(P, dtype=dtype)
    mass = np.zeros(P, dtype=np.float64)
    logp = np.float64(0.0)
    logp_mass = np.float64(0.0)

    def acc(target_block, index, values):
        b = plan.blocks[target_block]
        view = grad[b.offset:b.offset + b.size]
        if index is None:
            view += values
            mass[b.offset:b.offset + b.size] += np.abs(values)
        else:
            np.add.at(view, index, values)                       # AdvancedIncSubtensor1
            np.add.at(mass[b.offset:b.offset + b.size], index, np.abs(values).astype(np.float64))

    def acc_absolute(index, values):
        np.add.at(grad, index, values)
        np.add.at(mass, index, np.abs(values).astype(np.float64))

    # ---- priors, in creation order -----------------------------------------------------------
    for number, b in enumerate(plan.blocks):
        raw = theta[b.offset:b.offset + b.size]
        if b.family == 'Normal':
            mu_p, sd_p = b.params['mu'], b.params['sigma']
            mu = np.broadcast_to(_param_value(plan, theta, mu_p, dtype), raw.shape)
            sd = np.broadcast_to(_param_value(plan, theta, sd_p, dtype), raw.shape)
            z = (raw - mu) / sd
            terms = T(-0.5) * z * z - T(LOG_SQRT_2PI) - np.log(sd)
            logp += np.sum(terms, dtype=np.float64)
            logp_mass += np.sum(np.abs(terms), dtype=np.float64)
            acc(number, None, -z / sd)
            if mu_p.is_mixed:
                # member-wise parameter: per element a constant (index -1) or an element of theta
                linked = mu_p.index >= 0
                d_mu = z / sd
                acc_absolute(mu_p.index[linked], np.where(_is_log(plan)[mu_p.index], d_mu * mu, d_mu)[linked])
            elif not mu_p.is_const:
                # a mu that is a positive variable is sampled as u = log(mu): chain rule d mu / d u = mu
                d_mu = z / sd
                acc(mu_p.block, mu_p.index, d_mu * mu if plan.blocks[mu_p.block].transform == 'log' else d_mu)
            if sd_p.is_mixed:
                linked = sd_p.index >= 0
                acc_absolute(sd_p.index[linked], (z * z - T(1))[linked])
            elif not sd_p.is_const:
                # d/d sigma = (z^2 - 1)/sigma, times d sigma / d u = sigma
                acc(sd_p.block, sd_p.index, z * z - T(1))
        elif b.family == 'HalfNormal':
            tau = np.broadcast_to(b.params['sigma'].const.astype(dtype), raw.shape)
            s = np.exp(raw)
            q = s / tau
            terms = T(-0.5) * q * q + T(LOG_SQRT_2_OVER_PI) - np.log(tau) + raw
            logp += np.sum(terms, dtype=np.float64)
            logp_mass += np.sum(np.abs(terms), dtype=np.float64)
            acc(number, None, -q * q + T(1))
        elif b.family == 'Exponential':
            lam = np.broadcast_to(b.params['lam'].const.astype(dtype), raw.shape)
            s = np.exp(raw)
            terms = np.log(lam) - lam * s + raw
            logp += np.sum(terms, dtype=np.float64)
            logp_mass += np.sum(np.abs(terms), dtype=np.float64)
            acc(number, None, -lam * s + T(1))
        elif b.family in ('HalfCauchy', 'Gamma', 'InverseGamma', 'Cauchy', 'Laplace', 'LogNormal', 'HalfStudentT'):
            # constant parameters only; PyMC's logp of the family (pymc/distributions/continuous.py) at the
            # back-transformed value, plus the log-Jacobian of the LogTransform (= raw) for the positive ones
            def const(name):
                return np.broadcast_to(b.params[name].const.astype(dtype), raw.shape)
            if b.family == 'HalfCauchy':
                beta = const('beta')
                x = np.exp(raw)
                terms = T(np.log(2.0 / np.pi)) - np.log(beta) - np.log1p((x / beta) ** 2) + raw
                g = -T(2) * x * x / (beta * beta + x * x) + T(1)
            elif b.family == 'Gamma':
                alpha, beta = const('alpha'), const('beta')
                x = np.exp(raw)
                terms = alpha * np.log(beta) - gammaln(alpha).astype(dtype) + (alpha - T(1)) * raw - beta * x + raw
                g = (alpha - T(1)) - beta * x + T(1)
            elif b.family == 'InverseGamma':
                alpha, beta = const('alpha'), const('beta')
                x = np.exp(raw)
                terms = alpha * np.log(beta) - gammaln(alpha).astype(dtype) - (alpha + T(1)) * raw - beta / x + raw
                g = -(alpha + T(1)) + beta / x + T(1)
            elif b.family == 'Cauchy':
                loc, beta = const('alpha'), const('beta')
                z = (raw - loc) / beta
                terms = -T(np.log(np.pi)) - np.log(beta) - np.log1p(z * z)
                g = -T(2) * z / (beta * (T(1) + z * z))
            elif b.family == 'Laplace':
                mu, scale = const('mu'), const('b')
                terms = -np.log(T(2) * scale) - np.abs(raw - mu) / scale
                g = -np.sign(raw - mu) / scale
            elif b.family == 'LogNormal':
                mu, sd = const('mu'), const('sigma')
                z = (raw - mu) / sd                             # raw = log(x)
                terms = T(-0.5) * z * z - T(LOG_SQRT_2PI) - np.log(sd) - raw + raw
                g = -z / sd
            else:                                               # HalfStudentT
                nu, sd = const('nu'), const('sigma')
                x = np.exp(raw)
                q = (x / sd) ** 2 / nu
                terms = (T(np.log(2.0)) + (gammaln((nu + 1) / 2) - gammaln(nu / 2)).astype(dtype)
                         - T(0.5) * np.log(nu * T(np.pi) * sd * sd) - (nu + T(1)) / T(2) * np.log1p(q) + raw)
                g = -(nu + T(1)) * q / (T(1) + q) + T(1)
            logp += np.sum(terms, dtype=np.float64)
            logp_mass += np.sum(np.abs(terms), dtype=np.float64)
            acc(number, None, g.astype(dtype))
        else:
            raise NotImplementedError(b.family)

    # ---- likelihood ----------------------------------------------------------------------------
    lik = plan.likelihood
    y = lik.y.astype(dtype)
    eta = np.zeros(plan.n_rows, dtype=dtype) if lik.offset is None else lik.offset.astype(dtype).copy()
    columns = []
    for t in plan.terms:
        if t.block < 0:     # member-wise coefficient: absolute elements of theta, several blocks
            gathered = _all_values(plan, theta)[t.index]
        else:
            gathered = _block_value(plan, theta, t.block)[t.index]   # AdvancedSubtensor1
        x = None if t.x is None else t.x.astype(dtype)
        columns.append(x)
        eta += gathered if x is None else gathered * x

    if lik.family == 'Normal' and getattr(lik, 'sigma_rows_element', None) is not None:
        # sigma differs between rows: a constant, or exp(theta[element]) (sigma per group as a parameter)
        element = lik.sigma_rows_element
        linked = element >= 0
        sigma = np.where(linked, np.exp(theta[np.maximum(element, 0)]), lik.sigma_rows_const.astype(dtype)).astype(dtype)
        r = (y - eta) / sigma
        terms = T(-0.5) * r * r - T(LOG_SQRT_2PI) - np.log(sigma)
        d_eta = r / sigma
        acc_absolute(element[linked], (r * r - T(1))[linked])       # d/d log(sigma) = r^2 - 1
    elif lik.family == 'Normal':
        if lik.sigma_block is None:
            sigma = T(lik.sigma_const)
        else:
            sigma = _block_value(plan, theta, lik.sigma_block)[lik.sigma_index]
        r = (y - eta) / sigma
        terms = T(-0.5) * r * r - T(LOG_SQRT_2PI) - np.log(sigma)
        d_eta = r / sigma
        if lik.sigma_block is not None:
            sb = plan.blocks[lik.sigma_block]
            contrib = r * r - T(1)
            grad[sb.offset + lik.sigma_index] += np.sum(contrib, dtype=np.float64)
            mass[sb.offset + lik.sigma_index] += np.sum(np.abs(contrib), dtype=np.float64)
    elif lik.family == 'StudentT':
        # pymc StudentT(nu, mu, sigma): lgamma((nu+1)/2) - lgamma(nu/2) - log(nu pi)/2 - log(sigma) - (nu+1)/2 log1p(r^2/nu)
        nu = T(lik.nu)
        element = getattr(lik, 'sigma_rows_element', None)
        if element is None:
            sigma = np.ones(plan.n_rows, dtype=dtype)            # rows standardised by their constant sigma
            linked = np.zeros(plan.n_rows, dtype=bool)
        else:
            linked = element >= 0
            sigma = np.where(linked, np.exp(theta[np.maximum(element, 0)]), lik.sigma_rows_const.astype(dtype)).astype(dtype)
        r = (y - eta) / sigma
        norm = T(gammaln(0.5 * (float(nu) + 1.0)) - gammaln(0.5 * float(nu)) - 0.5 * np.log(float(nu) * np.pi))
        terms = norm - np.log(sigma) - T(0.5) * (nu + T(1)) * np.log1p(r * r / nu)
        w = (nu + T(1)) / (nu + r * r)
        d_eta = w * r / sigma
        if element is not None:
            acc_absolute(element[linked], (w * r * r - T(1))[linked])
    elif lik.family == 'Bernoulli':
        terms = y * eta - softplus(eta)
        d_eta = y - sigmoid(eta)
    elif lik.family == 'Poisson':
        mu = np.exp(eta)
        terms = y * eta - mu - gammaln(lik.y + 1.0).astype(dtype)
        d_eta = y - mu
    elif lik.family == 'NegativeBinomial':
        # pymc NegativeBinomial(mu, alpha): binomln(y + alpha - 1, y) + y log(mu / (mu + alpha)) + alpha log(alpha / (mu + alpha))
        alpha = T(lik.sigma_const)
        mu = np.exp(eta)
        log_sum = np.logaddexp(eta, np.log(alpha))                 # log(mu + alpha)
        const = (gammaln(lik.y + float(alpha)) - gammaln(lik.y + 1.0) - gammaln(float(alpha))).astype(dtype)
        terms = const + y * (eta - log_sum) + alpha * (np.log(alpha) - log_sum)
        d_eta = y - (y + alpha) * mu / (mu + alpha)
    elif lik.family == 'Gamma':
        # pymc Gamma(alpha, beta): alpha log(beta) - lgamma(alpha) + (alpha - 1) log(y) - beta y, with the rate of
        # the log-link GLM beta = alpha / mu, mu = exp(eta)
        alpha = T(lik.sigma_const)
        w = y * np.exp(-eta)                                       # y / mu
        const = ((float(alpha) - 1.0) * np.log(lik.y) + float(alpha) * np.log(float(alpha)) - gammaln(float(alpha))).astype(dtype)
        terms = const - alpha * (eta + w)
        d_eta = alpha * (w - T(1))
    else:
        raise NotImplementedError(lik.family)
    logp += np.sum(terms, dtype=np.float64)
    logp_mass += np.sum(np.abs(terms), dtype=np.float64)

    for t, x in zip(plan.terms, columns):
        if t.block < 0:
            if _is_log(plan)[t.index].any():
                raise NotImplementedError('positive blocks in the linear predictor')
            acc_absolute(t.index, d_eta if x is None else d_eta * x)
            continue
        if plan.blocks[t.block].transform != 'identity':
            raise NotImplementedError('positive blocks in the linear predictor')
        acc(t.block, t.index, d_eta if x is None else d_eta * x)

    logp += np.float64(getattr(plan, 'logp_const', 0.0))      # rows masked out for NaN observations
    logp_mass += abs(np.float64(getattr(plan, 'logp_const', 0.0)))
    logp_out = np.dtype(dtype).type(logp)
    if return_mass:
        return logp_out, grad, logp_mass, mass
    return logp_out, grad


def logp_dlogp_batch(plan, thetas: np.ndarray, dtype=np.float64) -> Tuple[np.ndarray, np.ndarray]:
    """Row-wise evaluation of a ``[B, P]`` batch of parameter vectors."""
    thetas = np.atleast_2d(thetas)
    out_l = np.empty(thetas.shape[0], dtype=dtype)
    out_g = np.empty(thetas.shape, dtype=dtype)
    for i, th in enumerate(thetas):
        out_l[i], out_g[i] = logp_dlogp(plan, th, dtype=dtype)
    return out_l, out_g


def parity_error(got_logp, got_grad, plan, theta, floor: float = 1.0) -> Tuple[float, float]:
    """
    Error of ``(got_logp, got_grad)`` against the fp64 oracle in the metric of SURVEY.md §8d:
    ``|got - ref| / max(|ref|, L1 mass of the reduction)`` -- plain relative error is ill-defined
    for gradient entries that cancel to ~0 near the mode.  Returns ``(err_logp, max err_grad)``.
    """
    ref_l, ref_g, mass_l, mass_g = logp_dlogp(plan, np.asarray(theta, dtype=np.float64), np.float64, True)
    den_l = max(abs(float(ref_l)), float(mass_l), floor * np.finfo(np.float64).tiny)
    err_l = abs(float(got_logp) - float(ref_l)) / den_l
    den_g = np.maximum(np.maximum(np.abs(ref_g), mass_g), np.finfo(np.float64).tiny)
    err_g = float(np.max(np.abs(np.asarray(got_grad, dtype=np.float64) - ref_g) / den_g)) if len(ref_g) else 0.0
    return err_l, err_g
