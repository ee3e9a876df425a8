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

    # ---- priors, in creation order -----------------------------------------------------------
    for number, b in enumerate(plan.blocks):
        raw = theta[b.offset:b.offset + b.size]
        if b.family == 'Normal':
            mu_p, sd_p = b.params['mu'], b.params['sigma']
            mu = T(1) * mu_p.const.astype(dtype) if mu_p.is_const else _block_value(plan, theta, mu_p.block)[mu_p.index]
            sd = T(1) * sd_p.const.astype(dtype) if sd_p.is_const else _block_value(plan, theta, sd_p.block)[sd_p.index]
            mu = np.broadcast_to(mu, raw.shape)
            sd = np.broadcast_to(sd, raw.shape)
            z = (raw - mu) / sd
            terms = T(-0.5) * z * z - T(LOG_SQRT_2PI) - np.log(sd)
            logp += np.sum(terms, dtype=np.float64)
            logp_mass += np.sum(np.abs(terms), dtype=np.float64)
            acc(number, None, -z / sd)
            if not mu_p.is_const:
                # a mu that is a positive variable is sampled as u = log(mu): chain rule d mu / d u = mu
                d_mu = z / sd
                acc(mu_p.block, mu_p.index, d_mu * mu if plan.blocks[mu_p.block].transform == 'log' else d_mu)
            if not sd_p.is_const:
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
        else:
            raise NotImplementedError(b.family)

    # ---- likelihood ----------------------------------------------------------------------------
    lik = plan.likelihood
    y = lik.y.astype(dtype)
    eta = np.zeros(plan.n_rows, dtype=dtype) if lik.offset is None else lik.offset.astype(dtype).copy()
    columns = []
    for t in plan.terms:
        gathered = _block_value(plan, theta, t.block)[t.index]   # AdvancedSubtensor1
        x = None if t.x is None else t.x.astype(dtype)
        columns.append(x)
        eta += gathered if x is None else gathered * x

    if lik.family == 'Normal':
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
    elif lik.family == 'Bernoulli':
        terms = y * eta - softplus(eta)
        d_eta = y - sigmoid(eta)
    elif lik.family == 'Poisson':
        mu = np.exp(eta)
        terms = y * eta - mu - gammaln(lik.y + 1.0).astype(dtype)
        d_eta = y - mu
    else:
        raise NotImplementedError(lik.family)
    logp += np.sum(terms, dtype=np.float64)
    logp_mass += np.sum(np.abs(terms), dtype=np.float64)

    for t, x in zip(plan.terms, columns):
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
