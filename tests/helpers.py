"""Shared test utilities."""
import numpy as np

import sakkara_b200.model as api
from sakkara_b200.model import dist, sym

# tolerances of BASELINE.json north_star, in the metric of SURVEY.md §8d
# (|got - ref| / max(|ref|, L1 mass of the reduction))
RTOL = {'float64': 1e-10, 'float32': 1e-4}


def build_model(entry):
    df, make = entry
    return api.build(df, make(api, dist, sym))


def theta_points(plan, n, seed=0, scale=0.5, include_zero=True):
    rng = np.random.default_rng(seed)
    pts = rng.normal(0.0, scale, (n, plan.n_params))
    if include_zero and n > 1:
        pts[0] = 0.0
    return pts


def assert_parity(got_logp, got_grad, plan, theta, dtype):
    """Compare one evaluation with the fp64 oracle at the same (dtype-rounded) point."""
    from oracle.logp_numpy import parity_error
    theta_r = np.asarray(theta, dtype=dtype).astype(np.float64)
    err_l, err_g = parity_error(got_logp, got_grad, plan, theta_r)
    tol = RTOL[np.dtype(dtype).name]
    assert np.isfinite(got_logp) and np.all(np.isfinite(got_grad))
    assert err_l <= tol, f'logp error {err_l:.3e} > {tol}'
    assert err_g <= tol, f'gradient error {err_g:.3e} > {tol}'
    return err_l, err_g
