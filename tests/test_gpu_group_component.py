"""
GroupComponent models at sizes that reach the tiled paths of the row kernel (the zoo model has 20 rows): a member-wise
coefficient is one term whose slot -> theta table points into several blocks, member-wise prior scales are per-element
"constant or element of theta"; built through the DSL, compared with the oracle, and with torch autograd over the
recorded graph at a smaller size.
"""
import numpy as np
import pandas as pd
import pytest

import sakkara_b200.model as api
from helpers import assert_parity
from sakkara_b200.model import dist as pm, sym

pytestmark = pytest.mark.gpu


def frame(n, sensors=4, times=60, seed=3):
    rng = np.random.default_rng(seed)
    sensor = rng.integers(0, sensors, n)
    time = rng.integers(0, times, n)
    x = rng.normal(size=n)
    eta = 0.3 * x * (sensor - 1.5) + 0.05 * np.sin(time)
    return pd.DataFrame({'sensor': [f's{v}' for v in sensor], 'time': time, 'x': x,
                         'y': rng.poisson(np.exp(np.clip(eta, -3, 3)))})


def likelihood(df, family):
    DC, GC = api.DistributionComponent, api.GroupComponent
    data = api.data_components(df[['x', 'y']])
    slope = GC('sensor', 'slope', {
        's0': DC(pm.Normal, name='slope_0', group='time', sigma=0.5),
        's1': DC(pm.Normal, name='slope_1', group='time', mu=DC(pm.Normal, name='level_1', sigma=0.3), sigma=0.5),
        's2': DC(pm.Normal, name='slope_2', sigma=0.4),
        's3': 0.25})
    spread = GC('sensor', 'spread', {'s0': 0.5, 's1': DC(pm.HalfNormal, name='spread_1', sigma=1.0),
                                     's2': DC(pm.Gamma, name='spread_2', alpha=3.0, beta=4.0), 's3': 0.8})
    level = DC(pm.Normal, name='level', group='sensor', mu=0.1, sigma=spread)
    eta = slope * data['x'] + level
    if family == 'Poisson':
        return api.Likelihood(pm.Poisson, mu=api.f_(sym.exp)(eta), observed=data['y'])
    return api.Likelihood(pm.Normal, mu=eta, sigma=DC(pm.HalfCauchy, name='noise', beta=1.0), observed=data['y'])


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('family', ['Poisson', 'Normal'])
@pytest.mark.parametrize('n', [3_000, 120_000])
def test_member_wise_model_parity(n, family, dtype):
    df = frame(n)
    model = api.build(df, likelihood(df, family), backend='b200')
    plan = model.plan
    assert any(t.block < 0 for t in plan.terms)                    # the member-wise coefficient reads several blocks
    assert any(p.is_mixed for b in plan.blocks for p in b.params.values())
    f = model.logp_dlogp_function(dtype=dtype, device=0)
    rng = np.random.default_rng(1)
    for _ in range(2):
        theta = rng.normal(0, 0.2, plan.n_params)
        logp, grad = f(theta.astype(dtype))
        assert_parity(logp, grad, plan, theta, dtype)
    f.close()
    if n <= 3_000 and dtype == 'float64':
        from oracle.autograd_check import logp_dlogp_autograd
        from oracle.logp_numpy import parity_error
        theta = rng.normal(0, 0.2, plan.n_params)
        l, g = logp_dlogp_autograd(model.recorded, theta)
        err_l, err_g = parity_error(l, g, plan, theta)
        assert err_l < 1e-12 and err_g < 1e-12


def sigma_frame(n, groups, seed=5):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, groups, n)
    x = rng.normal(size=n)
    sd = rng.uniform(0.3, 1.5, groups)[g]
    return pd.DataFrame({'group': g, 'x': x, 'y': 0.6 * x + 0.1 + sd * rng.normal(size=n)})


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('n,groups', [(2_000, 7), (150_000, 300)])
def test_sigma_per_group_as_a_parameter(n, groups, dtype):
    """sigma of the Normal likelihood is a positive block gathered by group: the row kernel gathers log(sigma) per
    row (SKK_LIK_NORMAL_LOGSIGMA); cross-checked with one standard plan per sigma on the small case."""
    df = sigma_frame(n, groups)
    DC = api.DistributionComponent

    def lik():
        data = api.data_components(df[['x', 'y']])
        k = DC(pm.Normal, name='k', group='group', mu=DC(pm.Normal, name='k_mean'), sigma=DC(pm.HalfNormal, name='k_sd'))
        sigma = DC(pm.HalfNormal, name='noise', group='group', sigma=1.0)
        return api.Likelihood(pm.Normal, mu=k * data['x'] + DC(pm.Normal, name='icpt'), sigma=sigma, observed=data['y'])
    model = api.build(df, lik(), backend='b200')
    plan = model.plan
    assert plan.likelihood.sigma_rows_element is not None
    assert len(np.unique(plan.likelihood.sigma_rows_element)) == groups
    f = model.logp_dlogp_function(dtype=dtype, device=0, max_batch=2)
    assert f.kernel_plan.family == 4               # one plan, priors included: nothing to add up on the host
    rng = np.random.default_rng(2)
    theta = rng.normal(0, 0.2, (2, plan.n_params))
    logp, grad = f(theta.astype(dtype))
    for b in range(2):
        assert_parity(logp[b], grad[b], plan, theta[b], dtype)
    one, g_one = f(theta[0].astype(dtype))
    assert_parity(one, g_one, plan, theta[0], dtype)
    f.close()
    if groups <= 16:
        g = model.logp_dlogp_function(dtype=dtype, device=0, per_element=True)
        assert len(g.parts) == groups and all(p.kernel_plan.family == 0 for p in g.parts)
        logp2, grad2 = g(theta[0].astype(dtype))
        assert_parity(logp2, grad2, plan, theta[0], dtype)
        g.close()
