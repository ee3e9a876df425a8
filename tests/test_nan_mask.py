"""
NaN-masked likelihoods (reference ``sakkara/model/composable/hierarchical/likelihood.py:32-57``):
rows whose observation is NaN get constant distribution parameters (``nan_param_mask``) and a
constant observation (``nan_data_mask``), i.e. they add a constant to logp and nothing to the
gradient.  Here the rows are dropped from the plan and the constant is kept.  Checked three ways:
the plan-level oracle against torch autograd over the *recorded graph* (which still holds all rows,
with ``where(mask, fill, value)`` parameters), against scipy's densities, and against the same model
built on the frame without the NaN rows; the kernel plan is pinned with the NumPy evaluator.
"""
import warnings

import numpy as np
import pandas as pd
import pytest
from scipy import stats

import sakkara_b200.model as api
from helpers import assert_parity, theta_points
from oracle.autograd_check import logp_dlogp_autograd
from oracle.kernel_plan_eval import likelihood_from_kernel_plan
from oracle.logp_c import priors_only
from oracle.logp_numpy import logp_dlogp
from sakkara_b200.model import dist, sym
from sakkara_b200.plan.kernel_plan import make_kernel_plan


def frame(n=240, groups=6, seed=3, nan_rows=(5, 17, 40, 41, 199)):
    rng = np.random.default_rng(seed)
    code = np.concatenate([np.arange(groups), rng.integers(0, groups, n - groups)])   # every group appears early
    x = rng.normal(size=n)
    df = pd.DataFrame({'g': [f'g{c}' for c in code], 'x': x})
    eta = 0.6 * x + 0.2 * (code % 3)
    df['y_normal'] = eta + 0.3 * rng.normal(size=n)
    df['y_bern'] = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    df['y_pois'] = rng.poisson(np.exp(eta)).astype(float)
    nan_rows = np.asarray(nan_rows)
    for col in ('y_normal', 'y_bern', 'y_pois'):
        df.loc[nan_rows, col] = np.nan
    return df, nan_rows


def likelihood(df, family, masked=True):
    data = api.data_components(df)
    k = api.DistributionComponent(dist.Normal, name='k', group='g',
                                  mu=api.DistributionComponent(dist.Normal, name='mu_k'),
                                  sigma=api.DistributionComponent(dist.Exponential, name='sd_k', lam=1.0))
    icpt = api.DistributionComponent(dist.Normal, name='icpt')
    eta = k * data['x'] + icpt
    mask = {}
    if family == 'normal':
        kw = dict(mu=eta, sigma=api.DistributionComponent(dist.Exponential, name='sd', lam=1.0))
        mask = {'mu': 0.5, 'sigma': 2.0}
        gen, obs = dist.Normal, data['y_normal']
    elif family == 'bernoulli':
        kw = dict(logit_p=eta)
        mask = {'logit_p': 0.3}
        gen, obs = dist.Bernoulli, data['y_bern']
    else:
        kw = dict(mu=api.f_(sym.exp)(eta))
        mask = {'mu': 1.7}
        gen, obs = dist.Poisson, data['y_pois']
    if masked:
        return api.Likelihood(gen, observed=obs, nan_param_mask=mask, nan_data_mask=0, **kw)
    return api.Likelihood(gen, observed=obs, **kw)


def build(df, family, masked=True):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', UserWarning)
        return api.build(df, likelihood(df, family, masked))


EXPECTED = {
    'normal': lambda: stats.norm.logpdf(0.0, 0.5, 2.0),
    'bernoulli': lambda: stats.bernoulli.logpmf(0, 1 / (1 + np.exp(-0.3))),
    'poisson': lambda: stats.poisson.logpmf(0, 1.7),
}


@pytest.mark.parametrize('family', ['normal', 'bernoulli', 'poisson'])
def test_masked_rows_are_a_constant(family):
    df, nan_rows = frame()
    model = build(df, family)
    plan = model.plan
    assert plan.n_rows == len(df) - len(nan_rows) and plan.n_masked_rows == len(nan_rows)
    assert plan.logp_const == pytest.approx(len(nan_rows) * EXPECTED[family](), rel=1e-13)
    assert all(len(t.index) == plan.n_rows for t in plan.terms) and len(plan.likelihood.y) == plan.n_rows
    for theta in theta_points(plan, 4, seed=2):
        logp, grad = logp_dlogp(plan, theta)
        ref_logp, ref_grad = logp_dlogp_autograd(model.recorded, theta)     # all rows, where(mask, fill, value)
        assert logp == pytest.approx(ref_logp, rel=1e-12)
        np.testing.assert_allclose(grad, ref_grad, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('family', ['normal', 'bernoulli', 'poisson'])
def test_same_as_the_model_without_those_rows(family):
    df, nan_rows = frame()
    masked = build(df, family).plan
    dropped = build(df.drop(index=nan_rows).reset_index(drop=True), family, masked=False).plan
    assert [b.name for b in masked.blocks] == [b.name for b in dropped.blocks]
    for theta in theta_points(masked, 3, seed=5):
        lm, gm = logp_dlogp(masked, theta)
        ld, gd = logp_dlogp(dropped, theta)
        assert lm == pytest.approx(ld + masked.logp_const, rel=1e-13)
        np.testing.assert_allclose(gm, gd, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('family', ['normal', 'bernoulli', 'poisson'])
def test_kernel_plan_keeps_the_constant(family, dtype):
    """also: a group whose rows are all NaN keeps its parameter (prior only) and gets no segment"""
    df, _ = frame(nan_rows=(3, 17, 40))
    df.loc[df['g'] == 'g3', ['y_normal', 'y_bern', 'y_pois']] = np.nan
    plan = build(df, family).plan
    assert 'g3' in list(plan.coords['g'])
    kp = make_kernel_plan(plan, dtype=dtype)
    assert kp.n_rows == plan.n_rows
    assert kp.n_primary == 5 and np.all(np.diff(kp.seg_offsets) > 0)   # the group without rows has no segment
    for theta in theta_points(plan, 3, seed=7):
        theta = theta.astype(dtype).astype(np.float64)
        lik_l, lik_g = likelihood_from_kernel_plan(kp, theta)
        pr_l, pr_g = logp_dlogp(priors_only(plan), theta)
        ref_l, ref_g = logp_dlogp(plan, theta)
        tol = 1e-12 if dtype == 'float64' else 2e-6
        assert lik_l + pr_l == pytest.approx(ref_l, rel=tol)
        np.testing.assert_allclose(lik_g + pr_g, ref_g, rtol=tol, atol=tol)
        k3 = plan.block_slices()['k'].start + list(plan.coords['g']).index('g3')
        assert lik_g[k3] == 0.0                                      # no row of g3 is left


def test_sharded_constants_add_up():
    df, _ = frame()
    plan = build(df, 'poisson').plan
    whole = make_kernel_plan(plan, dtype='float64')
    rows = np.arange(plan.n_rows)
    parts = [make_kernel_plan(plan, dtype='float64', rows=rows[r::3], n_rows_total=plan.n_rows) for r in range(3)]
    assert sum(p.logp_const for p in parts) == pytest.approx(whole.logp_const, rel=1e-13)


def test_argument_checks_follow_the_reference():
    df, _ = frame()
    data = api.data_components(df)
    with pytest.raises(ValueError, match='nan_param_mask and nan_data_mask'):
        api.Likelihood(dist.Normal, mu=0.0, sigma=1.0, observed=data['y_normal'])
    with pytest.raises(ValueError, match='nan_param_mask and nan_data_mask'):          # likelihood.py:33: truthy data mask
        api.Likelihood(dist.Normal, mu=0.0, sigma=1.0, observed=data['y_normal'],
                       nan_param_mask={'mu': 0.0, 'sigma': 1.0}, nan_data_mask=1.0)
    with pytest.raises(ValueError, match='Mask values for all parameters'):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', UserWarning)
            api.Likelihood(dist.Normal, mu=0.0, sigma=1.0, observed=data['y_normal'],
                           nan_param_mask={'mu': 0.0}, nan_data_mask=0)
    with pytest.warns(UserWarning, match='NaN'):
        api.Likelihood(dist.Normal, mu=0.0, sigma=1.0, observed=data['y_normal'],
                       nan_param_mask={'mu': 0.0, 'sigma': 1.0}, nan_data_mask=0)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('family', ['normal', 'bernoulli', 'poisson'])
def test_gpu_parity_with_masked_rows(family, dtype):
    df, _ = frame(n=3000, groups=40, nan_rows=tuple(range(50, 3000, 37)))
    model = build(df, family)
    f = model.logp_dlogp_function(dtype=dtype)
    for theta in theta_points(model.plan, 3, seed=11):
        logp, grad = f(theta)
        assert_parity(logp, grad, model.plan, theta, dtype)
    f.close()


def test_random_masks_agree_with_the_recorded_graph():
    """random NaN positions (including none and whole groups), all three families"""
    rng = np.random.default_rng(99)
    for trial in range(6):
        n = int(rng.integers(20, 120))
        groups = int(rng.integers(2, 6))      # (a one-member group is a twin of 'global', as in the reference)
        k = int(rng.integers(0, n // 2))
        nan_rows = tuple(sorted(rng.choice(np.arange(groups, n), size=min(k, n - groups), replace=False).tolist()))
        df, _ = frame(n=n, groups=groups, seed=100 + trial, nan_rows=nan_rows if nan_rows else (groups,))
        family = ['normal', 'bernoulli', 'poisson'][trial % 3]
        model = build(df, family)
        assert model.plan.n_rows + model.plan.n_masked_rows == n
        theta = rng.normal(0.0, 0.4, model.plan.n_params)
        logp, grad = logp_dlogp(model.plan, theta)
        ref_logp, ref_grad = logp_dlogp_autograd(model.recorded, theta)
        assert logp == pytest.approx(ref_logp, rel=1e-11)
        np.testing.assert_allclose(grad, ref_grad, rtol=1e-9, atol=1e-11)
        kp = make_kernel_plan(model.plan, dtype='float64')
        lik_l, lik_g = likelihood_from_kernel_plan(kp, theta)
        pr_l, pr_g = logp_dlogp(priors_only(model.plan), theta)
        assert lik_l + pr_l == pytest.approx(logp, rel=1e-11)
        np.testing.assert_allclose(lik_g + pr_g, grad, rtol=1e-9, atol=1e-11)
