"""
The Python that composes several device plans, end to end on CPU against a stand-in device
(tests/cpu_device_double.py evaluates every ValueGradFunction from its own KernelPlan): what the public calls return
must equal the oracle of the whole model.  The GPU versions of these checks are in the ``-m gpu`` files.
"""
import numpy as np
import pytest

import cpu_device_double
from helpers import build_model, theta_points
from model_zoo import zoo
from oracle.logp_numpy import logp_dlogp

ZOO = zoo()


@pytest.fixture(autouse=True)
def device_double(monkeypatch):
    cpu_device_double.install(monkeypatch)


@pytest.mark.parametrize('name', sorted(ZOO))
def test_every_zoo_model_through_the_public_call(name):
    """model.logp_dlogp_function() -- whichever evaluator class it picks -- equals the oracle."""
    model = build_model(ZOO[name])
    f = model.logp_dlogp_function()
    for theta in theta_points(model.plan, 3, seed=21):
        logp, grad = f(theta)
        ref_logp, ref_grad = logp_dlogp(model.plan, theta)
        assert abs(logp - ref_logp) <= 1e-10 * max(1.0, abs(ref_logp))
        assert np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)
    f.close()


@pytest.mark.parametrize('four', [False, True])
def test_crossed_factors_beyond_two(four):
    from test_valuegrad import _crossed3_model
    model, _ = _crossed3_model(four=four)
    f = model.logp_dlogp_function()
    assert type(f).__name__ == 'PartitionedValueGrad' and len(f.parts) == (12 if four else 4)
    thetas = theta_points(model.plan, 3, seed=2, scale=0.3)
    for theta in thetas:
        logp, grad = f(theta)
        ref_logp, ref_grad = logp_dlogp(model.plan, theta)
        assert abs(logp - ref_logp) <= 1e-10 * abs(ref_logp) and np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)
    # a batch: the parts evaluate all points, the aliasing acts per point
    g = model.logp_dlogp_function(max_batch=4)
    logp, grad = g(thetas)
    assert logp.shape == (3,) and grad.shape == thetas.shape
    assert np.allclose(grad[2], logp_dlogp(model.plan, thetas[2])[1], rtol=1e-9, atol=1e-9)
    out = np.empty_like(thetas[0])
    assert f(thetas[1], grad_out=out)[1] is out


def test_oversized_secondary_level_through_the_model():
    from test_valuegrad import _crossed3_model      # (its frame: A 40 groups x B 9 groups)
    import pandas as pd
    import sakkara_b200.model as api
    from sakkara_b200.model import dist as pm, sym as pt
    rng = np.random.default_rng(0)
    n = 3000
    df = pd.DataFrame({'x': rng.standard_normal(n), 'A': rng.integers(0, 700, n), 'B': rng.integers(0, 600, n)})
    df['y'] = rng.poisson(np.exp(0.2 * df['x'])).astype(float)
    DC = api.DistributionComponent
    data = api.data_components(df[['x', 'y']])
    eta = DC(pm.Normal, name='beta') * data['x'] + DC(pm.Normal, name='a', group='A') + DC(pm.Normal, name='b', group='B')
    model = api.build(df, api.Likelihood(pm.Poisson, mu=api.f_(pt.exp)(eta), observed=data['y']), backend='b200')
    f = model.logp_dlogp_function(max_batch=2)          # batch tile of 4, fp64: 512 secondary groups per plan
    assert type(f).__name__ == 'PartitionedValueGrad' and len(f.parts) == 2
    assert max(p.kernel_plan.n_secondary for p in f.parts) <= 512
    theta = theta_points(model.plan, 2, seed=3, scale=0.2)[1]
    logp, grad = f(theta)
    ref_logp, ref_grad = logp_dlogp(model.plan, theta)
    assert abs(logp - ref_logp) <= 1e-10 * abs(ref_logp) and np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)
    # one point at a time the level fits one plan (2,048 groups in fp64)
    assert type(model.logp_dlogp_function()).__name__ == 'ValueGradFunction'


def test_minibatch_evaluator_and_fit():
    """MinibatchValueGrad composes priors + n / batch x the likelihood of the drawn rows; model.fit() runs on it."""
    from test_variational import _reference_minibatch_model
    model = _reference_minibatch_model()
    plan = model.plan
    f = model.logp_dlogp_function(seed=5)
    theta = np.array([0.3, -0.2])
    logp, grad = f(theta)
    rows = f.last_rows
    prior_l, prior_g = logp_dlogp(plan.take_rows([]), theta)
    sub_l, sub_g = logp_dlogp(plan.take_rows(rows), theta)
    assert logp == pytest.approx(prior_l + 6.0 * (sub_l - prior_l), rel=1e-10)
    assert np.allclose(grad, prior_g + 6.0 * (sub_g - prior_g), rtol=1e-10)
    f.close()
    approx = model.fit(n=4000, learning_rate=3e-3, random_seed=100)            # the GPU test's call
    k = approx.constrained(plan, approx.sample(10000, seed=100))['k']
    assert abs(k[:, 0].mean() + 1.0) < 2e-2 and abs(k[:, 1].mean() - 1.0) < 2e-2


def test_host_resident_nuts_on_a_sum_of_plans():
    """sample_nuts on a model that is not one device plan: chains in host memory, batched host calls."""
    from sakkara_b200.sampling import sample_nuts
    from test_valuegrad import _crossed3_model
    model, df = _crossed3_model(n=600)
    res = sample_nuts(model, chains=4, draws=30, tune=40, max_depth=5, seed=1)
    assert res.draws.shape == (4, 30, model.plan.n_params) and np.isfinite(res.logp).all()
    lp, _ = logp_dlogp(model.plan, res.draws[1, 7])
    assert res.logp[1, 7] == pytest.approx(lp, rel=1e-9)
    assert res.stats['depth'].max() <= 5 and res.accept_rate.mean() > 0.3


def test_nan_masked_rows_next_to_sigma_parameters():
    """A NaN-masked Normal likelihood whose sigma is a GroupComponent of a constant and two parameters: the evaluator is
    a sum of plans (SplitSigmaValueGrad), and the constant of the masked rows is added once (it used to be dropped)."""
    import warnings
    import pandas as pd
    import sakkara_b200.model as api
    from oracle.autograd_check import logp_dlogp_autograd
    from sakkara_b200.model import dist as pm
    rng = np.random.default_rng(0)
    n = 200
    g = rng.integers(0, 3, n)
    x = rng.normal(size=n)
    df = pd.DataFrame({'x': x, 'group': g, 'y': 0.8 * x - 0.2 + np.array([0.1, 0.6, 1.4])[g] * rng.normal(size=n)})
    df.loc[[3, 17, 50], 'y'] = np.nan
    DC, GC = api.DistributionComponent, api.GroupComponent
    data = api.data_components(df)
    R = GC(name='R', group='group', membercomponents={0: 1e-1, 1: DC(pm.Exponential, name='R_1', lam=10),
                                                       2: DC(pm.HalfNormal, name='R_2', sigma=2.0)})
    k = DC(pm.Normal, name='k', group='group', sigma=DC(pm.HalfNormal, name='k_sd', sigma=1.0))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', UserWarning)
        lik = api.Likelihood(pm.Normal, mu=k * data['x'] + DC(pm.Normal, name='icpt'), sigma=R, observed=data['y'],
                             nan_param_mask={'mu': 0.5, 'sigma': 2.0}, nan_data_mask=0)
        model = api.build(df, lik, backend='b200')
    plan = model.plan
    assert plan.n_masked_rows == 3 and plan.logp_const != 0.0
    f = model.logp_dlogp_function()
    assert type(f).__name__ == 'SplitSigmaValueGrad'
    theta = np.random.default_rng(1).normal(0, 0.3, plan.n_params)
    logp, grad = f(theta)
    ref_logp, ref_grad = logp_dlogp_autograd(model.recorded, theta)            # all rows, where(mask, fill, value)
    assert logp == pytest.approx(ref_logp, rel=1e-12)
    assert np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('name', ['group_sigma', 'studentt_group_sigma', 'studentt_one_sigma', 'studentt_const_sigma',
                                  'row_sigma', 'readme'])
def test_per_element_option_never_changes_the_result(name):
    """``per_element=True`` (one standard plan per sigma parameter: the cross-check of the log-sigma kernels) applies
    to Normal sigmas; StudentT sigma parameters keep the log-sigma plan (there is no standard StudentT plan with a
    sigma element -- it used to be evaluated with sigma = 1), other models ignore the option."""
    model = build_model(ZOO[name])
    f = model.logp_dlogp_function(per_element=True)
    theta = theta_points(model.plan, 2, seed=1)[1]
    logp, grad = f(theta)
    ref_logp, ref_grad = logp_dlogp(model.plan, theta)
    assert abs(logp - ref_logp) <= 1e-10 * abs(ref_logp) and np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)
