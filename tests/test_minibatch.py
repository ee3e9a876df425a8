"""
MinibatchLikelihood (reference ``sakkara/model/composable/hierarchical/likelihood.py:65-89``,
``sakkara/model/minibatch.py``): the component structure the reference's own tests pin
(``tests/components/test_minibatch.py``), the lowering over the whole data set, and -- on a GPU -- the
mini-batch evaluator against the oracle on the rows it drew.
"""
import numpy as np
import pandas as pd
import pytest

import sakkara_b200.model as api
from sakkara_b200.model import dist as pm
from sakkara_b200.model import DistributionComponent as DC, DeterministicComponent, FunctionComponent, Likelihood, \
    MinibatchLikelihood, build, data_components
from sakkara_b200.model.components import MinibatchComponent, MinibatchDeterministic
from sakkara_b200.plan.lower import UnsupportedModelError


@pytest.fixture
def udf():
    # /root/reference/tests/fixtures (udf / xdf): 30 time steps, two groups with slopes -1 and +1
    return pd.DataFrame({'time': np.arange(30), 'u': np.random.default_rng(100).normal(0, 1, 30)})


@pytest.fixture
def xdf(udf):
    merged = pd.DataFrame({'g': ['a', 'b'], 'k': [-1, 1]}).merge(udf, how='cross')
    merged['y'] = merged['u'] * merged['k']
    merged['obs'] = np.arange(len(merged))
    return merged


def test_minibatch_component_structure(udf, xdf):
    # the assertions of /root/reference/tests/components/test_minibatch.py:11-55, with shapes read off the
    # recorded variables instead of pm.draw
    udc = data_components(udf, 'time')
    xdc = data_components(xdf)
    k = DC(pm.Normal, name='k', group='g')
    predicted = DeterministicComponent('predicted', k * udc['u'])
    mbl = MinibatchLikelihood(pm.Normal, observed=xdc['y'], batch_size=1, mu=predicted, sigma=1e-15)
    model = build(xdf, mbl, backend='b200')

    assert mbl['total_size'].variable == 60
    assert mbl.variable.shape == (60,) and mbl.variable.total_size == 60
    assert mbl['mu'].variable.shape == (1,)
    assert isinstance(mbl['mu'], MinibatchDeterministic)
    assert isinstance(mbl['mu'].component, FunctionComponent)
    assert isinstance(mbl['mu'].component.args[0], MinibatchComponent)
    assert mbl['mu'].component.args[0].component == k
    assert isinstance(mbl['mu'].component.args[1], MinibatchComponent)
    assert mbl['mu'].component.args[1].component == udc['u']
    assert mbl['observed'].variable.shape == (1,)
    assert isinstance(mbl['observed'], MinibatchComponent)
    assert mbl['observed'].component == xdc['y']
    assert predicted.variable is None                       # only the mini-batched alias was built
    assert xdc['y'].variable.shape == (60,)
    assert udc['u'].variable.shape == (30,)
    assert k.variable.shape == (2,)
    assert model.plan.minibatch == 1 and model.plan.n_rows == 60

    ll = Likelihood(pm.Normal, observed=xdc['y'], mu=predicted, sigma=1e-15)
    ll.clear()
    build(xdf, ll, backend='b200')
    assert ll['mu'].variable.shape == (2, 30)
    assert ll.variable.shape == (60,)
    assert predicted.variable.shape == (2, 30)
    assert k.variable.shape == (2,)


def test_minibatch_plan_is_the_full_plan(udf, xdf):
    def likelihood(kind, **extra):
        udc, xdc = data_components(udf, 'time'), data_components(xdf)
        k = DC(pm.Normal, name='k', group='g', sigma=DC(pm.HalfNormal, name='spread', sigma=2.0))
        return kind(pm.Normal, observed=xdc['y'], mu=DeterministicComponent('predicted', k * udc['u']),
                    sigma=DC(pm.Exponential, name='noise', lam=1.0), **extra)
    mini = build(xdf, likelihood(MinibatchLikelihood, batch_size=10), backend='b200').plan
    full = build(xdf, likelihood(Likelihood), backend='b200').plan
    assert mini.minibatch == 10 and full.minibatch is None and mini.n_rows == full.n_rows == 60
    assert [b.name for b in mini.blocks] == [b.name for b in full.blocks]
    for a, b in zip(mini.terms, full.terms):
        assert a.block == b.block and np.array_equal(a.index, b.index) and np.array_equal(a.x, b.x)
    assert np.array_equal(mini.likelihood.y, full.likelihood.y)
    rows = np.array([3, 3, 17, 59])
    sub = mini.take_rows(rows)
    assert sub.n_rows == 4 and np.array_equal(sub.likelihood.y, full.likelihood.y[rows]) and sub.minibatch is None
    assert mini.take_rows([]).n_rows == 0


def test_minibatched_components_need_a_minibatch_likelihood(udf, xdf):
    udc, xdc = data_components(udf, 'time'), data_components(xdf)
    k = DC(pm.Normal, name='k', group='g')
    mu = (k * udc['u']).to_minibatch(5, 'obs')
    with pytest.raises((UnsupportedModelError, ValueError, IndexError)):
        build(xdf, Likelihood(pm.Normal, observed=xdc['y'], mu=mu, sigma=1.0), backend='b200').plan


@pytest.mark.gpu
@pytest.mark.parametrize('dtype,tol', [('float64', 1e-10), ('float32', 1e-4)])
def test_gpu_minibatch_evaluator_matches_oracle_on_its_rows(udf, xdf, dtype, tol):
    from oracle.logp_numpy import logp_dlogp
    udc, xdc = data_components(udf, 'time'), data_components(xdf)
    k = DC(pm.Normal, name='k', group='g', sigma=DC(pm.HalfNormal, name='spread', sigma=2.0))
    mbl = MinibatchLikelihood(pm.Normal, observed=xdc['y'], batch_size=16, mu=k * udc['u'] + DC(pm.Normal, name='icpt'),
                              sigma=DC(pm.Exponential, name='noise', lam=1.0))
    model = build(xdf, mbl, backend='b200')
    plan = model.plan
    f = model.logp_dlogp_function(dtype=dtype, device=0, seed=3)
    rng = np.random.default_rng(0)
    seen = set()
    for _ in range(4):
        theta = rng.normal(0, 0.4, plan.n_params)
        logp, grad = f(theta.astype(dtype))
        rows = f.last_rows
        assert rows.shape == (16,) and rows.min() >= 0 and rows.max() < 60
        seen.add(tuple(rows))
        prior_l, prior_g = logp_dlogp(plan.take_rows([]), theta)
        sub_l, sub_g = logp_dlogp(plan.take_rows(rows), theta)
        scale = 60 / 16
        ref_l = prior_l + scale * (sub_l - prior_l)
        ref_g = prior_g + scale * (sub_g - prior_g)
        assert abs(logp - ref_l) <= tol * max(1.0, abs(ref_l))
        assert np.all(np.abs(grad - ref_g) <= tol * np.maximum(1.0, np.abs(ref_g)) * 10)
    assert len(seen) == 4                                   # a fresh draw per call
    # the estimate is unbiased: with batch = all rows drawn once each it IS the full evaluation
    theta = rng.normal(0, 0.4, plan.n_params)
    logp, grad = f(theta.astype(dtype), rows=np.arange(60))
    full_l, full_g = logp_dlogp(plan.take_rows(np.arange(60)), theta)
    assert abs(logp - full_l) <= tol * max(1.0, abs(full_l))
    assert np.all(np.abs(grad - full_g) <= tol * np.maximum(1.0, np.abs(full_g)) * 10)
    f.close()


@pytest.mark.gpu
def test_gpu_plan_without_rows_evaluates_the_priors():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import build_model, assert_parity
    from model_zoo import zoo
    for name in ('nested_logistic', 'crossed_poisson', 'readme'):
        plan = build_model(zoo()[name]).plan.take_rows([])
        from sakkara_b200.valuegrad import ValueGradFunction
        f = ValueGradFunction(plan, dtype='float64', device=0)
        theta = np.random.default_rng(1).normal(0, 0.3, plan.n_params)
        logp, grad = f(theta)
        assert_parity(logp, grad, plan, theta, 'float64')
        f.close()
