"""Behaviour of the component DSL that model code written for the reference relies on."""
import numpy as np
import pandas as pd
import pytest

import sakkara_b200.model as api
from sakkara_b200.model import dist as pm, sym
from sakkara_b200.model import (DataComponent, DistributionComponent as DC, FunctionComponent, Likelihood,
                                UnrepeatableComponent, build, data_components, f_)
from sakkara_b200.plan.lower import UnsupportedModelError


@pytest.fixture
def df():
    # the frame of /root/reference/tests/workflows/test_linreg.py:14-29
    rng = np.random.default_rng(100)
    frame = pd.DataFrame({'g1': np.repeat(['a', 'b'], 4), 'g2': np.repeat(['a1', 'a2', 'b1', 'b2'], 2),
                          'x': rng.normal(0, 1, 8), 'y': 0.0})
    for g2, (icpt, k) in {'a1': (1, 1), 'a2': (.8, 1), 'b1': (.6, .4), 'b2': (.4, .4)}.items():
        rows = frame['g2'] == g2
        frame.loc[rows, 'y'] = icpt + frame.loc[rows, 'x'] * k
    return frame


def linreg(df):
    coeff = DC(pm.Normal, name='coeff', group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=1e-10))
    intercept = DC(pm.Normal, name='intercept', group='g2',
                   mu=DC(pm.Normal, group='g1', mu=DC(pm.Normal), sigma=DC(pm.HalfNormal, sigma=1e-5)),
                   sigma=DC(pm.HalfNormal, sigma=1e-10))
    data = data_components(df)
    return Likelihood(pm.Normal, mu=coeff * data['x'] + intercept, sigma=1e-15, observed=data['y'])


def test_component_tree_and_shapes(df):
    # mirrors tests/workflows/test_linreg.py:60-98 of the reference
    lik = linreg(df)
    assert isinstance(lik['mu'], FunctionComponent) and isinstance(lik['mu'].args[0], FunctionComponent)
    assert isinstance(lik['mu'].args[0].args[0]['sigma']['sigma'], UnrepeatableComponent)
    assert isinstance(lik['mu'].args[0].args[1], DataComponent)
    assert isinstance(lik['sigma'], UnrepeatableComponent)
    model = build(df, lik)
    assert lik.variable.shape == (8,)
    assert lik['mu'].variable.shape == (8,) and lik['mu'].args[0].args[1].variable.shape == (8,)
    assert lik['mu'].args[0].args[0].variable.shape == (2,)
    assert lik['mu'].args[0].args[0]['mu'].variable.shape == (1,)
    assert lik['mu'].args[1].variable.shape == (4,)
    assert lik['mu'].args[1]['mu'].variable.shape == (2,)
    assert lik['mu'].args[1]['mu']['mu'].variable.shape == (1,)
    assert model.plan.n_params == 1 + 1 + 2 + 1 + 1 + 2 + 1 + 4
    assert model.plan.likelihood.sigma_const == 1e-15


def test_auto_naming(df):
    lik = linreg(df)
    build(df, lik)
    assert lik['mu'].args[0].args[0]['mu'].get_name() == 'mu_coeff'
    assert lik['mu'].args[1]['mu']['sigma'].get_name() == 'sigma_mu_intercept'
    assert lik.get_name() == 'likelihood'


def test_rebuild_needs_clear(df):
    # reference tests/workflows/test_linreg.py:101-113: a built tree cannot be built again until clear()
    lik = linreg(df)
    build(df, lik)
    saved = lik.variable
    with pytest.raises(ValueError):
        build(df, lik)          # data components re-register their names
    lik.clear()
    build(df, lik)
    assert lik.variable is not saved


def test_name_clash_of_the_readme(df):
    # README.md:25 names the coefficient 'x' while column 'x' is registered as data (quirk D.1)
    data = data_components(df)
    coeff = DC(pm.Normal, name='x', group='g1')
    with pytest.raises(ValueError, match='already exists'):
        build(df, Likelihood(pm.Normal, mu=coeff * data['x'], sigma=1.0, observed=data['y']))


def test_non_minimal_groups_rejected(df):
    data = data_components(df)
    bad = DC(pm.Normal, name='b', group=('g1', 'g2'))
    with pytest.raises(ValueError, match='minimal'):
        build(df, Likelihood(pm.Normal, mu=bad, sigma=1.0, observed=data['y']))


def test_operators_and_scalars(df):
    data = data_components(df)
    k = DC(pm.Normal, name='k', group='g1')
    eta = -(2.0 * k) * data['x'] / 4.0 + 1.5 - k
    model = build(df, Likelihood(pm.Normal, mu=eta, sigma=2.0, observed=data['y']))
    plan = model.plan
    assert len(plan.terms) == 2 and all(plan.blocks[t.block].name == 'k' for t in plan.terms)
    assert np.allclose(plan.terms[0].x, -0.5 * df['x'].values)
    assert np.allclose(plan.terms[1].x, -1.0)
    assert np.allclose(plan.likelihood.offset, 1.5)


def test_unsupported_structures_raise(df):
    data = data_components(df)
    k = DC(pm.Normal, name='k', group='g1')
    j = DC(pm.Normal, name='j', group='g2')
    with pytest.raises(UnsupportedModelError):          # not affine in theta
        build(df, Likelihood(pm.Normal, mu=k * j, sigma=1.0, observed=data['y']))
    for c in (k, j, *data.values()):
        c.clear()
    with pytest.raises(UnsupportedModelError):
        build(df, Likelihood(pm.Normal, mu=k ** 2, sigma=1.0, observed=data['y']))
    for c in (k, *data.values()):
        c.clear()
    with pytest.raises(NotImplementedError):            # unknown distribution
        class StudentT:
            pass
        build(df, Likelihood(pm.Normal, mu=DC(StudentT, name='t'), sigma=1.0, observed=data['y']))
    with pytest.raises(NotImplementedError):
        api.GroupComponent('g1')
    with pytest.raises(NotImplementedError):
        api.MinibatchLikelihood(pm.Normal, data['y'], 4)


def test_pymc_generators_are_accepted_by_name(df):
    class Normal:               # stands for pm.Normal: resolved by class name
        pass

    class HalfNormal:
        pass
    data = data_components(df)
    k = DC(Normal, name='k', group='g1', sigma=DC(HalfNormal, sigma=2.0))
    plan = build(df, Likelihood(Normal, mu=k * data['x'], sigma=1.0, observed=data['y'])).plan
    assert [b.family for b in plan.blocks] == ['HalfNormal', 'Normal']


def test_poisson_needs_exp_link(df):
    frame = df.assign(y=np.arange(8))
    data = data_components(frame)
    k = DC(pm.Normal, name='k', group='g1')
    plan = build(frame, Likelihood(pm.Poisson, mu=f_(sym.exp)(k * data['x']), observed=data['y'])).plan
    assert plan.likelihood.family == 'Poisson'
    for c in (k, *data.values()):
        c.clear()
    with pytest.raises(UnsupportedModelError):
        build(frame, Likelihood(pm.Poisson, mu=k * data['x'], observed=data['y']))
