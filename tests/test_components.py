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
    for c in (k, *data.values()):
        c.clear()
    model = build(df, Likelihood(pm.Normal, mu=k ** 2, sigma=1.0, observed=data['y']), strict=False)
    assert not model.is_lowered and model['k'].shape == (2,)      # the recorded graph is there ...
    with pytest.raises(UnsupportedModelError):                     # ... its evaluation is not
        model.logp_dlogp_function()
    for c in (k, *data.values()):
        c.clear()
    with pytest.raises(NotImplementedError):            # unknown distribution
        class Weibull:
            pass
        build(df, Likelihood(pm.Normal, mu=DC(Weibull, name='t'), sigma=1.0, observed=data['y']))


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


# ---------------------------------------------------------------------------------------------
# GroupComponent: the cases of /root/reference/tests/components/test_group_component.py.  The reference draws
# from degenerate Uniform variables to read the member-wise values; here the members are constants (or
# Normal variables evaluated at a theta that numbers them) and the recorded graph is evaluated.
# ---------------------------------------------------------------------------------------------
@pytest.fixture
def simple_df():
    # /root/reference/tests/fixtures/simple_df.py:6-16
    return pd.DataFrame({
        'global': 'global',
        'building': np.repeat(list('ab'), 10),
        'sensor': np.repeat(list('stuv'), 5),
        'time': np.tile(pd.date_range('2018-01-01 00:00:00', tz='utc', periods=5, freq='h'), 4),
        'time_2': pd.date_range('2020-01-01 00:00:00', tz='utc', periods=20, freq='h'),
        'obs': np.arange(20)})


def _values(component, model, theta=None):
    """The component's variable evaluated at theta (free variables in creation order)."""
    from oracle.autograd_check import _evaluate
    import torch
    env, offset = {}, 0
    for rv in model.recorded.free_RVs:
        env[id(rv)] = torch.as_tensor(theta[offset:offset + rv.size], dtype=torch.float64).reshape(rv.shape)
        offset += rv.size
    return np.asarray(_evaluate(component.variable, env))


def test_group_component_fixed_values(simple_df):
    gc = api.GroupComponent('sensor', 'gc', {k: i for i, k in enumerate('stuv')})
    x = DC(pm.Normal, name='x', mu=gc, sigma=gc + 1.0, group='sensor')
    model = build(simple_df, x, backend='b200')
    assert [str(g) for g in gc.representation.get_groups()] == ['sensor']
    assert gc.variable.shape == (4,)
    assert _values(gc, model, np.zeros(4)).tolist() == [0, 1, 2, 3]
    assert not model.is_lowered                                    # no likelihood: the recorded graph only
    with pytest.raises(UnsupportedModelError):
        model.plan


def test_group_component_unrelated_groups(simple_df):
    members = {k: DC(pm.Normal, name=f'm{k}', group='time') for k in 'stuv'}
    gc = api.GroupComponent('sensor', 'gc', members)
    x = gc + DC(pm.Normal, 'x', group='building')
    model = build(simple_df, x, backend='b200')
    assert [str(g) for g in gc.representation.get_groups()] == ['sensor', 'time']
    assert gc.variable.shape == (4, 5)
    theta = np.concatenate([np.full(5, float(i)) for i in range(4)] + [np.zeros(2)])
    got = _values(gc, model, theta)
    assert all(got[i, j] == i for i in range(4) for j in range(5))


def test_group_component_mixed_groups(simple_df):
    gc = api.GroupComponent('sensor', 'gc', {k: DC(pm.Normal, name=f'm{k}', group='time') for k in 'st'})
    gc.add('u', DC(pm.Normal, name='mu', group='global'))
    gc.add('v', 3)
    model = build(simple_df, gc, backend='b200')
    assert [str(g) for g in gc.representation.get_groups()] == ['sensor', 'time']
    assert gc.variable.shape == (4, 5)
    theta = np.concatenate([np.zeros(5), np.ones(5), [2.0]])
    got = _values(gc, model, theta)
    assert all(got[i, j] == i for i in range(4) for j in range(5))
    assert gc['u'].get_name() == 'mu' and isinstance(gc['v'], DataComponent)


def test_group_component_parent_child(simple_df):
    with pytest.raises(ValueError):
        gc = api.GroupComponent('building', 'gc', {k: DC(pm.Normal, name=f'c{k}', group='sensor') for k in 'ab'})
        build(simple_df, gc, backend='b200')
    c = DC(pm.Normal, name='c', group='building')
    gc = api.GroupComponent('sensor', 'gc', {k: c for k in 'stuv'})
    model = build(simple_df, gc, backend='b200')
    assert _values(gc, model, np.array([0.0, 1.0])).tolist() == [0, 0, 1, 1]


def test_group_component_needs_every_member(simple_df):
    with pytest.raises(ValueError):
        gc = api.GroupComponent('sensor', 'gc', {k: DC(pm.Normal, name=f'm{k}', group='time') for k in 'st'})
        build(simple_df, gc, backend='b200')


def test_group_component_lowers_to_one_term(simple_df):
    # different blocks, a scalar variable and a constant in different rows: one gather over theta + an offset
    frame = simple_df.assign(x=np.linspace(-1, 1, 20), y=np.cos(np.arange(20.0)))
    data = data_components(frame[['x', 'y']])
    gc = api.GroupComponent('sensor', 'gc', {'s': DC(pm.Normal, name='ms', group='time'),
                                            't': DC(pm.Normal, name='mt', group='time'),
                                            'u': DC(pm.Normal, name='mu'), 'v': 3.0})
    plan = build(frame, Likelihood(pm.Normal, mu=gc * data['x'], sigma=1.0, observed=data['y']), backend='b200').plan
    assert len(plan.terms) == 1 and plan.terms[0].block == -1
    t = plan.terms[0]
    assert t.index[:5].tolist() == [0, 1, 2, 3, 4] and t.index[5:10].tolist() == [5, 6, 7, 8, 9]
    assert (t.index[10:15] == 10).all()
    assert np.array_equal(t.x[:15], frame['x'].values[:15]) and (t.x[15:] == 0).all()
    assert np.allclose(plan.likelihood.offset[15:], 3.0 * frame['x'].values[15:]) and (plan.likelihood.offset[:15] == 0).all()


def test_gamma_glm_spellings_and_rejections():
    """Gamma(alpha, beta = alpha / exp(eta)): the three accepted spellings give the same terms (a constant other than
    alpha becomes an offset of eta); anything else raises."""
    import pandas as pd
    rng = np.random.default_rng(0)
    frame = pd.DataFrame({'g': rng.integers(0, 3, 50), 'x': rng.normal(size=50), 'y': rng.gamma(2.0, 1.0, 50)})

    def plan_of(beta_of, alpha=2.0, y='y'):
        data = api.data_components(frame)
        k = api.DistributionComponent(pm.Normal, name='k', group='g')
        eta = k * data['x'] + api.DistributionComponent(pm.Normal, name='icpt')
        return api.build(frame, api.Likelihood(pm.Gamma, alpha=alpha, beta=beta_of(eta), observed=data[y]),
                         backend='b200').plan

    exp = api.f_(sym.exp)
    over = plan_of(lambda eta: 2.0 / exp(eta))
    times = plan_of(lambda eta: exp(-eta) * 2.0)
    shifted = plan_of(lambda eta: 0.5 / exp(eta))
    assert over.likelihood.family == 'Gamma' and over.likelihood.sigma_const == 2.0 and over.likelihood.offset is None
    for a, b in zip(over.terms, times.terms):
        assert a.block == b.block and np.array_equal(a.index, b.index)
        assert (a.x is None and b.x is None) or np.allclose(a.x, b.x)
    assert np.allclose(shifted.likelihood.offset, np.log(2.0 / 0.5))
    with pytest.raises(UnsupportedModelError):
        plan_of(lambda eta: exp(eta) + 1.0)                           # not a rate of the log-link GLM
    with pytest.raises(UnsupportedModelError):
        plan_of(lambda eta: 2.0 / eta)                                # no link
    with pytest.raises(UnsupportedModelError):
        plan_of(lambda eta: 2.0 / exp(eta), alpha=api.DistributionComponent(pm.Exponential, name='shape', lam=1.0))
    frame['y0'] = frame['y'].where(frame.index != 3, 0.0)
    with pytest.raises(ValueError):
        plan_of(lambda eta: 2.0 / exp(eta), y='y0')                   # observations must be positive
