"""
The Python that composes several device plans, end to end on CPU against a stand-in device
(tests/cpu_device_double.py evaluates every ValueGradFunction from its own KernelPlan): what the public calls return
must equal the oracle of the whole model.  The GPU versions of these checks are in the ``-m gpu`` files.
"""
import numpy as np
import pytest

import pandas as pd

import cpu_device_double
import sakkara_b200.model as api
from sakkara_b200.model import dist as pm, sym as pt
from helpers import build_model, theta_points
from model_zoo import zoo
from oracle.logp_numpy import logp_dlogp

ZOO = zoo()
_DC = api.DistributionComponent


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
    approx = model.fit(n=1500, learning_rate=8e-3, random_seed=100)            # the GPU test's call
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


# ---------------------------------------------------------------------------------------------------------------
# random models: likelihood family x sigma form x a random subset of covariate / intercept / group / random-slope /
# parent-level / crossed / cell (A x B) terms in random order, random prior families for the group-level scales, NaN masks


def random_model(seed, api=api, pm=pm, pt=pt):
    """A random model on a random frame, written against the DSL modules it is given (ours by default; the reference's,
    through oracle/ref_harness.py, in tests/test_build_parity.py).  Returns ``(frame, likelihood, description)``."""
    _DC = api.DistributionComponent
    rng=np.random.default_rng(seed)
    n=int(rng.integers(40,400))
    nA=int(rng.integers(2,12)); per=int(rng.integers(2,4)); nB=int(rng.integers(2,7)); nC=int(rng.integers(2,4))
    A=rng.integers(0,nA,n); P=A//per   # P parent of A (if per>1 and distinct)
    df=pd.DataFrame({'x1':rng.standard_normal(n),'x2':rng.standard_normal(n),'A':A,'PA':P,'B':rng.integers(0,nB,n),'C':rng.integers(0,nC,n)})
    fam=rng.choice(['normal_const','normal_param','normal_group','bernoulli','poisson','negbin','studentt','gamma'])
    def sd_prior(name):
        k=rng.integers(0,4)
        if k==0: return _DC(pm.HalfNormal,name=name,sigma=1.0)
        if k==1: return _DC(pm.Exponential,name=name,lam=1.5)
        if k==2: return _DC(pm.HalfCauchy,name=name,beta=1.0)
        return 0.7
    data=api.data_components(df[['x1','x2']])
    terms=[]
    desc=[]
    if rng.random()<0.7: terms.append(_DC(pm.Normal,name='beta1')*data['x1']); desc.append('b1x1')
    if rng.random()<0.3: terms.append(_DC(pm.Normal,name='beta2',sigma=2.0)*data['x2']); desc.append('b2x2')
    if rng.random()<0.5: terms.append(_DC(pm.Normal,name='icpt')); desc.append('icpt')
    if rng.random()<0.8:
        mu = _DC(pm.Normal,name='a_mu',group='PA') if rng.random()<0.4 else (_DC(pm.Normal,name='a_mu0') if rng.random()<0.5 else 0.0)
        terms.append(_DC(pm.Normal,name='a',group='A',mu=mu,sigma=sd_prior('a_sd'))); desc.append('a[A]')
    if rng.random()<0.3: terms.append(_DC(pm.Normal,name='s',group='A',sigma=sd_prior('s_sd'))*data['x2']); desc.append('s[A]x2')
    if rng.random()<0.3: terms.append(_DC(pm.Normal,name='p',group='PA')); desc.append('p[PA]')
    if rng.random()<0.5: terms.append(_DC(pm.Normal,name='b',group='B',sigma=sd_prior('b_sd'))); desc.append('b[B]')
    if rng.random()<0.25: terms.append(_DC(pm.Normal,name='c',group='C')); desc.append('c[C]')
    if rng.random()<0.25: terms.append(_DC(pm.Normal,name='ab',group=('A','B'),sigma=sd_prior('ab_sd'))); desc.append('ab[A,B]')
    if not terms: terms.append(_DC(pm.Normal,name='icpt')); desc.append('icpt')
    order=rng.permutation(len(terms))
    eta=terms[order[0]]
    for i in order[1:]: eta=eta+terms[i]
    if rng.random()<0.3: eta = eta*0.5 + 0.1; desc.append('affine')
    e=0.3*df['x1'].values
    if fam.startswith('normal') or fam=='studentt':
        df['y']=e+0.5*rng.standard_normal(n)
        if fam=='normal_const': sigma=0.8
        elif fam=='normal_param' or (fam=='studentt' and rng.random()<0.5): sigma=_DC(pm.HalfNormal,name='noise',sigma=1.0)
        else: sigma=_DC(pm.Exponential,name='noise',group='B' if rng.random()<0.5 else 'A',lam=1.0)
        y=api.data_components(df[['y']])['y']
        kw={}
        if rng.random()<0.3:
            df.loc[rng.choice(n,3,replace=False),'y']=np.nan; y=api.data_components(df[['y']])['y']; desc.append('nan')
            kw=dict(nan_param_mask={'mu':0.2,'sigma':1.5,'nu':5.0} if fam=='studentt' else {'mu':0.2,'sigma':1.5}, nan_data_mask=0)
        if fam=='studentt': lik=api.Likelihood(pm.StudentT,nu=5.0,mu=eta,sigma=sigma,observed=y,**kw)
        else: lik=api.Likelihood(pm.Normal,mu=eta,sigma=sigma,observed=y,**kw)
    elif fam=='bernoulli':
        df['y']=(rng.random(n)<0.5).astype(float); y=api.data_components(df[['y']])['y']
        lik=api.Likelihood(pm.Bernoulli,logit_p=eta,observed=y) if rng.random()<0.5 else api.Likelihood(pm.Bernoulli,p=api.f_(pt.sigmoid)(eta),observed=y)
    elif fam=='poisson':
        df['y']=rng.poisson(np.exp(e)).astype(float); y=api.data_components(df[['y']])['y']
        if rng.random()<0.4:
            df.loc[rng.choice(n,3,replace=False),'y']=np.nan; y=api.data_components(df[['y']])['y']; desc.append('nan')
            lik=api.Likelihood(pm.Poisson,mu=api.f_(pt.exp)(eta),observed=y,nan_param_mask={'mu':1.2},nan_data_mask=0)
        else:
            lik=api.Likelihood(pm.Poisson,mu=api.f_(pt.exp)(eta),observed=y)
    elif fam=='negbin':
        df['y']=rng.poisson(np.exp(e)).astype(float); y=api.data_components(df[['y']])['y']
        lik=api.Likelihood(pm.NegativeBinomial,mu=api.f_(pt.exp)(eta),alpha=1.7,observed=y)
    else:
        df['y']=rng.gamma(2.0,np.exp(e)/2.0); y=api.data_components(df[['y']])['y']
        lik=api.Likelihood(pm.Gamma,alpha=2.0,beta=2.0/api.f_(pt.exp)(eta),observed=y)
    return df, lik, fam + ' ' + '+'.join(desc)


@pytest.mark.parametrize('seed', range(0, 80))
def test_random_models_equal_autograd_over_the_recorded_graph(seed):
    """Whatever evaluator model.logp_dlogp_function() picks for a random model -- one plan, a sum of plans, folded
    terms, split sigmas -- equals torch autograd over the graph build() recorded (oracle/autograd_check.py shares no
    code with the lowering, the planner or the composition).  Models outside the supported class must say so."""
    import warnings
    from oracle.autograd_check import logp_dlogp_autograd
    from sakkara_b200.plan.lower import UnsupportedModelError
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            df, lik, description = random_model(seed)
            model = api.build(df, lik, backend='b200')
            f = model.logp_dlogp_function(max_batch=int(seed % 2) * 3 + 1)
            theta = np.random.default_rng(seed).normal(0, 0.4, model.plan.n_params)
            logp, grad = f(theta)                  # (plans are made when first needed: they may refuse only here)
    except (UnsupportedModelError, NotImplementedError) as err:
        pytest.skip(f'outside the supported class: {err}')
    except ValueError as err:
        if 'minimal' in str(err):
            pytest.skip('the random frame made a parent column a twin (the reference raises the same)')
        raise
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref_logp, ref_grad = logp_dlogp_autograd(model.recorded, theta)
    assert abs(logp - ref_logp) <= 1e-9 * max(1.0, abs(ref_logp)), description
    assert np.abs(grad - ref_grad).max() <= 1e-8 * max(1.0, np.abs(ref_grad).max()), description


@pytest.mark.parametrize('name', ['readme', 'crossed_poisson', 'group_sigma', 'crossed_gamma'])
def test_pymc_model_over_the_real_evaluators(name):
    """``build(df, likelihood)`` as the reference's user writes it, with PyMC importable (tests/pymc_stub.py evaluates
    the PyMC / PyTensor interfaces the glue uses): the returned ``pm.Model``'s logp_dlogp_function -- value variables
    in PyMC's raveled order, Jacobians removed from the Potential and re-added by the transforms -- over the repo's own
    evaluator classes (here on the stand-in device) equals the oracle."""
    import pymc_stub
    import sakkara_b200.model as api_
    pymc_stub.install()
    try:
        df, make = ZOO[name]
        model = api_.build(df, make(api_, api_.dist, api_.sym))
        assert type(model).__name__ == 'Model' and hasattr(model, 'b200')       # a pm.Model, not the B200Model
        f = model.logp_dlogp_function()
        plan = model.b200.plan
        for theta in theta_points(plan, 2, seed=4, scale=0.3):
            logp, grad = f(theta)
            ref_logp, ref_grad = logp_dlogp(plan, theta)
            assert abs(logp - ref_logp) <= 1e-10 * max(1.0, abs(ref_logp))
            assert np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)
    finally:
        pymc_stub.uninstall()


def test_pymc_model_over_a_sum_of_plans():
    import pymc_stub
    from test_valuegrad import _crossed3_model
    pymc_stub.install()
    try:
        b200, _ = _crossed3_model(n=500)
        from sakkara_b200.pytensor_op import to_pymc_model
        model = to_pymc_model(b200)
        assert type(model.b200_value_grad).__name__ == 'PartitionedValueGrad'
        f = model.logp_dlogp_function()
        theta = theta_points(b200.plan, 2, seed=9, scale=0.3)[1]
        logp, grad = f(theta)
        ref_logp, ref_grad = logp_dlogp(b200.plan, theta)
        assert abs(logp - ref_logp) <= 1e-10 * abs(ref_logp) and np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)
    finally:
        pymc_stub.uninstall()
