"""
Models that take several plans (valuegrad.PartitionedValueGrad) on the GPU: a crossed model whose secondary factor is
split by ranges of groups, and a model with three crossed factors split by the smallest one.
(The composition is pinned on CPU in tests/test_valuegrad.py from the parts' kernel plans; this file was written
after the round's GPU budget was spent and sorts last so that it cannot mask another file's result under ``-x``.)
"""
import numpy as np
import pytest

from helpers import assert_parity
from sakkara_b200 import synth
from sakkara_b200.valuegrad import ChunkedSecondaryValueGrad

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]      # (pytest-timeout, where installed: a stuck test fails alone)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_secondary_level_split_over_three_plans(dtype):
    cols, _ = synth.poisson_columns(120_000, 3000, 300, seed=8)
    plan = synth.poisson_plan(cols)
    f = ChunkedSecondaryValueGrad(plan, dtype=dtype, chunk_groups=100)
    assert len(f.parts) == 3
    theta = synth.start_point(plan, 1, seed=1, scale=0.3, dtype=dtype)[0]
    logp, grad = f(theta)
    assert_parity(logp, grad, plan, theta, dtype)
    again = f(theta)
    assert again[0] == logp and np.array_equal(again[1], grad)
    f.close()


def test_secondary_level_too_large_for_one_plan():
    """The shape tests/test_gpu_shapes.py::test_secondary_level_too_large_is_rejected shows one plan cannot hold."""
    rng = np.random.default_rng(0)
    n = 400_000
    cols = {'x': rng.standard_normal(n), 'y': rng.poisson(1.0, n).astype(float),
            'A': rng.integers(0, 150_000, n), 'B': rng.integers(0, 120_000, n)}
    plan = synth.poisson_plan(cols)
    f = ChunkedSecondaryValueGrad(plan, dtype='float64')
    assert len(f.parts) > 1 and max(p.kernel_plan.n_secondary for p in f.parts) <= f.chunk_groups
    theta = synth.start_point(plan, 1, seed=2, scale=0.2, dtype='float64')[0]
    logp, grad = f(theta)
    assert_parity(logp, grad, plan, theta, 'float64')
    f.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_three_crossed_factors(dtype):
    from helpers import theta_points
    from test_valuegrad import _crossed3_model
    model, _ = _crossed3_model()
    f = model.logp_dlogp_function(dtype=dtype)
    assert type(f).__name__ == 'PartitionedValueGrad' and len(f.parts) == 4
    for theta in theta_points(model.plan, 3, seed=4, scale=0.3):
        logp, grad = f(theta)
        assert_parity(logp, grad, model.plan, theta, dtype)
    f.close()


def test_stats_name_the_kernels_of_the_last_evaluation():
    """skk_stats_t.kernel_variant: 2 when the chain-per-lane kernels evaluated the batch, 1 for the row kernel."""
    from sakkara_b200.valuegrad import ValueGradFunction
    cols, _ = synth.linreg_columns(20_000, 50, seed=3)
    plan = synth.linreg_plan(cols)
    f = ValueGradFunction(plan, dtype='float64', max_batch=40)
    theta = synth.start_point(plan, 40, seed=2, scale=0.3, dtype='float64')
    f(theta)
    assert f.stats()['kernel_variant'] == 2 and f.stats()['row_kernel_launches'] == 1
    f(theta[:3])
    assert f.stats()['kernel_variant'] == 1
    f.close()


def test_advi_on_a_minibatched_model():
    """model.fit(): mean-field ADVI whose every step evaluates a fresh mini-batch of rows with the fused kernels."""
    import pandas as pd
    import sakkara_b200.model as api
    from sakkara_b200.model import dist as pm
    rng = np.random.default_rng(1)
    n = 3000
    group = rng.integers(0, 2, n)
    x = rng.standard_normal(n)
    df = pd.DataFrame({'x': x, 'g': group, 'y': np.array([-1.0, 1.0])[group] * x + 0.2 + 0.4 * rng.standard_normal(n)})
    data = api.data_components(df)
    DC = api.DistributionComponent
    k = DC(pm.Normal, name='k', group='g')
    lik = api.MinibatchLikelihood(pm.Normal, observed=data['y'], batch_size=256, mu=k * data['x'] + DC(pm.Normal, name='icpt'),
                                  sigma=DC(pm.Exponential, name='noise', lam=1.0))
    model = api.build(df, lik, backend='b200')
    assert model.plan.minibatch == 256
    approx = model.fit(n=800, learning_rate=0.03, random_seed=3, start_sigma=0.1)
    post = approx.constrained(model.plan, approx.sample(2000))
    order = list(model.plan.coords['g'])
    assert np.allclose(post['k'].mean(0), np.array([-1.0, 1.0])[order], atol=0.15)
    assert abs(post['icpt'].mean() - 0.2) < 0.1 and abs(post['noise'].mean() - 0.4) < 0.1
    assert approx.elbo[-100:].mean() > approx.elbo[:100].mean()


def test_four_crossed_factors_next_to_a_covariate():
    """Two peeled factors are evaluated as one global term of a part (valuegrad.merge_constant_terms)."""
    from helpers import theta_points
    from test_valuegrad import _crossed3_model
    model, _ = _crossed3_model(four=True)
    f = model.logp_dlogp_function(dtype='float64')
    assert type(f).__name__ == 'PartitionedValueGrad' and len(f.parts) == 12 and all(a is not None for a in f.aliases)
    for theta in theta_points(model.plan, 2, seed=6, scale=0.3):
        logp, grad = f(theta)
        assert_parity(logp, grad, model.plan, theta, 'float64')
    f.close()


def test_minibatch_vi_workflow_of_the_reference():
    """/root/reference/tests/workflows/test_minibatch_vi.py:8-24 with model.fit() on the mini-batched evaluator (fewer,
    larger steps than the reference's 10,000 x 1e-3: every step creates a plan for its ten rows)."""
    from test_variational import _reference_minibatch_model
    model = _reference_minibatch_model()
    approx = model.fit(n=1500, learning_rate=8e-3, random_seed=100)
    k = approx.constrained(model.plan, approx.sample(10000, seed=100))['k']
    assert abs(k[:, 0].mean() + 1.0) < 2e-2 and abs(k[:, 1].mean() - 1.0) < 2e-2


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_more_terms_than_one_level_takes(dtype):
    """Two covariates + intercept + a[A] + b[B] (three global terms: the planner moves one to the primary level), and
    the seven-term model of tests/test_valuegrad.py::test_terms_beyond_the_kernels_limits_are_folded."""
    import pandas as pd
    import sakkara_b200.model as api
    from helpers import theta_points
    from sakkara_b200.model import dist as pm
    rng = np.random.default_rng(3)
    n = 20_000
    A = rng.integers(0, 120, n)
    df = pd.DataFrame({'x1': rng.standard_normal(n), 'x2': rng.standard_normal(n), 'A': A, 'PA': A // 3,
                       'B': rng.integers(0, 5, n)})
    df['y'] = 0.4 * df['x1'] - 0.2 * df['x2'] + 0.1 * (df['PA'] % 3) + 0.5 * rng.standard_normal(n)
    DC = api.DistributionComponent

    def model_of(seven):
        data = api.data_components(df[['x1', 'x2', 'y']])
        eta = (DC(pm.Normal, name='beta1') * data['x1'] + DC(pm.Normal, name='beta2') * data['x2']
               + DC(pm.Normal, name='icpt') + DC(pm.Normal, name='a', group='A', sigma=DC(pm.HalfNormal, name='a_sd'))
               + DC(pm.Normal, name='b', group='B'))
        if seven:
            eta = eta + DC(pm.Normal, name='s', group='A') * data['x2'] + DC(pm.Normal, name='p', group='PA')
        return api.build(df, api.Likelihood(pm.Normal, mu=eta, sigma=DC(pm.Exponential, name='noise', lam=1.0),
                                            observed=data['y']), backend='b200')

    for seven, kind in ((False, 'ValueGradFunction'), (True, 'PartitionedValueGrad')):
        model = model_of(seven)
        f = model.logp_dlogp_function(dtype=dtype)
        assert type(f).__name__ == kind
        for theta in theta_points(model.plan, 2, seed=8, scale=0.3):
            logp, grad = f(theta)
            assert_parity(logp, grad, model.plan, theta, dtype)
        f.close()


@pytest.mark.parametrize('seed', [0, 3, 5, 6, 9, 14, 16, 22, 24, 37, 41, 49, 57, 63])
def test_random_models_on_the_device(seed):
    """Models of tests/test_composition_cpu.py's generator that one plan did not take before the planner moved terms
    between levels / the evaluator folded terms / summed plans (and a few ordinary ones), on the real device."""
    import warnings
    import sakkara_b200.model as api
    from sakkara_b200.plan.lower import UnsupportedModelError
    from test_composition_cpu import random_model
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            df, lik, description = random_model(seed)
            model = api.build(df, lik, backend='b200')
            f = model.logp_dlogp_function(dtype='float64')
            theta = np.random.default_rng(seed).normal(0, 0.4, model.plan.n_params)
            logp, grad = f(theta)
    except (UnsupportedModelError, NotImplementedError) as err:
        pytest.skip(f'outside the supported class: {err}')
    except ValueError as err:
        if 'minimal' in str(err):
            pytest.skip('the random frame made a parent column a twin')
        raise
    assert_parity(logp, grad, model.plan, theta, 'float64')
    f.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_intercept_and_two_random_slopes_by_one_group(dtype):
    """Three primary-level terms: the planner moves one to the free secondary level (tests/test_kernel_plan.py)."""
    from sakkara_b200.valuegrad import ValueGradFunction
    from test_kernel_plan import _mixed_model_plan
    plan = _mixed_model_plan(n=60_000, groups=300)
    f = ValueGradFunction(plan, dtype=dtype)
    assert f.kernel_plan.n_secondary == 300
    theta = np.random.default_rng(2).normal(0, 0.3, plan.n_params)
    logp, grad = f(theta)
    assert_parity(logp, grad, plan, theta, dtype)
    f.close()
