"""Host-side behaviour of the ValueGradFunction / runtime wrapper that does not need a GPU."""
import os
import pickle

import numpy as np
import pytest

from sakkara_b200 import synth
from sakkara_b200.valuegrad import ValueGradFunction


@pytest.fixture
def plan():
    return synth.linreg_plan(synth.linreg_columns(500, 5, seed=1)[0])


def test_surface_of_pymc_valuegradfunction(plan):
    f = ValueGradFunction(plan, dtype='float32', max_batch=3)
    assert f.size == plan.n_params and f.dtype == np.float32
    f.set_extra_values({})
    assert f.get_extra_values() == {}
    kp = f.kernel_plan
    assert kp.dtype == np.float32 and kp.n_rows == 500 and kp is f.kernel_plan      # cached


def test_pickle_drops_the_device_handle(plan):
    f = ValueGradFunction(plan)
    f._skk, f._pid = object(), os.getpid()            # stands for a live device plan
    g = pickle.loads(pickle.dumps(f))
    assert g._skk is None and g._pid is None and g.size == f.size
    assert np.array_equal(g.model_plan.terms[0].index, plan.terms[0].index)


def test_a_plan_is_bound_to_its_process(plan):
    from sakkara_b200.runtime import SkkPlan
    fake = SkkPlan.__new__(SkkPlan)
    import ctypes
    fake._handle = ctypes.c_void_p(1)
    fake._pid = os.getpid() + 1                      # as if created in the parent of a fork
    with pytest.raises(RuntimeError, match='forked'):
        fake._guard()
    fake._handle = ctypes.c_void_p()
    fake._pid = os.getpid()
    with pytest.raises(RuntimeError, match='destroyed'):
        fake._guard()
    fake.lib = None
    fake.close()


def test_row_shards(plan):
    f = ValueGradFunction(plan, rows=slice(0, 200), n_rows_total=500)
    assert f.kernel_plan.n_rows == 200 and f.kernel_plan.n_rows_total == 500


def test_chunked_secondary_parts_add_up_to_the_model():
    """A crossed model whose secondary factor is split over several plans (valuegrad.ChunkedSecondaryValueGrad): the
    parts partition the rows, every part's secondary level respects the limit, and priors + parts -- each evaluated
    from its KERNEL plan, i.e. what the device would compute -- add up to the oracle of the whole model."""
    from oracle.kernel_plan_eval import likelihood_from_kernel_plan
    from oracle.logp_numpy import logp_dlogp
    from sakkara_b200 import synth
    from sakkara_b200.plan.kernel_plan import make_kernel_plan
    from sakkara_b200.valuegrad import secondary_chunk_groups, split_rows_by_secondary
    cols, _ = synth.poisson_columns(6000, 90, 50, seed=4)
    plan = synth.poisson_plan(cols)
    assert len(split_rows_by_secondary(plan, 64)) == 1              # fits: one part, all rows
    parts = split_rows_by_secondary(plan, 16)
    assert len(parts) == 4                                          # 50 groups -> 4 parts of <= 13
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(plan.n_rows))
    theta = synth.start_point(plan, 1, seed=3, scale=0.3)[0].astype(np.float64)
    logp, grad = logp_dlogp(plan.take_rows([]), theta)              # the priors
    grad = np.array(grad, dtype=np.float64)
    for rows in parts:
        kp = make_kernel_plan(plan.take_rows(rows), dtype='float64', include_priors=False)
        assert 0 < kp.n_secondary <= 13 and kp.n_rows == len(rows)
        lp, g = likelihood_from_kernel_plan(kp, theta)
        logp, grad = logp + lp, grad + g
    ref_logp, ref_grad = logp_dlogp(plan, theta)
    assert abs(logp - ref_logp) < 1e-10 * abs(ref_logp)
    assert np.allclose(grad, ref_grad, rtol=1e-10, atol=1e-10)
    # the limits: tables of up to 16 KB, so that a CTA of eight warps still fits an SM
    assert secondary_chunk_groups('float32', 1) == 4096 and secondary_chunk_groups('float64', 4) == 512


def _crossed3_model(n=4000, seed=1, four=False):
    """beta * x + a[A] + b[B] + c[C] (+ d[D]) with independently drawn factors, through the DSL."""
    import pandas as pd
    import sakkara_b200.model as api
    from sakkara_b200.model import dist as pm, sym as pt
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({'x': rng.standard_normal(n), 'A': rng.integers(0, 40, n), 'B': rng.integers(0, 9, n),
                       'C': rng.integers(0, 4, n), 'D': rng.integers(0, 3, n)})
    df['y'] = rng.poisson(np.exp(0.2 * df['x'] + 0.1 * (df['C'] - 1.5))).astype(float)
    DC = api.DistributionComponent
    data = api.data_components(df[['x', 'y']])
    eta = DC(pm.Normal, name='beta') * data['x']
    for name in ('A', 'B', 'C') + (('D',) if four else ()):
        eta = eta + DC(pm.Normal, name=name.lower(), group=name, sigma=DC(pm.HalfNormal, name='sd_' + name, sigma=1.0))
    return api.build(df, api.Likelihood(pm.Poisson, mu=api.f_(pt.exp)(eta), observed=data['y']), backend='b200'), df


def test_more_than_two_crossed_factors_become_a_sum_of_plans(four=False):
    """Three crossed factors: rows split by the smallest factor; inside a part that factor's term is a single element
    of theta.  Parts evaluated from their kernel plans + priors = the oracle of the whole model."""
    from oracle.kernel_plan_eval import likelihood_from_kernel_plan
    from oracle.logp_numpy import logp_dlogp
    from sakkara_b200.plan.kernel_plan import CrossedFactorsError, make_kernel_plan
    from sakkara_b200.valuegrad import partition_rows
    model, df = _crossed3_model(four=four)
    plan = model.plan
    with pytest.raises(CrossedFactorsError) as info:
        make_kernel_plan(plan, 'float64')
    assert info.value.size == (3 if four else 4) and len(info.value.codes) == plan.n_rows
    parts = partition_rows(plan, chunk_groups=4096)
    assert len(parts) == (12 if four else 4)
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(plan.n_rows))
    for rows in parts:                                               # a part = one group of C (and of D)
        assert df['C'].values[rows].min() == df['C'].values[rows].max()
        assert not four or df['D'].values[rows].min() == df['D'].values[rows].max()
    theta = np.random.default_rng(3).normal(0, 0.3, plan.n_params)
    logp, grad = logp_dlogp(plan.take_rows([]), theta)
    grad = np.array(grad, dtype=np.float64)
    for rows in parts:
        kp = make_kernel_plan(plan.take_rows(rows), dtype='float64', include_priors=False)
        assert kp.n_primary <= 40 and kp.n_secondary <= 9
        lp, g = likelihood_from_kernel_plan(kp, theta)
        logp, grad = logp + lp, grad + g
    ref_logp, ref_grad = logp_dlogp(plan, theta)
    assert abs(logp - ref_logp) < 1e-10 * abs(ref_logp)
    assert np.allclose(grad, ref_grad, rtol=1e-10, atol=1e-10)
    # the model hands out the sum of plans (constructed lazily: no device needed to get this far)
    f = model.logp_dlogp_function()
    assert type(f).__name__ == 'PartitionedValueGrad' and len(f.parts) == len(parts)
    with pytest.raises(Exception):
        partition_rows(plan, chunk_groups=4096, max_parts=2)


def test_four_crossed_factors_next_to_a_covariate():
    """Two peeled factors and beta * x are three terms that are global inside a part, the kernels take two: the terms
    without a data column are evaluated as one (the kept element holds their sum, its gradient is copied to the others)."""
    from oracle.kernel_plan_eval import likelihood_from_kernel_plan
    from oracle.logp_numpy import logp_dlogp
    from sakkara_b200.plan.kernel_plan import make_kernel_plan
    from sakkara_b200.plan.lower import UnsupportedModelError
    from sakkara_b200.valuegrad import alias_grad, alias_theta, merge_constant_terms, partition_rows
    model, _ = _crossed3_model(four=True)
    plan = model.plan
    parts = partition_rows(plan, chunk_groups=4096)
    assert len(parts) == 12
    # (as it stands a part has three global terms; the planner moves one to the primary level, the sum of plans folds
    #  the two without a data column into one before it gets there)
    unfolded = make_kernel_plan(plan.take_rows(parts[0]), dtype='float64', include_priors=False)
    assert sorted(t.level for t in unfolded.terms) == [0, 0, 1, 1, 2]
    theta = np.random.default_rng(5).normal(0, 0.3, plan.n_params)
    logp, grad = logp_dlogp(plan.take_rows([]), theta)
    grad = np.array(grad, dtype=np.float64)
    sl = plan.block_slices()
    for rows in parts:
        part, alias = merge_constant_terms(plan.take_rows(rows))
        (host, folded, ratio), = alias                              # one fold: d's element into c's
        assert ratio == 1.0
        assert len(host) == 1 and sl['c'].start <= host[0] < sl['c'].stop and sl['d'].start <= folded[0] < sl['d'].stop
        kp = make_kernel_plan(part, dtype='float64', include_priors=False)
        lp, g = likelihood_from_kernel_plan(kp, alias_theta(theta, alias))
        logp, grad = logp + lp, grad + alias_grad(g, alias)
    ref_logp, ref_grad = logp_dlogp(plan, theta)
    assert abs(logp - ref_logp) < 1e-10 * abs(ref_logp)
    assert np.allclose(grad, ref_grad, rtol=1e-10, atol=1e-10)
    # batches: the aliasing acts on the last axis
    batch = np.stack([theta, theta + 0.1])
    assert np.array_equal(alias_theta(batch, alias)[0], alias_theta(theta, alias))
    assert alias_theta(theta, None) is theta and alias_grad(grad, None) is grad


def test_terms_beyond_the_kernels_limits_are_folded():
    """beta1 x1 + beta2 x2 + icpt + a[A] + s[A] x2 + p[parent of A] + b[B]: 3 global / 3 primary-level / 1 secondary-level
    terms.  One global term moves to the primary level in the planner; the intercept and the parent-level term -- no
    data column, determined by A -- fold into a[A].  Parts from their kernel plans + priors = the oracle."""
    import pandas as pd
    import sakkara_b200.model as api
    from oracle.kernel_plan_eval import likelihood_from_kernel_plan
    from oracle.logp_numpy import logp_dlogp
    from sakkara_b200.model import dist as pm
    from sakkara_b200.plan.kernel_plan import make_kernel_plan
    from sakkara_b200.plan.lower import UnsupportedModelError
    from sakkara_b200.valuegrad import alias_grad, alias_theta, fold_terms
    rng = np.random.default_rng(3)
    n = 900
    A = rng.integers(0, 12, n)
    df = pd.DataFrame({'x1': rng.standard_normal(n), 'x2': rng.standard_normal(n), 'A': A, 'PA': A // 3,
                       'B': rng.integers(0, 5, n)})
    df['y'] = 0.4 * df['x1'] - 0.2 * df['x2'] + 0.1 * df['PA'] + 0.5 * rng.standard_normal(n)
    DC = api.DistributionComponent
    data = api.data_components(df[['x1', 'x2', 'y']])
    eta = (DC(pm.Normal, name='beta1') * data['x1'] + DC(pm.Normal, name='beta2') * data['x2'] + DC(pm.Normal, name='icpt')
           + DC(pm.Normal, name='a', group='A', sigma=DC(pm.HalfNormal, name='a_sd')) + DC(pm.Normal, name='s', group='A') * data['x2']
           + DC(pm.Normal, name='p', group='PA') + DC(pm.Normal, name='b', group='B'))
    model = api.build(df, api.Likelihood(pm.Normal, mu=eta, sigma=DC(pm.Exponential, name='noise', lam=1.0),
                                         observed=data['y']), backend='b200')
    plan = model.plan
    with pytest.raises(UnsupportedModelError, match='compiled kernels cover'):
        make_kernel_plan(plan, 'float64')
    part, folds = fold_terms(plan, only_constant=False)
    assert len(part.terms) == len(plan.terms) - 2 and len(folds) == 2
    kp = make_kernel_plan(part, dtype='float64', include_priors=False)
    assert sorted(t.level for t in kp.terms) == [0, 0, 1, 1, 2]          # 2 global / 2 primary / 1 secondary
    theta = np.random.default_rng(4).normal(0, 0.3, plan.n_params)
    logp, grad = logp_dlogp(plan.take_rows([]), theta)
    lp, g = likelihood_from_kernel_plan(kp, alias_theta(theta, folds))
    logp, grad = logp + lp, np.array(grad, dtype=np.float64) + alias_grad(g, folds)
    ref_logp, ref_grad = logp_dlogp(plan, theta)
    assert abs(logp - ref_logp) < 1e-10 * abs(ref_logp)
    assert np.allclose(grad, ref_grad, rtol=1e-10, atol=1e-10)
    f = model.logp_dlogp_function()
    assert type(f).__name__ == 'PartitionedValueGrad' and len(f.parts) == 1 and f.aliases[0] is not None


def test_sum_of_plans_pickles_without_device_state():
    """PyMC hands the evaluator to its worker processes: a PartitionedValueGrad travels as its plans (no device needed
    to build or to unpickle it), like ValueGradFunction."""
    import pickle
    model, _ = _crossed3_model(n=400)
    f = model.logp_dlogp_function()
    assert type(f).__name__ == 'PartitionedValueGrad'
    g = pickle.loads(pickle.dumps(f))
    assert len(g.parts) == len(f.parts) and g.size == f.size and g.dtype == f.dtype
    assert all(p._skk is None for p in g.parts)
    assert all(np.array_equal(a.kernel_plan.seg_offsets, b.kernel_plan.seg_offsets) for a, b in zip(f.parts, g.parts))
