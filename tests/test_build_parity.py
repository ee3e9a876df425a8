"""
Integer half of the hot path: ``sakkara_b200.model.build`` must produce the theta layout and the
gather indices the reference's ``sakkara.model.build`` produces -- bit-exact.

* against golden fixtures generated from the reference (tests/golden/make_golden.py), always;
* against the reference itself, run live through the recording fake PyMC, where /root/reference
  exists (not on the GPU box).
"""
import os

import numpy as np
import pytest

from golden.plan_io import assert_same, plan_to_arrays
from helpers import build_model
from model_zoo import zoo
from oracle import ref_harness

ZOO = zoo()
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', sorted(ZOO))
def test_matches_golden(name):
    mine = plan_to_arrays(build_model(ZOO[name]).plan)
    with np.load(os.path.join(GOLDEN, f'{name}.npz')) as stored:
        assert_same({k: stored[k] for k in stored.files}, mine)


@pytest.mark.skipif(not ref_harness.reference_available(), reason='reference sources not present')
@pytest.mark.parametrize('name', sorted(ZOO))
def test_matches_live_reference(name):
    df, make = ZOO[name]
    assert_same(plan_to_arrays(ref_harness.reference_plan(make, df)), plan_to_arrays(build_model(ZOO[name]).plan))


@pytest.mark.skipif(not ref_harness.reference_available(), reason='reference sources not present')
@pytest.mark.parametrize('seed', range(4))
def test_random_frames_match_live_reference(seed):
    """Random nested + crossed structure, shuffled rows, string and integer labels."""
    import pandas as pd
    from model_zoo import crossed_poisson_likelihood, nested_logistic_likelihood
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(30, 120))
    g2 = rng.integers(0, 9, n)
    frame = pd.DataFrame({'x': rng.standard_normal(n), 'y': rng.integers(0, 2, n),
                          'g1': [f'p{v}' for v in g2 // 3], 'g2': g2 * 7})
    make = nested_logistic_likelihood(frame)
    assert_same(plan_to_arrays(ref_harness.reference_plan(make, frame)), plan_to_arrays(build_model((frame, make)).plan))
    frame = pd.DataFrame({'x': rng.standard_normal(n), 'y': rng.poisson(1.0, n),
                          'A': rng.integers(0, 5, n), 'B': [f'b{v}' for v in rng.integers(0, 3, n)]})
    for first in (True, False):
        make = crossed_poisson_likelihood(frame, first)
        try:
            ref = ref_harness.reference_plan(make, frame)
        except Exception as err:          # e.g. a missing (A, B) combination in the dense spelling
            with pytest.raises(type(err)):
                build_model((frame, make))
            continue
        assert_same(plan_to_arrays(ref), plan_to_arrays(build_model((frame, make)).plan))


def test_theta_layouts():
    """Creation order of SURVEY.md §A."""
    assert build_model(ZOO['readme']).value_var_names == \
        ['mu_k', 'sigma_k_log__', 'k', 'intercept', 'sigma_est_log__']
    assert build_model(ZOO['nested_logistic']).value_var_names == \
        ['mu_slope', 'sigma_slope_log__', 'slope', 'mu_mu_icpt', 'sigma_mu_icpt_log__', 'mu_icpt',
         'sigma_icpt_log__', 'icpt']
    assert build_model(ZOO['crossed_poisson']).value_var_names == \
        ['beta', 'sigma_a_log__', 'a', 'sigma_b_log__', 'b']
    model = build_model(ZOO['readme'])
    assert [rv.dims for rv in model.free_RVs] == [('global',), ('global',), ('group',), ('global',), ('global',)]
    assert model.observed_RVs[0].dims == ('obs',) and model.observed_RVs[0].shape == (10,)
    assert set(model.coords) == {'global', 'obs', 'group'}


def test_both_crossed_spellings_lower_to_the_same_gathers():
    a, b = build_model(ZOO['crossed_poisson']).plan, build_model(ZOO['crossed_poisson_dense']).plan
    by_name = lambda p: {p.blocks[t.block].name: t for t in p.terms}
    ta, tb = by_name(a), by_name(b)
    assert ta.keys() == tb.keys() == {'beta', 'a', 'b'}
    for name in ta:
        assert np.array_equal(ta[name].index, tb[name].index)
