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


@pytest.mark.skipif(not ref_harness.reference_available(), reason='reference sources not present')
@pytest.mark.parametrize('seed', range(0, 60))
def test_random_models_match_live_reference(seed):
    """The random models of tests/test_composition_cpu.py built by the reference's own ``build()`` (recording fake PyMC)
    and by ours: the lowered plans -- theta layout, every gather index, data columns -- are identical.  A NaN-masked
    likelihood is the one designed difference (the reference keeps the masked rows with constant parameters through a
    per-row GroupComponent, we drop them and keep their constant): there logp and gradient must agree instead."""
    import warnings
    import numpy as np
    from oracle.logp_numpy import logp_dlogp
    from sakkara_b200.model import build
    from sakkara_b200.plan.lower import UnsupportedModelError
    from test_composition_cpu import random_model
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            frame, likelihood, description = random_model(seed)
            ours = build(frame, likelihood, backend='b200').plan
            ref = ref_harness.reference_plan(lambda api, pm, pt: random_model(seed, api, pm, pt)[1], frame)
    except (UnsupportedModelError, NotImplementedError) as err:
        pytest.skip(f'outside the supported class: {err}')
    except ValueError as err:
        if 'minimal' in str(err):
            pytest.skip('the random frame made a parent column a twin (both implementations raise)')
        raise
    if 'nan' not in description.split('+'):
        assert_same(plan_to_arrays(ref), plan_to_arrays(ours))
        return
    theta = np.random.default_rng(seed).normal(0, 0.4, ours.n_params)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        (ref_logp, ref_grad), (logp, grad) = logp_dlogp(ref, theta), logp_dlogp(ours, theta)
    assert logp == pytest.approx(ref_logp, rel=1e-10)
    assert np.allclose(grad, ref_grad, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('seed', [0, 1, 3, 5, 6, 10, 14, 16, 19, 22, 28, 35])
def test_random_models_match_golden(seed):
    """Twelve of the random models, pinned by fixtures the reference's own ``build()`` produced
    (tests/golden/make_golden.py: random_models) -- the same check where /root/reference is not present."""
    import warnings
    import numpy as np
    from sakkara_b200.model import build
    from test_composition_cpu import random_model
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        frame, likelihood, _ = random_model(seed)
        ours = build(frame, likelihood, backend='b200').plan
    golden = dict(np.load(os.path.join(GOLDEN, f'random_{seed:02d}.npz'), allow_pickle=False))
    assert_same(golden, plan_to_arrays(ours))
