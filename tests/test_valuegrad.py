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
