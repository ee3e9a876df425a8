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
