"""
The PyMC boundary: ``build()`` must hand back a ``pm.Model`` that ``pm.sample()`` can drive --
reference usage ``with build(df, likelihood): pm.sample()`` (``/root/reference/README.md:33-34``,
``sakkara/model/utils.py:31-33``).  PyMC / PyTensor are not installable here, so the glue of
``sakkara_b200/pytensor_op.py`` runs against ``tests/pymc_stub.py``, an *evaluating* stand-in for the
interfaces it uses; what the stub's ``Model.logp_dlogp_function()`` returns is what NUTS would see.
"""
import pickle

import numpy as np
import pytest

import pymc_stub
from helpers import build_model, theta_points
from model_zoo import zoo
from oracle.logp_numpy import logp_dlogp

ZOO = zoo()


class OracleValueGrad:
    """CPU stand-in for ValueGradFunction (the CUDA evaluator needs a GPU): same calling convention."""
    dtype = np.dtype('float64')

    def __init__(self, plan):
        self.plan = plan
        self.calls = 0

    def __call__(self, theta):
        self.calls += 1
        return logp_dlogp(self.plan, np.asarray(theta, dtype=np.float64))


@pytest.fixture()
def stub():
    pm = pymc_stub.install()
    yield pm
    pymc_stub.uninstall()


def test_without_pymc_build_returns_the_b200_model():
    import sakkara_b200.model as api
    pymc_stub.uninstall()
    model = build_model(ZOO['readme'])
    assert isinstance(model, api.B200Model)
    with pytest.raises(ImportError):
        model.to_pymc()


@pytest.mark.parametrize('name', sorted(ZOO))
def test_build_returns_a_pymc_model_with_the_reference_layout(stub, name):
    import sakkara_b200.model as api
    model = build_model(ZOO[name])                       # backend='auto': PyMC is importable now
    assert isinstance(model, stub.Model) and isinstance(model.b200, api.B200Model)
    plan = model.b200.plan
    # one free variable per block, creation order, <name>_log__ value variables for positive blocks
    assert [rv.name for rv in model.free_RVs] == [b.name for b in plan.blocks]
    assert [v.name for v in model.value_vars] == [b.value_name for b in plan.blocks]
    assert [n for n in (v.name for v in model.value_vars) if n.endswith('_log__')] == \
        [b.name + '_log__' for b in plan.blocks if b.transform == 'log']
    # dims and coords as the reference registers them (pm.Model(coords=groupset.coords()))
    for b in plan.blocks:
        assert model.named_vars_to_dims[b.name] == (tuple(b.dims) if b.dims else None)
    assert set(model.coords) == set(model.b200.coords)
    # data columns under the reference's names (fixed/data.py:40-41)
    assert set(model.data) == {d.name for d in model.b200.recorded.data_vars}
    with model:                                          # usable as a context, like the reference's return value
        assert stub.Model.get_context() is model


@pytest.mark.parametrize('name', sorted(ZOO))
def test_value_and_gradient_through_the_op_equal_the_oracle(stub, name):
    from sakkara_b200.pytensor_op import to_pymc_model
    b200 = build_model((ZOO[name][0], ZOO[name][1])).b200
    vg = OracleValueGrad(b200.plan)
    model = to_pymc_model(b200, value_grad=vg)
    f = model.logp_dlogp_function()
    assert f.size == b200.plan.n_params
    for theta in theta_points(b200.plan, 3, seed=3):
        before = vg.calls
        logp, grad = f(theta)
        # value AND gradient of one point cost one evaluation: grad() reuses the forward application
        assert vg.calls == before + 1
        ref_l, ref_g = logp_dlogp(b200.plan, theta)
        # PyMC's own Jacobian of HalfFlat and the one inside our logp must not be counted twice
        assert abs(logp - ref_l) <= 1e-12 * max(1.0, abs(ref_l))
        assert np.allclose(grad, ref_g, rtol=1e-12, atol=1e-12)


def test_op_outputs_have_the_evaluator_dtype(stub):
    from sakkara_b200.pytensor_op import make_op
    b200 = build_model(ZOO['readme']).b200
    vg = OracleValueGrad(b200.plan)
    op = make_op(vg)
    import pytensor.tensor as pt
    logp, dlogp = op(pt.vector('theta'))
    assert logp.ndim == 0 and dlogp.ndim == 1 and logp.dtype == dlogp.dtype == 'float64'
    assert logp.owner is dlogp.owner                     # two outputs of one application
    assert op.connection_pattern(logp.owner) == [[True, False]]


def test_value_grad_function_pickles_without_its_device_plan():
    """PyMC sends the step method to one worker process per chain: the device plan must not travel."""
    import sakkara_b200.model as api
    pymc_stub.uninstall()
    model = build_model(ZOO['nested_logistic'])
    f = model.logp_dlogp_function(dtype='float32', max_batch=2)
    f._skk, f._pid = object(), 12345                     # pretend a plan exists in this process
    g = pickle.loads(pickle.dumps(f, protocol=pickle.HIGHEST_PROTOCOL))
    assert g._skk is None and g._pid is None
    assert g.dtype == np.float32 and g.max_batch == 2 and g.size == f.size
    assert [b.name for b in g.model_plan.blocks] == [b.name for b in model.plan.blocks]
    for t_f, t_g in zip(f.model_plan.terms, g.model_plan.terms):
        assert np.array_equal(t_f.index, t_g.index)
    f._skk = None


@pytest.mark.gpu
@pytest.mark.parametrize('name,dtype,tol', [('readme', 'float64', 1e-10), ('nested_logistic', 'float64', 1e-10),
                                            ('crossed_poisson', 'float32', 1e-4), ('positive_mu', 'float64', 1e-10)])
def test_cuda_evaluator_behind_the_pymc_model(stub, name, dtype, tol):
    """the model build() returns, with the CUDA library behind the Op, against the oracle"""
    import sakkara_b200.model as api
    from helpers import assert_parity
    df, make = ZOO[name]
    model = api.build(df, make(api, api.dist, api.sym), dtype=dtype, device=0)
    f = model.logp_dlogp_function()
    plan = model.b200.plan
    for theta in theta_points(plan, 2, seed=4, scale=0.3):
        theta = theta.astype(dtype).astype(np.float64)
        logp, grad = f(theta)
        assert_parity(logp, grad, plan, theta, dtype)
    assert model.b200_value_grad.stats()['launches'] > 0
    model.b200_value_grad.close()
