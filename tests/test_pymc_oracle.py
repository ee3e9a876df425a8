"""
The floating-point half against the REAL reference (sakkara on PyMC/PyTensor) -- runs only where PyMC
is importable (it is not in the build image).  When it runs, it pins the NumPy oracle,
and with it every parity test of the CUDA path, to ``model.logp_dlogp_function()`` of the reference.
"""
import numpy as np
import pytest

from helpers import build_model, theta_points
from model_zoo import zoo
from oracle import pymc_oracle
from oracle.logp_numpy import logp_dlogp

pytestmark = pytest.mark.skipif(not pymc_oracle.available(), reason=pymc_oracle.describe())

ZOO = zoo()


@pytest.mark.parametrize('name', sorted(ZOO))
def test_numpy_oracle_equals_pymc(name):
    df, make = ZOO[name]
    model = build_model(ZOO[name])
    f, names = pymc_oracle.reference_logp_dlogp(make, df)
    assert names == model.value_var_names
    for theta in theta_points(model.plan, 4, seed=17):
        ref_logp, ref_grad = f(theta)
        logp, grad = logp_dlogp(model.plan, theta)
        assert logp == pytest.approx(ref_logp, rel=1e-10, abs=1e-10)
        np.testing.assert_allclose(grad, ref_grad, rtol=1e-10, atol=1e-9)


def test_describe_names_the_oracle():
    assert 'PyMC' in pymc_oracle.describe()
