"""
Parity of the CUDA path (through the C ABI, host buffers) with the CPU oracle.
fp64: 1e-10, fp32: 1e-4, relative to the L1 mass of each reduction (tests/helpers.py).
"""
import numpy as np
import pytest

from helpers import assert_parity, build_model, theta_points
from model_zoo import zoo

pytestmark = pytest.mark.gpu

ZOO = zoo()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('name', sorted(ZOO))
def test_single_point_parity(name, dtype):
    model = build_model(ZOO[name])
    f = model.logp_dlogp_function(dtype=dtype)
    for theta in theta_points(model.plan, 4, seed=11):
        logp, grad = f(theta)
        assert_parity(logp, grad, model.plan, theta, dtype)
    assert f.stats()['launches'] > 0
    f.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('batch', [2, 4, 7])
@pytest.mark.parametrize('name', ['readme', 'nested_logistic', 'crossed_poisson'])
def test_batch_parity(name, batch, dtype):
    model = build_model(ZOO[name])
    f = model.logp_dlogp_function(dtype=dtype, max_batch=8)
    thetas = theta_points(model.plan, batch, seed=5)
    logp, grad = f(thetas)
    assert logp.shape == (batch,) and grad.shape == thetas.shape
    for b in range(batch):
        assert_parity(logp[b], grad[b], model.plan, thetas[b], dtype)
    # a batch is bit-identical to evaluating the points one by one with the same kernel shape
    logp2, grad2 = f(thetas)
    assert np.array_equal(logp, logp2) and np.array_equal(grad, grad2)
    f.close()


def test_bit_reproducible():
    model = build_model(ZOO['crossed_poisson'])
    theta = theta_points(model.plan, 2, seed=3)[1]
    f = model.logp_dlogp_function(dtype='float32')
    first = f(theta)
    for _ in range(5):
        again = f(theta)
        assert first[0] == again[0] and np.array_equal(first[1], again[1])
    f.close()
