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


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('switch', ['SKK_NO_FUSED_FINISH', 'SKK_GENERIC', 'SKK_NO_GRAPH'])
@pytest.mark.parametrize('name', sorted(ZOO))
def test_alternative_paths_parity(name, switch, dtype, monkeypatch):
    """
    The paths a default single-GPU run does not take: the separate prior / slot-total / theta kernels
    (what sharded runs use), the run-time (non-specialised) row kernels, and direct launches instead
    of a replayed CUDA graph.  Each must agree with the oracle and, bit for bit in fp64 up to the
    summation order of the priors, with the default path.
    """
    model = build_model(ZOO[name])
    thetas = theta_points(model.plan, 3, seed=23)
    f = model.logp_dlogp_function(dtype=dtype, max_batch=4)
    base = [f(theta) for theta in thetas]
    base_batch = f(thetas)
    f.close()
    monkeypatch.setenv(switch, '1')
    g = model.logp_dlogp_function(dtype=dtype, max_batch=4)
    for theta, (logp0, grad0) in zip(thetas, base):
        logp, grad = g(theta)
        assert_parity(logp, grad, model.plan, theta, dtype)
        tol = 1e-12 if dtype == 'float64' else 1e-5
        assert abs(float(logp) - float(logp0)) <= tol * max(1.0, abs(float(logp0)))
    logp_b, grad_b = g(thetas)
    for b in range(len(thetas)):
        assert_parity(logp_b[b], grad_b[b], model.plan, thetas[b], dtype)
    if switch == 'SKK_NO_GRAPH':   # same kernels, same order: bit-identical
        assert np.array_equal(base_batch[0], logp_b) and np.array_equal(base_batch[1], grad_b)
    g.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('name,batch', [('crossed_gamma', 1), ('crossed_gamma', 6), ('grouped_gamma_times', 3),
                                        ('grouped_gamma_shifted', 40)])
def test_gamma_glm_batches(name, batch, dtype):
    """The Gamma-log family through the row kernel (batch tiles of 1 and 4) and the chain-per-lane kernel (>= 32
    vectors, no crossed factor), on a frame repeated to a few ragged tiles."""
    import pandas as pd
    df, make = ZOO[name]
    big = pd.concat([df] * 9, ignore_index=True).iloc[:-7]
    from model_zoo import crossed_gamma_likelihood, grouped_gamma_likelihood
    make = crossed_gamma_likelihood(big) if name == 'crossed_gamma' else \
        grouped_gamma_likelihood(big, 'times' if name.endswith('times') else 'shifted')
    model = build_model((big, make))
    f = model.logp_dlogp_function(dtype=dtype, max_batch=max(batch, 1))
    thetas = theta_points(model.plan, max(batch, 2), seed=6, scale=0.3)[:batch]
    logp, grad = f(thetas if batch > 1 else thetas[0])
    logp, grad = np.atleast_1d(logp), np.atleast_2d(grad)
    for b in sorted({0, batch // 2, batch - 1}):
        assert_parity(logp[b], grad[b], model.plan, thetas[b], dtype)
    if batch >= 32:
        assert f.stats()['row_kernel_launches'] == 1         # chain-per-lane: one pass over the rows for the whole batch
    f.close()
