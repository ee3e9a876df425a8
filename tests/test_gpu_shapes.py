"""
Parity of the CUDA path on shapes that stress the tiling: rows not a multiple of the tile, tiles
that straddle many groups (the slow path of the row kernel), one group per row, one giant group,
crossed factors with tiny primary groups, multi-CTA grids, and the BASELINE shapes at reduced size.
Inputs are seeded; the oracle is evaluated at the same (dtype-rounded) point.
"""
import numpy as np
import pytest

from helpers import assert_parity
from sakkara_b200 import synth
from sakkara_b200.valuegrad import ValueGradFunction

pytestmark = pytest.mark.gpu


def check(plan, dtype, batch=1, seed=0, **kw):
    f = ValueGradFunction(plan, dtype=dtype, max_batch=max(batch, 1), **kw)
    theta = synth.start_point(plan, batch, seed=seed, scale=0.3, dtype=dtype)
    logp, grad = f(theta)
    errs = [assert_parity(logp[b], grad[b], plan, theta[b], dtype) for b in range(batch)]
    again = f(theta)
    assert np.array_equal(again[0], logp) and np.array_equal(again[1], grad), 'not bit-reproducible'
    f.close()
    return errs


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('n', [1, 2, 31, 543, 544, 545, 1088, 5000, 70_001])
def test_linreg_row_counts(n, dtype):
    cols, _ = synth.linreg_columns(n, max(1, min(n // 3, 50)), seed=n)
    check(synth.linreg_plan(cols), dtype)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('n,groups', [(5000, 5000), (20_000, 7000), (100_000, 10_000), (300_000, 3), (200_000, 1)])
def test_linreg_group_sizes(n, groups, dtype):
    """one row per group ... one group for all rows"""
    cols, _ = synth.linreg_columns(n, groups, seed=groups)
    check(synth.linreg_plan(cols), dtype, batch=2)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_linreg_unbalanced_groups(dtype):
    rng = np.random.default_rng(5)
    n = 250_000
    sizes = np.concatenate([[120_000], rng.integers(1, 40, 3000)])
    group = np.repeat(np.arange(len(sizes)), sizes)[:n]
    group = group[rng.permutation(len(group))]
    x = rng.standard_normal(len(group))
    y = 0.5 * x + rng.standard_normal(len(group))
    check(synth.linreg_plan({'x': x, 'y': y, 'group': group}), dtype)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('n,g1,per', [(3000, 3, 4), (400_000, 100, 100), (150_000, 50, 1000)])
def test_nested_logistic(n, g1, per, dtype):
    cols, _ = synth.logistic_columns(n, g1, per, seed=n)
    check(synth.logistic_plan(cols), dtype)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('n,a,b', [(2000, 5, 3), (500_000, 1000, 100), (300_000, 30_000, 40), (200_000, 20, 1000),
                                    (100_000, 300, 1)])
def test_crossed_poisson(n, a, b, dtype):
    """large and tiny primary groups; the secondary table from 1 to 1000 entries"""
    cols, _ = synth.poisson_columns(n, a, b, seed=a)
    check(synth.poisson_plan(cols), dtype)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('batch', [1, 3])
@pytest.mark.parametrize('n,a,b', [(120_000, 60, 7), (90_000, 100, 2), (64_000, 40, 300)])
def test_crossed_poisson_two_group_tiles(n, a, b, batch, dtype):
    """
    primary groups of a few tiles: most straddling tiles hold exactly two groups and take the two-pass
    path; with few secondary codes the same code sits on both sides of the split (and in one lane)
    """
    cols, _ = synth.poisson_columns(n, a, b, seed=n % 97)
    check(synth.poisson_plan(cols), dtype, batch=batch)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_batch_of_chains(dtype):
    cols, _ = synth.linreg_columns(300_000, 1000, seed=3)
    check(synth.linreg_plan(cols), dtype, batch=9)
    cols, _ = synth.poisson_columns(200_000, 500, 50, seed=3)
    check(synth.poisson_plan(cols), dtype, batch=5)


def test_wide_storage_of_observations():
    """counts above 255 keep the observed column in floating point"""
    cols, _ = synth.poisson_columns(50_000, 100, 10, seed=1)
    cols['y'] = cols['y'] * 300.0
    plan = synth.poisson_plan(cols)
    f = ValueGradFunction(plan, dtype='float64')
    assert f.kernel_plan.y_kind == 0
    f.close()
    check(plan, 'float64', seed=2)
    check(synth.poisson_plan(synth.poisson_columns(50_000, 100, 10, seed=1)[0]), 'float32', narrow_y=False)


def test_secondary_level_too_large_is_rejected():
    from sakkara_b200.runtime import SkkError
    rng = np.random.default_rng(0)
    n = 400_000
    cols = {'x': rng.standard_normal(n), 'y': rng.poisson(1.0, n).astype(float),
            'A': rng.integers(0, 150_000, n), 'B': rng.integers(0, 120_000, n)}
    f = ValueGradFunction(synth.poisson_plan(cols), dtype='float64')
    with pytest.raises(SkkError, match='shared memory'):
        f.skk


# ---- chain-per-lane path (K2): batches of >= 32 parameter vectors, models without a secondary level ----
@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('batch', [32, 45, 128])
def test_chain_per_lane_linreg(batch, dtype):
    cols, _ = synth.linreg_columns(60_000, 300, seed=batch)
    plan = synth.linreg_plan(cols)
    f = ValueGradFunction(plan, dtype=dtype, max_batch=batch)
    theta = synth.start_point(plan, batch, seed=2, scale=0.3, dtype=dtype)
    logp, grad = f(theta)
    assert f.stats()['row_kernel_launches'] == 1          # one pass over the rows for the whole batch
    for b in (0, 1, batch // 2, batch - 1):
        assert_parity(logp[b], grad[b], plan, theta[b], dtype)
    again = f(theta)
    assert np.array_equal(again[0], logp) and np.array_equal(again[1], grad)
    # the same points through the row-parallel kernels (batch tiles of 4) agree
    g = ValueGradFunction(plan, dtype=dtype, max_batch=8)
    lp4, gr4 = g(theta[:8])
    tol = 1e-12 if dtype == 'float64' else 2e-5
    assert np.allclose(lp4, logp[:8], rtol=tol) and np.allclose(gr4, grad[:8], rtol=tol, atol=tol * np.abs(grad[:8]).max())
    f.close()
    g.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('n,groups', [(1, 1), (513, 2), (5000, 5000), (40_000, 7), (100_000, 1)])
def test_chain_per_lane_group_sizes(n, groups, dtype):
    cols, _ = synth.linreg_columns(n, groups, seed=n)
    plan = synth.linreg_plan(cols)
    f = ValueGradFunction(plan, dtype=dtype, max_batch=40)
    theta = synth.start_point(plan, 40, seed=3, scale=0.3, dtype=dtype)
    logp, grad = f(theta)
    for b in (0, 17, 39):
        assert_parity(logp[b], grad[b], plan, theta[b], dtype)
    f.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_chain_per_lane_nested_logistic(dtype):
    cols, _ = synth.logistic_columns(80_000, 10, 20, seed=4)
    plan = synth.logistic_plan(cols)
    f = ValueGradFunction(plan, dtype=dtype, max_batch=64)
    theta = synth.start_point(plan, 64, seed=4, scale=0.3, dtype=dtype)
    logp, grad = f(theta)
    for b in (0, 31, 32, 63):
        assert_parity(logp[b], grad[b], plan, theta[b], dtype)
    f.close()


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('shape', ['crossed_long_segments', 'crossed_two_group_tiles', 'nested', 'linreg'])
def test_generic_kernels_on_tiled_shapes(shape, dtype, monkeypatch):
    """
    The run-time (non-specialised) row kernels on shapes with whole tiles inside one segment and with
    two-group tiles: the same paths as the specialised kernels of the BASELINE shapes (hand-pipelined
    common path, one-pass two-group path), which the small models of the test zoo never reach.
    """
    monkeypatch.setenv('SKK_GENERIC', '1')
    if shape == 'crossed_long_segments':
        plan = synth.poisson_plan(synth.poisson_columns(200_000, 20, 300, seed=20)[0])
    elif shape == 'crossed_two_group_tiles':
        plan = synth.poisson_plan(synth.poisson_columns(120_000, 60, 7, seed=23)[0])
    elif shape == 'nested':
        plan = synth.logistic_plan(synth.logistic_columns(150_000, 50, 2, seed=9)[0])
    else:
        plan = synth.linreg_plan(synth.linreg_columns(100_000, 40, seed=3)[0])
    check(plan, dtype)
    check(plan, dtype, batch=3)
