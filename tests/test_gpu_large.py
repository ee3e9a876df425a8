"""
Parity at the sizes BASELINE.json names -- the paths that only long shards reach: tens of
tiles per warp (the TMA ring in steady state), fp32 accumulator tables that collect many tiles, shards
whose primary groups span many tiles.  The oracle here is the fused C restatement (oracle/logp_c.c, fp64
likelihood) + the NumPy priors; the NumPy oracle alone would take minutes at these sizes.  Error metric:
|got - ref| / max(|ref|, 1e-3 max|ref|) per gradient entry, relative error of logp; bound 1e-4 (fp32) /
1e-10 (fp64) as north_star states.
"""
import os

import numpy as np
import pytest

from sakkara_b200 import synth
from sakkara_b200.valuegrad import ValueGradFunction

pytestmark = pytest.mark.gpu
TOL = {'float32': 1e-4, 'float64': 1e-10}


def c_oracle(plan, theta):
    from oracle.logp_c import COracle
    f = COracle(plan, dtype='float64', threads=os.cpu_count() or 1)
    return f(np.asarray(theta, dtype=np.float64))


def compare(plan, dtype, batch=1, peer_self=False, seed=0):
    f = ValueGradFunction(plan, dtype=dtype, max_batch=batch, device=0)
    if peer_self:                              # the sharded pipeline (push, wait, combine) with this rank as its only peer
        f.skk.peer_init([f.skk.peer_alloc(1)], 0, 1)
    theta = synth.start_point(plan, batch, seed=seed, scale=0.1, dtype=dtype)
    logp, grad = f(theta)
    for b in range(batch):
        ref_l, ref_g = c_oracle(plan, theta[b])
        err_l = abs(float(logp[b]) - ref_l) / max(abs(ref_l), 1.0)
        scale = np.maximum(np.abs(ref_g), np.abs(ref_g).max() * 1e-3)
        err_g = float(np.max(np.abs(np.asarray(grad[b], dtype=np.float64) - ref_g) / scale))
        assert np.isfinite(err_l) and err_l <= TOL[dtype], f'logp error {err_l:.3e}'
        assert err_g <= TOL[dtype], f'gradient error {err_g:.3e}'
    again = f(theta)
    assert np.array_equal(again[0], logp) and np.array_equal(again[1], grad), 'not bit-reproducible'
    st = f.stats()
    f.close()
    return st


def test_c3_full_size():
    """BASELINE configuration 3 as bench.py builds it: 10M rows, nested 100 x 100, Bernoulli-logit, fp32"""
    import bench
    plan = bench.make_plan('c3', 10_000_000, 0, 1)
    st = compare(plan, 'float32')
    assert st['n_tiles'] > 15_000


@pytest.mark.parametrize('peer_self', [False, True])
def test_c4_shard_of_eight(peer_self):
    """the rows of rank 0 of an 8-way sharded configuration 4: 12.5M rows, 1,250 primary groups x 1,000 codes"""
    import bench
    plan = bench.make_plan('c4', 100_000_000, 0, 8)
    assert plan.n_rows == 12_500_000
    compare(plan, 'float32', peer_self=peer_self)


def test_long_shard_with_few_secondary_codes():
    """
    20M rows x 50 primary groups x 2 secondary codes: every entry of a warp's fp32 accumulator table collects
    the warp's whole share of the shard (the worst case for the table's precision).
    """
    cols, _ = synth.poisson_columns(20_000_000, 50, 2, seed=11)
    compare(synth.poisson_plan(cols), 'float32')


def test_c2_full_size_fp64_batch():
    import bench
    plan = bench.make_plan('c2', 1_000_000, 0, 1)
    compare(plan, 'float64', batch=4)
