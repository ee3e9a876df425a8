"""
Kernel planner on the CPU: sort permutation, segment offsets and slot tables are exact, and a
KernelPlan evaluated with plain NumPy (oracle/kernel_plan_eval.py) reproduces the oracle on the
ModelPlan it came from -- for whole plans, for row shards, and for a 2-rank gloo all-reduce.
"""
import os

import numpy as np
import pytest

from helpers import build_model, theta_points
from model_zoo import zoo
from oracle.kernel_plan_eval import likelihood_from_kernel_plan
from oracle.logp_c import priors_only
from oracle.logp_numpy import logp_dlogp
from sakkara_b200 import synth
from sakkara_b200.distributed import shard_bounds, sorted_shard
from sakkara_b200.plan.kernel_plan import (LEVEL_GLOBAL, LEVEL_PRIMARY, LEVEL_SECONDARY, make_kernel_plan,
                                            sort_rows, stable_argsort_small)
from sakkara_b200.plan.lower import UnsupportedModelError

ZOO = zoo()


def full_eval(plan, kplans, theta):
    lp, gp = logp_dlogp(plan.take_rows([]), theta)
    for kp in kplans:
        l, g = likelihood_from_kernel_plan(kp, theta)
        lp, gp = lp + l, gp + g
    return lp, gp


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
@pytest.mark.parametrize('name', sorted(ZOO))
def test_kernel_plan_reproduces_oracle(name, dtype):
    plan = build_model(ZOO[name]).plan
    # (a sigma that differs between rows: one standard kernel plan per sigma, ModelPlan.split_by_sigma)
    kps = [make_kernel_plan(part, dtype) for part in plan.split_by_sigma()]
    for theta in theta_points(plan, 3, seed=2):
        ref = logp_dlogp(plan, theta)
        got = full_eval(plan, kps, theta)
        tol = 1e-11 if dtype == 'float64' else 1e-5
        assert abs(got[0] - ref[0]) <= tol * max(1.0, abs(ref[0]))
        assert np.allclose(got[1], ref[1], rtol=tol, atol=tol * max(1.0, np.abs(ref[1]).max()))


@pytest.mark.parametrize('name', sorted(ZOO))
def test_plan_without_rows_and_sigma_parts_are_plannable(name):
    # what the mini-batch and split-sigma evaluators hand to the planner
    plan = build_model(ZOO[name]).plan
    kp = make_kernel_plan(plan.take_rows([]), 'float64')
    assert kp.n_rows == 0 and kp.terms == [] and len(kp.blocks) == len(plan.blocks)
    parts = plan.split_by_sigma()
    assert sum(p.n_rows for p in parts) == plan.n_rows
    for part in parts:
        assert make_kernel_plan(part, 'float32', include_priors=False).n_rows == part.n_rows


def test_levels_and_storage():
    kp = make_kernel_plan(synth.poisson_plan(synth.poisson_columns(5000, 40, 7, seed=1)[0]), 'float32')
    assert [t.level for t in kp.terms] == [LEVEL_GLOBAL, LEVEL_PRIMARY, LEVEL_SECONDARY]
    assert kp.n_primary == 40 and kp.n_secondary == 7          # the larger factor is sorted, the smaller tabled
    assert kp.sec_codes.dtype == np.uint16 and kp.y.dtype == np.uint8 and kp.y_kind == 1
    assert kp.stored_bytes_per_row == 4 + 1 + 2
    kp = make_kernel_plan(synth.logistic_plan(synth.logistic_columns(5000, 5, 6, seed=1)[0]), 'float32')
    assert [t.level for t in kp.terms] == [LEVEL_PRIMARY, LEVEL_PRIMARY] and kp.n_secondary == 0
    assert kp.stored_bytes_per_row == 4 + 1                    # nested parent level needs no column
    kp = make_kernel_plan(synth.linreg_plan(synth.linreg_columns(5000, 9, seed=1)[0]), 'float64')
    assert kp.stored_bytes_per_row == 16 and kp.y_kind == 0


def test_sorted_rows_and_segments_are_exact():
    cols, _ = synth.poisson_columns(20_000, 50, 9, seed=4)
    plan = synth.poisson_plan(cols)
    kp = make_kernel_plan(plan, 'float64')
    a_codes, b_codes = plan.terms[1].index, plan.terms[2].index
    perm = kp.perm
    assert sorted(perm.tolist()) == list(range(plan.n_rows))             # a permutation
    slot_of_theta = {int(t): s for s, t in enumerate(kp.terms[1].theta_index)}
    a_off = plan.blocks[plan.terms[1].block].offset
    p_slot = np.array([slot_of_theta[a_off + int(c)] for c in a_codes[perm]])
    assert (np.diff(p_slot) >= 0).all()
    assert np.array_equal(np.searchsorted(p_slot, np.arange(kp.n_primary + 1)), kp.seg_offsets)
    # inside every segment the secondary codes ascend, and ties keep the original row order (stable)
    for s in range(kp.n_primary):
        lo, hi = kp.seg_offsets[s], kp.seg_offsets[s + 1]
        sec = kp.sec_codes[lo:hi].astype(int)
        assert (np.diff(sec) >= 0).all()
        ties = np.diff(sec) == 0
        assert (np.diff(perm[lo:hi])[ties] > 0).all()
    # secondary slot table maps back to the block's elements
    b_off = plan.blocks[plan.terms[2].block].offset
    assert np.array_equal(kp.terms[2].theta_index[kp.sec_codes.astype(int)] - b_off, b_codes[perm])


def test_nested_parents_are_adjacent_in_slot_order():
    cols, _ = synth.logistic_columns(30_000, 6, 5, seed=9)
    kp = make_kernel_plan(synth.logistic_plan(cols), 'float64')
    slope_slots = kp.terms[0].theta_index          # theta index of the g1 parent of every g2 slot
    assert (np.diff(slope_slots) != 0).sum() == 6 - 1            # each parent's children are contiguous


def test_already_sorted_rows_need_no_permutation():
    cols, _ = synth.linreg_columns(4000, 11, seed=2)
    order = np.argsort(cols['group'], kind='stable')
    cols = {k: v[order] for k, v in cols.items()}
    assert make_kernel_plan(synth.linreg_plan(cols), 'float64').perm is None


def test_radix_argsort_matches_numpy():
    rng = np.random.default_rng(0)
    for n_keys in (3, 70_000, 5_000_000):
        keys = rng.integers(0, n_keys, 200_000)
        assert np.array_equal(stable_argsort_small(keys, n_keys), np.argsort(keys, kind='stable'))
    p = rng.integers(0, 100_000, 50_000)
    s = rng.integers(0, 30, 50_000)
    assert np.array_equal(sort_rows(p, 100_000, s, 30), np.lexsort((s, p)))


def test_three_crossed_factors_are_rejected():
    rng = np.random.default_rng(0)
    n = 3000
    cols = {'x': rng.standard_normal(n), 'y': rng.poisson(1.0, n).astype(float),
            'A': rng.integers(0, 11, n), 'B': rng.integers(0, 7, n)}
    plan = synth.poisson_plan(cols)
    from sakkara_b200.plan.model_plan import Term
    plan.terms.append(Term(plan.terms[1].block, rng.integers(0, 5, n), None))
    with pytest.raises(UnsupportedModelError, match='crossed'):
        make_kernel_plan(plan, 'float64')


@pytest.mark.parametrize('world', [2, 3, 8])
def test_row_shards_add_up(world):
    plan = synth.poisson_plan(synth.poisson_columns(30_000, 60, 8, seed=6)[0])
    theta = synth.start_point(plan, 1, seed=1)[0]
    ref = logp_dlogp(plan, theta)
    for shards in ([slice(*shard_bounds(plan.n_rows, r, world)) for r in range(world)],
                   [sorted_shard(plan, r, world) for r in range(world)]):
        kps = [make_kernel_plan(plan, 'float64', rows=rows, n_rows_total=plan.n_rows) for rows in shards]
        assert sum(kp.n_rows for kp in kps) == plan.n_rows
        got = full_eval(plan, kps, theta)
        assert abs(got[0] - ref[0]) <= 1e-10 * abs(ref[0])
        assert np.allclose(got[1], ref[1], rtol=1e-10, atol=1e-9)
    # sorting before cutting keeps the primary groups (almost) whole on one rank
    sorted_groups = [make_kernel_plan(plan, 'float64', rows=sorted_shard(plan, r, world)).n_primary for r in range(world)]
    assert sum(sorted_groups) <= 60 + world - 1


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        plan = synth.poisson_plan(synth.poisson_columns(20_000, 40, 6, seed=8)[0])
        theta = synth.start_point(plan, 1, seed=2)[0]
        kp = make_kernel_plan(plan, 'float64', rows=sorted_shard(plan, rank, world), n_rows_total=plan.n_rows)
        lp, g = likelihood_from_kernel_plan(kp, theta)
        vec = torch.from_numpy(np.concatenate([g, [lp]]))
        dist.all_reduce(vec)                       # the one collective of the row-sharded path
        pl, pg = logp_dlogp(priors_only(plan), theta)
        ref = logp_dlogp(plan, theta)
        got_g, got_l = vec[:-1].numpy() + pg, float(vec[-1]) + pl
        ok = abs(got_l - ref[0]) <= 1e-10 * abs(ref[0]) and np.allclose(got_g, ref[1], rtol=1e-10, atol=1e-9)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce():
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as manager:
        out = manager.dict()
        mp.spawn(_gloo_worker, args=(world, 29533 + os.getpid() % 500, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _mixed_model_plan(n=2500, groups=37, covariates=2, seed=4):
    """intercept + a[g] + s1[g] x1 + s2[g] x2 (+ beta_k x_k): more terms than one level takes."""
    import pandas as pd
    import sakkara_b200.model as api
    from sakkara_b200.model import dist as pm
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({'x1': rng.standard_normal(n), 'x2': rng.standard_normal(n), 'x3': rng.standard_normal(n),
                       'g': rng.integers(0, groups, n)})
    df['y'] = 0.3 * df['x1'] - 0.2 * df['x2'] + 0.5 * rng.standard_normal(n)
    DC = api.DistributionComponent
    data = api.data_components(df[['x1', 'x2', 'x3', 'y']])
    eta = DC(pm.Normal, name='a', group='g', sigma=DC(pm.HalfNormal, name='a_sd')) \
        + DC(pm.Normal, name='s1', group='g') * data['x1'] + DC(pm.Normal, name='s2', group='g') * data['x2']
    for k in range(covariates):
        eta = eta + DC(pm.Normal, name=f'beta{k + 1}') * data[f'x{k + 1}']
    if covariates == 3:
        eta = eta + DC(pm.Normal, name='icpt')
    lik = api.Likelihood(pm.Normal, mu=eta, sigma=DC(pm.Exponential, name='noise', lam=1.0), observed=data['y'])
    return api.build(df, lik, backend='b200').plan


@pytest.mark.parametrize('covariates,levels', [(2, [0, 0, 1, 1, 2]), (3, [0, 0, 1, 1, 1, 2])])
def test_terms_move_between_levels(covariates, levels):
    """An intercept and two random slopes by one group are three primary-level terms: the third moves to the free
    secondary level on a copy of the group's column.  A third covariate and a global intercept next to them are four
    global terms: two would have to ride on the primary level, which is full -- still more than one plan takes."""
    from oracle.kernel_plan_eval import likelihood_from_kernel_plan
    plan = _mixed_model_plan(covariates=covariates)
    if covariates == 3:
        with pytest.raises(UnsupportedModelError, match='compiled kernels cover'):
            make_kernel_plan(plan, 'float64')
        return
    kp = make_kernel_plan(plan, 'float64', include_priors=False)
    assert sorted(t.level for t in kp.terms) == levels
    assert kp.n_secondary == kp.n_primary == 37 and kp.sec_codes.dtype == np.uint16
    # the copy is sorted inside every primary segment: it is constant there
    for k in range(kp.n_primary):
        segment = kp.sec_codes[kp.seg_offsets[k]:kp.seg_offsets[k + 1]]
        assert segment.min() == segment.max()
    theta = np.random.default_rng(2).normal(0, 0.3, plan.n_params)
    logp, grad = likelihood_from_kernel_plan(kp, theta)
    prior_logp, prior_grad = logp_dlogp(plan.take_rows([]), theta)
    ref_logp, ref_grad = logp_dlogp(plan, theta)
    assert abs(logp + prior_logp - ref_logp) < 1e-10 * abs(ref_logp)
    assert np.allclose(grad + prior_grad, ref_grad, rtol=1e-10, atol=1e-10)
