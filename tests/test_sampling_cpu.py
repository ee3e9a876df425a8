"""
The many-chain HMC loop (``sakkara_b200.sampling.hmc_chains``) on CPU tensors with the NumPy oracle as the
evaluator: step-size adaptation, mass-matrix update, acceptance and bookkeeping are device-agnostic torch
code, so they can be checked without a GPU.  (``sample_hmc`` itself -- the same loop fed by ``skk_eval``
with device pointers -- is covered by tests/test_gpu_sampling.py.)
"""
import numpy as np
import pandas as pd
import torch

from helpers import build_model
from model_zoo import readme_likelihood
from oracle.logp_numpy import logp_dlogp
from sakkara_b200.sampling import hmc_chains


def test_hmc_loop_recovers_the_posterior_mean():
    rng = np.random.default_rng(5)
    n, groups = 300, 3
    code = rng.integers(0, groups, n)
    k = np.array([1.0, -1.0, 0.5])
    x = rng.standard_normal(n)
    y = k[code] * x + 0.3 + 0.5 * rng.standard_normal(n)
    df = pd.DataFrame({'x': x, 'y': y, 'group': code})
    plan = build_model((df, readme_likelihood(df))).plan
    calls = []

    def evaluate(q, logp_out, grad_out):
        calls.append(q.shape[0])
        for c in range(q.shape[0]):
            lp, g = logp_dlogp(plan, q[c].numpy())
            logp_out[c] = float(lp)
            grad_out[c] = torch.from_numpy(np.asarray(g, dtype=np.float64))

    chains, draws, tune, leap = 6, 120, 160, 6
    gen = torch.Generator().manual_seed(1)
    theta = 0.1 * torch.randn(chains, plan.n_params, dtype=torch.float64, generator=gen)
    kept, kept_logp, accepted, eps = hmc_chains(evaluate, theta, draws, tune, leap, 0.01, 0.8, generator=gen)
    assert kept.shape == (chains, draws, plan.n_params) and kept_logp.shape == (chains, draws)
    assert len(calls) == 1 + (tune + draws) * leap and set(calls) == {chains}     # one batched call per leapfrog step
    assert 0.0 < eps < 1.0
    rate = (accepted / draws).numpy()
    assert 0.4 < rate.mean() <= 1.0
    # kept logp values are the oracle's at the kept positions
    lp, _ = logp_dlogp(plan, kept[2, 7].numpy())
    assert abs(float(kept_logp[2, 7]) - float(lp)) < 1e-9
    sl = plan.block_slices()
    order = pd.factorize(df['group'])[1]
    post_k = kept[:, :, sl['k']].mean((0, 1)).numpy()
    assert np.allclose(post_k, k[np.asarray(order)], atol=0.15)
    assert abs(float(kept[:, :, sl['intercept']].mean()) - 0.3) < 0.12
    assert abs(float(kept[:, :, sl['sigma_est']].exp().mean()) - 0.5) < 0.08


# ---------------------------------------------------------------------------------------------------------------
# NUTS (nuts_chains): the same loop sample_nuts feeds with skk_eval on the GPU


def _gaussian(mu, sd, calls=None):
    def evaluate(q, logp_out, grad_out):
        if calls is not None:
            calls.append(q.shape[0])
        z = (q - mu) / sd
        logp_out.copy_(-0.5 * (z * z).sum(1))
        grad_out.copy_(-z / sd)
    return evaluate


def test_nuts_samples_a_gaussian_with_unequal_scales():
    from sakkara_b200.sampling import ess_bulk, nuts_chains, split_rhat
    P, chains, draws, tune = 8, 24, 300, 300
    sd = torch.linspace(0.05, 4.0, P, dtype=torch.float64)            # two orders of magnitude: needs the mass matrix
    mu = torch.arange(P, dtype=torch.float64) - 3.0
    calls = []
    gen = torch.Generator().manual_seed(11)
    theta = 0.1 * torch.randn(chains, P, dtype=torch.float64, generator=gen)
    kept, kept_logp, stats, eps = nuts_chains(_gaussian(mu, sd, calls), theta, draws, tune, max_depth=8, generator=gen)
    assert set(calls) == {chains}                                     # every evaluation is one call for all chains
    k = kept.numpy()
    assert kept.shape == (chains, draws, P) and kept_logp.shape == (chains, draws)
    # moments within Monte-Carlo error of the exact ones
    assert np.all(np.abs(k.mean((0, 1)) - mu.numpy()) < 0.08 * sd.numpy())
    assert np.allclose(k.reshape(-1, P).std(0) / sd.numpy(), 1.0, atol=0.05)
    # kept logp is the target's at the kept position
    z = (kept[3, 5] - mu) / sd
    assert abs(float(kept_logp[3, 5]) + 0.5 * float((z * z).sum())) < 1e-9
    # sampler statistics: adapted to the target acceptance, no divergence on a Gaussian, trees inside the depth limit
    assert abs(float(stats['accept'].mean()) - 0.8) < 0.1
    assert int(stats['diverging'].sum()) == 0
    depth, steps = stats['depth'].numpy(), stats['n_steps'].numpy()
    assert depth.min() >= 1 and depth.max() <= 8
    assert np.all(steps <= 2 ** depth - 1) and np.all(steps >= 2 ** (depth - 1))
    # after the mass matrix is adapted the trajectories are short (a standardised Gaussian turns after a few steps)
    assert steps.mean() < 16
    assert max(split_rhat(k[:, :, j]) for j in range(P)) < 1.02
    assert min(ess_bulk(k[:, :, j]) for j in range(P)) > 0.3 * chains * draws


def test_nuts_stops_at_max_depth_and_flags_divergences():
    from sakkara_b200.sampling import nuts_chains
    P, chains = 3, 5
    mu, sd = torch.zeros(P, dtype=torch.float64), torch.ones(P, dtype=torch.float64)
    gen = torch.Generator().manual_seed(2)
    # a step far too small to turn within 2**3 - 1 steps: every tree is complete
    theta = torch.randn(chains, P, dtype=torch.float64, generator=gen)
    _, _, stats, eps = nuts_chains(_gaussian(mu, sd), theta, 20, 0, max_depth=3, step_size=1e-4, generator=gen)
    assert eps == 1e-4                                                # no tuning: the step size is the caller's
    assert np.all(stats['depth'].numpy() == 3) and np.all(stats['n_steps'].numpy() == 7)
    assert float(stats['accept'].min()) > 0.999
    # a step far too large: the first leaf's energy error exceeds the limit, the chain stays where it is
    theta = torch.randn(chains, P, dtype=torch.float64, generator=gen)
    start = theta.clone()
    kept, _, stats, _ = nuts_chains(_gaussian(mu, 1e-3 * sd), theta, 4, 0, max_depth=5, step_size=10.0, generator=gen)
    assert bool(stats['diverging'].all()) and np.all(stats['n_steps'].numpy() == 1)
    assert torch.equal(kept[:, -1], start)


def test_nuts_checkpoint_indices():
    """The balanced sub-subtrees that end at leaf n start at the leaves whose checkpoints _ckpt_range names."""
    from sakkara_b200.sampling import _ckpt_range
    slot_leaf = {}
    for leaf in range(256):
        idx_min, idx_max = _ckpt_range(leaf)
        if leaf % 2 == 0:
            slot_leaf[idx_max] = leaf                                # an even leaf opens a sub-subtree
            assert idx_min == idx_max + 1                            # ... and closes none
            continue
        starts = [slot_leaf[i] for i in range(idx_min, idx_max + 1)]
        # expected: for every k >= 1 with 2**k dividing leaf + 1, the aligned block of 2**k leaves ending here
        expected = []
        k = 1
        while (leaf + 1) % (1 << k) == 0:
            expected.append(leaf + 1 - (1 << k))
            k += 1
        assert sorted(starts) == sorted(expected)


def test_nuts_on_the_oracle_recovers_the_regression():
    from sakkara_b200.sampling import nuts_chains
    rng = np.random.default_rng(5)
    n, groups = 200, 3
    code = rng.integers(0, groups, n)
    k = np.array([1.0, -1.0, 0.5])
    x = rng.standard_normal(n)
    y = k[code] * x + 0.3 + 0.5 * rng.standard_normal(n)
    df = pd.DataFrame({'x': x, 'y': y, 'group': code})
    plan = build_model((df, readme_likelihood(df))).plan

    def evaluate(q, logp_out, grad_out):
        for c in range(q.shape[0]):
            lp, g = logp_dlogp(plan, q[c].numpy())
            logp_out[c] = float(lp)
            grad_out[c] = torch.from_numpy(np.asarray(g, dtype=np.float64))

    chains, draws, tune = 4, 100, 120
    gen = torch.Generator().manual_seed(4)
    theta = 0.1 * torch.randn(chains, plan.n_params, dtype=torch.float64, generator=gen)
    kept, kept_logp, stats, eps = nuts_chains(evaluate, theta, draws, tune, max_depth=6, generator=gen)
    assert int(stats['diverging'].sum()) <= 2
    sl = plan.block_slices()
    order = pd.factorize(df['group'])[1]
    assert np.allclose(kept[:, :, sl['k']].mean((0, 1)).numpy(), k[np.asarray(order)], atol=0.15)
    assert abs(float(kept[:, :, sl['intercept']].mean()) - 0.3) < 0.12
    assert abs(float(kept[:, :, sl['sigma_est']].exp().mean()) - 0.5) < 0.08


def test_warmup_windows_follow_stans_schedule():
    from sakkara_b200.sampling import warmup_windows
    assert warmup_windows(1000) == [(75, 100), (100, 150), (150, 250), (250, 450), (450, 950)]
    assert warmup_windows(500) == [(75, 100), (100, 150), (150, 250), (250, 450)]
    assert warmup_windows(100) == [(15, 90)]
    assert warmup_windows(10) == []
    for tune in (20, 57, 149, 150, 151, 333, 2000):
        w = warmup_windows(tune)
        assert w[0][0] >= 0 and w[-1][1] <= tune
        assert all(a < b for a, b in w) and all(w[i][1] == w[i + 1][0] for i in range(len(w) - 1))


def test_diagnostics_on_known_processes():
    from sakkara_b200.sampling import ess_bulk, split_rhat
    rng = np.random.default_rng(0)
    iid = rng.standard_normal((8, 1000))
    assert abs(split_rhat(iid) - 1.0) < 0.01
    assert 0.8 * iid.size < ess_bulk(iid) < 1.25 * iid.size
    shifted = iid + np.arange(8)[:, None]                       # chains that sit in different places
    assert split_rhat(shifted) > 1.5 and ess_bulk(shifted) < 50
    rho = 0.9                                                   # AR(1): ESS = N (1 - rho) / (1 + rho)
    ar = np.empty((8, 4000))
    ar[:, 0] = rng.standard_normal(8)
    eps = rng.standard_normal(ar.shape) * np.sqrt(1 - rho * rho)
    for t in range(1, ar.shape[1]):
        ar[:, t] = rho * ar[:, t - 1] + eps[:, t]
    expected = ar.size * (1 - rho) / (1 + rho)
    assert 0.75 * expected < ess_bulk(ar) < 1.3 * expected
    assert split_rhat(ar) < 1.05
    drift = iid + np.linspace(0, 3, 1000)[None, :]              # every chain drifts: the halves disagree
    assert split_rhat(drift) > 1.15


# ---------------------------------------------------------------------------------------------------------------
# chain-sharded run: two processes (gloo), every rank its own chains, adaptation pooled


def _sharded_nuts_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    from sakkara_b200.sampling import chain_pool, nuts_chains
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pool, chains, r = chain_pool(16)
        assert chains == 8 and r == rank and type(pool).__name__ == 'DistributedChainPool'
        P = 5
        sd = torch.linspace(0.2, 3.0, P, dtype=torch.float64)
        mu = torch.arange(P, dtype=torch.float64)
        gen = torch.Generator().manual_seed(100 + rank)
        theta = 0.1 * torch.randn(chains, P, dtype=torch.float64, generator=gen)
        kept, _, stats, eps = nuts_chains(_gaussian(mu, sd), theta, 250, 250, max_depth=7, generator=gen, pool=pool)
        torch.save({'kept': kept, 'eps': eps, 'accept': float(stats['accept'].mean())}, os.path.join(out, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_chain_sharded_nuts_adapts_as_one_ensemble(tmp_path):
    import os
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_sharded_nuts_worker, args=(world, 29100 + os.getpid() % 700, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), f'r{r}.pt')) for r in range(world)]
    assert res[0]['eps'] == res[1]['eps']                        # one step size for the whole ensemble
    assert not torch.equal(res[0]['kept'], res[1]['kept'])       # ... but different chains
    k = torch.cat([r['kept'] for r in res]).numpy()
    P = k.shape[2]
    sd = np.linspace(0.2, 3.0, P)
    assert np.all(np.abs(k.mean((0, 1)) - np.arange(P)) < 0.1 * sd)
    assert np.allclose(k.reshape(-1, P).std(0) / sd, 1.0, atol=0.06)
    assert all(abs(r['accept'] - 0.8) < 0.1 for r in res)


def test_samplers_run_host_resident_for_models_that_are_a_sum_of_plans():
    """sample_nuts / sample_hmc with an evaluator that is not one device plan (no ``skk``): chains in host memory, one
    batched call with host buffers per leapfrog step -- here a stand-in evaluator with a Gaussian target."""
    from types import SimpleNamespace
    from sakkara_b200.sampling import sample_hmc, sample_nuts
    P = 4
    mu, sd = np.arange(P, dtype=np.float64), np.array([0.5, 1.0, 2.0, 4.0])
    calls = []

    class Evaluator:
        size = P
        dtype = np.dtype('float64')

        def __call__(self, theta):
            calls.append(theta.shape)
            z = (theta - mu) / sd
            return -0.5 * (z * z).sum(1), -z / sd

        def close(self):
            calls.append('closed')

    model = SimpleNamespace(plan=SimpleNamespace(blocks=[SimpleNamespace(value_name='v', name='v', offset=0, size=P,
                                                                         transform='identity')]),
                            logp_dlogp_function=lambda **kw: Evaluator())
    res = sample_nuts(model, chains=16, draws=200, tune=200, seed=3)
    assert res.draws.shape == (16, 200, P) and calls[-1] == 'closed' and set(calls[:-1]) == {(16, P)}
    assert np.allclose(res.draws.reshape(-1, P).std(0) / sd, 1.0, atol=0.08)
    assert np.all(np.abs(res.draws.mean((0, 1)) - mu) < 0.15 * sd)
    assert res.constrained(model.plan)['v'].shape == (16, 200, P)
    res = sample_hmc(model, chains=8, draws=50, tune=80, n_leapfrog=5, step_size=0.2, seed=1)
    assert res.draws.shape == (8, 50, P) and 0.3 < res.accept_rate.mean() <= 1.0
