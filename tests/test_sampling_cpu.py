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
