"""Many-chain HMC and NUTS on the batch evaluator: the fused integrator kernel, and both samplers recover the
coefficients of a small hierarchical regression."""
import numpy as np
import pandas as pd
import pytest

from helpers import build_model
from model_zoo import readme_likelihood

pytestmark = pytest.mark.gpu


def _regression(n=4000, seed=3):
    rng = np.random.default_rng(seed)
    groups = 4
    code = rng.integers(0, groups, n)
    k = np.array([1.0, -1.0, 0.5, 2.0])
    x = rng.standard_normal(n)
    y = k[code] * x + 0.3 + 0.5 * rng.standard_normal(n)
    df = pd.DataFrame({'x': x, 'y': y, 'group': code})
    return df, k[np.asarray(pd.factorize(df['group'])[1])]       # members in first-appearance order


def _check_posterior(res, model, k_true, chains, draws):
    assert res.draws.shape == (chains, draws, model.plan.n_params)
    post = res.constrained(model.plan)
    assert np.allclose(post['k'].mean((0, 1)), k_true, atol=0.06)
    assert abs(post['intercept'].mean() - 0.3) < 0.05
    assert abs(post['sigma_est'].mean() - 0.5) < 0.05


@pytest.mark.parametrize('fused', [True, False])
def test_hmc_recovers_group_slopes(fused):
    from sakkara_b200.sampling import sample_hmc
    df, k_true = _regression()
    model = build_model((df, readme_likelihood(df)))
    res = sample_hmc(model, chains=64, draws=300, tune=400, n_leapfrog=12, seed=1, fused_leapfrog=fused)
    assert 0.4 < res.accept_rate.mean() <= 1.0
    _check_posterior(res, model, k_true, 64, 300)


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_leapfrog_kernel_matches_torch(dtype):
    """skk_leapfrog against the torch statement of the same update (sampling.kick_drift_torch), all three (a, b) forms."""
    import torch
    from sakkara_b200.sampling import kick_drift_torch
    df, _ = _regression(n=500)
    model = build_model((df, readme_likelihood(df)))
    chains = 37
    f = model.logp_dlogp_function(dtype=dtype, max_batch=chains)
    skk, P = f.skk, f.size
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    dev = torch.device('cuda', 0)
    stream = torch.cuda.ExternalStream(skk.stream, device=dev)
    gen = torch.Generator(device=dev).manual_seed(5)
    stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(stream):
        q, r, g = (torch.randn(chains, P, dtype=tdt, device=dev, generator=gen) for _ in range(3))
        eps = 0.1 * torch.randn(chains, 1, dtype=tdt, device=dev, generator=gen)      # both signs
        inv_mass = torch.rand(P, dtype=tdt, device=dev, generator=gen) + 0.5
        for a, b in ((0.5, 1.0), (0.5, 0.0), (1.0, 1.0)):
            q_ref, r_ref = q.clone(), r.clone()
            kick_drift_torch(q_ref, r_ref, g, eps, inv_mass, a, b)
            q_got, r_got = q.clone(), r.clone()
            skk.leapfrog(q_got.data_ptr(), r_got.data_ptr(), g.data_ptr(), eps.data_ptr(), inv_mass.data_ptr(), chains, a, b)
            tol = 1e-13 if dtype == 'float64' else 1e-5
            assert torch.allclose(r_got, r_ref, rtol=tol, atol=tol)
            assert torch.allclose(q_got, q_ref, rtol=tol, atol=tol)
            if b == 0.0:
                assert torch.equal(q_got, q)                      # the kick leaves the positions alone
    torch.cuda.synchronize()
    with pytest.raises(Exception):
        skk.leapfrog(0, r.data_ptr(), g.data_ptr(), eps.data_ptr(), inv_mass.data_ptr(), chains, 0.5, 1.0)
    f.close()


def test_nuts_recovers_group_slopes():
    from sakkara_b200.sampling import sample_nuts
    df, k_true = _regression()
    model = build_model((df, readme_likelihood(df)))
    chains, draws = 64, 200
    res = sample_nuts(model, chains=chains, draws=draws, tune=300, max_depth=7, seed=2)
    _check_posterior(res, model, k_true, chains, draws)
    st = res.stats
    assert abs(st['accept'].mean() - 0.8) < 0.12
    assert st['diverging'].mean() < 0.01
    assert st['depth'].max() <= 7 and np.all(st['n_steps'] <= 2 ** st['depth'] - 1)
    summary = res.summary(model.plan)
    assert max(v['r_hat'].max() for v in summary.values()) < 1.05
    assert min(v['ess_bulk'].min() for v in summary.values()) > 0.1 * chains * draws
    # kept logp values are the evaluator's at the kept positions
    f = model.logp_dlogp_function()
    lp, _ = f(res.draws[5, 17])
    assert abs(float(lp) - float(res.logp[5, 17])) < 1e-8 * abs(float(lp))
    f.close()


def test_model_sample_is_nuts_with_pm_sample_arguments():
    df, k_true = _regression(n=1000)
    model = build_model((df, readme_likelihood(df)))
    res = model.sample(draws=100, tune=150, chains=32, target_accept=0.85, random_seed=7)
    assert res.draws.shape == (32, 100, model.plan.n_params) and set(res.stats) >= {'depth', 'diverging', 'energy'}
    assert np.allclose(res.constrained(model.plan)['k'].mean((0, 1)), k_true, atol=0.12)
