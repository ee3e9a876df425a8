"""Many-chain HMC on the batch evaluator: recovers the coefficients of a small hierarchical regression."""
import numpy as np
import pandas as pd
import pytest

from helpers import build_model
from model_zoo import readme_likelihood

pytestmark = pytest.mark.gpu


def test_hmc_recovers_group_slopes():
    from sakkara_b200.sampling import sample_hmc
    rng = np.random.default_rng(3)
    n, groups = 4000, 4
    code = rng.integers(0, groups, n)
    k = np.array([1.0, -1.0, 0.5, 2.0])
    x = rng.standard_normal(n)
    y = k[code] * x + 0.3 + 0.5 * rng.standard_normal(n)
    df = pd.DataFrame({'x': x, 'y': y, 'group': code})
    model = build_model((df, readme_likelihood(df)))
    res = sample_hmc(model, chains=64, draws=300, tune=400, n_leapfrog=12, seed=1)
    assert res.draws.shape == (64, 300, model.plan.n_params)
    assert 0.4 < res.accept_rate.mean() <= 1.0
    post = res.constrained(model.plan)
    order = pd.factorize(df['group'])[1]                      # members in first-appearance order
    assert np.allclose(post['k'].mean((0, 1)), k[np.asarray(order)], atol=0.06)
    assert abs(post['intercept'].mean() - 0.3) < 0.05
    assert abs(post['sigma_est'].mean() - 0.5) < 0.05
