"""Mean-field ADVI (sakkara_b200/variational.py) on CPU: exact Gaussian targets, a mini-batch-style noisy evaluator, and
the oracle of a small regression as the evaluator (the GPU evaluators are drop-ins for these callables)."""
import numpy as np
import pandas as pd

from helpers import build_model
from model_zoo import readme_likelihood
from oracle.logp_numpy import logp_dlogp
from sakkara_b200.variational import AdagradWindow, fit_advi


class Gaussian:
    def __init__(self, mean, sd, noise=0.0, seed=0):
        self.mean, self.sd, self.size = np.asarray(mean, float), np.asarray(sd, float), len(mean)
        self.noise, self.rng, self.calls = noise, np.random.default_rng(seed), 0

    def __call__(self, theta):
        self.calls += 1
        z = (theta - self.mean) / self.sd
        grad = -z / self.sd
        if self.noise:                         # an unbiased, noisy gradient: what a mini-batch gives
            grad = grad + self.noise * self.rng.standard_normal(self.size)
        return -0.5 * float(z @ z), grad


def test_advi_recovers_a_factorised_gaussian():
    target = Gaussian([1.0, -2.0, 0.5], [0.3, 1.5, 0.05])
    approx = fit_advi(target, n=6000, learning_rate=0.02, seed=1)
    assert target.calls == 6000 and approx.elbo.shape == (6000,)
    assert np.allclose(approx.mean, target.mean, atol=0.05 * target.sd + 0.01)
    assert np.allclose(approx.sd / target.sd, 1.0, atol=0.12)
    assert approx.elbo[-500:].mean() > approx.elbo[:500].mean()
    draws = approx.sample(4000)
    assert draws.shape == (4000, 3) and np.allclose(draws.std(0) / approx.sd, 1.0, atol=0.05)


def test_advi_with_noisy_gradients_still_finds_the_mean():
    target = Gaussian([0.7, -0.3], [1.0, 0.4], noise=0.5, seed=3)
    approx = fit_advi(target, n=8000, learning_rate=0.01, seed=2)
    assert np.allclose(approx.mean, target.mean, atol=0.12)


def test_adagrad_window_follows_pymc():
    opt = AdagradWindow(2, learning_rate=0.1, n_win=3, epsilon=0.1)
    g = np.array([1.0, -2.0])
    steps = [opt.step(g) for _ in range(5)]
    # the accumulator holds the last three squared gradients
    for k, s in enumerate(steps):
        acc = min(k + 1, 3) * g * g
        assert np.allclose(s, 0.1 * g / (np.sqrt(acc) + 0.1))


def test_advi_on_the_regression_posterior():
    rng = np.random.default_rng(5)
    n, groups = 400, 3
    code = rng.integers(0, groups, n)
    k = np.array([1.0, -1.0, 0.5])
    x = rng.standard_normal(n)
    df = pd.DataFrame({'x': x, 'y': k[code] * x + 0.3 + 0.5 * rng.standard_normal(n), 'group': code})
    plan = build_model((df, readme_likelihood(df))).plan

    class Oracle:
        size = plan.n_params

        def __call__(self, theta):
            return logp_dlogp(plan, np.asarray(theta, dtype=np.float64))

    approx = fit_advi(Oracle(), n=4000, learning_rate=0.02, seed=0, start_sigma=0.1)
    sl = plan.block_slices()
    order = pd.factorize(df['group'])[1]
    post = approx.constrained(plan, approx.sample(2000))
    assert np.allclose(post['k'].mean(0), k[np.asarray(order)], atol=0.15)
    assert abs(post['intercept'].mean() - 0.3) < 0.1
    assert abs(post['sigma_est'].mean() - 0.5) < 0.08
    assert np.all(approx.sd[sl['k']] < 0.2)                 # the data are informative: tight factors
