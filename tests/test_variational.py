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


def _reference_minibatch_model():
    """The model of /root/reference/tests/workflows/test_minibatch_vi.py:8-16 on its fixtures (udf / xdf)."""
    from sakkara_b200.model import (DeterministicComponent, DistributionComponent as DC, MinibatchLikelihood, build,
                                    data_components, dist as pm)
    udf = pd.DataFrame({'time': np.arange(30), 'u': np.random.default_rng(100).normal(0, 1, 30)})
    xdf = pd.DataFrame({'g': ['a', 'b'], 'k': [-1, 1]}).merge(udf, how='cross')
    xdf['y'] = xdf['u'] * xdf['k']
    xdf['obs'] = np.arange(len(xdf))
    udc, xdc = data_components(udf, 'time'), data_components(xdf)
    k = DC(pm.Normal, name='k', group='g')
    predicted = DeterministicComponent('predicted', k * udc['u'])
    mbl = MinibatchLikelihood(pm.Normal, observed=xdc['y'], batch_size=10, mu=predicted, sigma=1e-15)
    return build(xdf, mbl, backend='b200')


def test_hierarchical_minibatch_workflow_of_the_reference():
    """The reference's own mini-batch VI workflow (tests/workflows/test_minibatch_vi.py:8-24): ``pm.fit(n=10000,
    random_seed=100)``, 10,000 draws, posterior means of k within 1e-2 of -1 / +1.  Here the loop is fit_advi and the
    evaluator is the oracle on a fresh mini-batch per step, composed exactly like valuegrad.MinibatchValueGrad
    composes the kernels' results (priors + n_rows / batch x likelihood of the drawn rows)."""
    plan = _reference_minibatch_model().plan
    assert plan.minibatch == 10 and plan.n_rows == 60 and list(plan.coords['g']) == ['a', 'b']

    class MiniBatchOracle:
        size = plan.n_params
        rng = np.random.default_rng(100)

        def __call__(self, theta):
            rows = np.sort(self.rng.integers(0, plan.n_rows, plan.minibatch))
            prior_l, prior_g = logp_dlogp(plan.take_rows([]), theta)
            sub_l, sub_g = logp_dlogp(plan.take_rows(rows), theta)
            scale = plan.n_rows / plan.minibatch
            return prior_l + scale * (sub_l - prior_l), prior_g + scale * (sub_g - prior_g)

    approx = fit_advi(MiniBatchOracle(), n=10000, seed=100)
    k = approx.constrained(plan, approx.sample(10000, seed=100))['k']
    assert abs(k[:, 0].mean() + 1.0) < 1e-2 and abs(k[:, 1].mean() - 1.0) < 1e-2
