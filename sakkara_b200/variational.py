"""
Mean-field ADVI on top of the value-and-gradient evaluator.

The reference fits mini-batched models with ``pm.fit()`` inside ``with build(df, MinibatchLikelihood(...))``
(``/root/reference/sakkara/model/composable/hierarchical/likelihood.py:65-89``, ``sakkara/model/minibatch.py:9-33``):
PyMC's ADVI draws theta from the variational family, asks the model for logp and its gradient -- for a
mini-batched likelihood an unbiased estimate from a fresh random subset of the rows, scaled by ``total_size`` -- and
takes a stochastic gradient step on the ELBO.  What PyMC asks per step is what ``model.logp_dlogp_function()`` returns
here (``MinibatchValueGrad`` for a mini-batched model, the full-data evaluators otherwise); this module is the loop
around it for use without PyMC, with PyMC's defaults: a fully factorised Gaussian on the unconstrained scale,
``sigma = softplus(rho)``, one Monte-Carlo draw per step, the ``adagrad_window`` optimiser.

    approx = fit_advi(model, n=20_000)            # model: a built B200Model (or an evaluator f(theta) -> (logp, dlogp))
    approx.mean, approx.sd                         # of theta, unconstrained scale
    approx.sample(1000)                            # draws [1000, P];  approx.constrained(plan, draws)
"""
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np


def _softplus(x):
    return np.logaddexp(0.0, x)


def _sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


@dataclass
class MeanFieldApproximation:
    mu: np.ndarray                # [P]
    rho: np.ndarray               # [P]; sd = softplus(rho)
    elbo: np.ndarray              # [n] stochastic ELBO estimates, one per step
    seed: int = 0

    @property
    def mean(self) -> np.ndarray:
        return self.mu

    @property
    def sd(self) -> np.ndarray:
        return _softplus(self.rho)

    def sample(self, draws: int = 1000, seed: Optional[int] = None) -> np.ndarray:
        rng = np.random.default_rng(self.seed + 1 if seed is None else seed)
        return self.mu + self.sd * rng.standard_normal((draws, self.mu.size))

    def constrained(self, plan, draws: np.ndarray) -> Dict[str, np.ndarray]:
        """Draws per block on the constrained scale (exp for log-transformed blocks)."""
        out = {}
        for b in plan.blocks:
            v = draws[..., b.offset:b.offset + b.size]
            out[b.name] = np.exp(v) if b.transform == 'log' else v
        return out


class AdagradWindow:
    """PyMC's default optimiser for variational inference: Adagrad over a window of the last ``n_win`` gradients."""

    def __init__(self, size: int, learning_rate: float = 1e-3, n_win: int = 10, epsilon: float = 0.1):
        self.lr, self.n_win, self.eps = learning_rate, n_win, epsilon
        self.window = np.zeros((n_win, size))
        self.i = 0

    def step(self, grad: np.ndarray) -> np.ndarray:
        """The update to ADD to the parameters for the gradient of an objective that is MAXIMISED."""
        self.window[self.i % self.n_win] = grad * grad
        self.i += 1
        return self.lr * grad / (np.sqrt(self.window.sum(0)) + self.eps)


def fit_advi(model, n: int = 10_000, learning_rate: float = 1e-3, n_win: int = 10, seed: int = 0, dtype: str = 'float64',
             device: Optional[int] = None, start: Optional[np.ndarray] = None, start_sigma: float = 1.0,
             callback=None) -> MeanFieldApproximation:
    """
    Mean-field ADVI (Kucukelbir et al. 2017) with the reparameterisation gradient, one draw per step:

        theta = mu + softplus(rho) * eps,  eps ~ N(0, I)
        ELBO  = E[logp(theta)] + sum log softplus(rho)                                  (+ const)
        d/dmu = dlogp(theta),   d/drho = (dlogp(theta) * eps + 1 / sd) * sigmoid(rho)

    :param model: a built model (``model.logp_dlogp_function()`` is asked for the evaluator: the mini-batched one for
        a ``MinibatchLikelihood``) or the evaluator itself, ``f(theta[P]) -> (logp, dlogp[P])`` with ``f.size``.
    :param n: number of steps (``pm.fit(n=...)``); ``learning_rate`` / ``n_win``: ``pm.adagrad_window``'s.
    :param start: initial mean (default zeros, PyMC's initial point for these families); ``start_sigma``: initial sd.
    :param callback: called as ``callback(step, mu, rho, elbo)`` every step (PyMC's ``callbacks``).
    """
    own = hasattr(model, 'logp_dlogp_function')
    if own:
        # (a mini-batched model draws its rows per call: seeded like the variational draws)
        options = {'seed': seed} if getattr(model.plan, 'minibatch', None) is not None else {}
        f = model.logp_dlogp_function(dtype=dtype, device=device, **options)
    else:
        f = model
    size = int(f.size)
    rng = np.random.default_rng(seed)
    mu = np.zeros(size) if start is None else np.array(start, dtype=np.float64)
    rho = np.full(size, float(np.log(np.expm1(start_sigma))))
    opt_mu, opt_rho = AdagradWindow(size, learning_rate, n_win), AdagradWindow(size, learning_rate, n_win)
    elbo = np.empty(n)
    try:
        for step in range(n):
            sd = _softplus(rho)
            eps = rng.standard_normal(size)
            logp, grad = f((mu + sd * eps).astype(dtype))
            grad = np.asarray(grad, dtype=np.float64)
            elbo[step] = float(logp) + float(np.sum(np.log(sd)))
            if not np.isfinite(elbo[step]) or not np.all(np.isfinite(grad)):
                continue                        # a draw outside the support of the numerics: no step
            g_mu = grad
            g_rho = (grad * eps + 1.0 / sd) * _sigmoid(rho)
            mu = mu + opt_mu.step(g_mu)
            rho = rho + opt_rho.step(g_rho)
            if callback is not None:
                callback(step, mu, rho, elbo[step])
    finally:
        if own and hasattr(f, 'close'):
            f.close()
    return MeanFieldApproximation(mu=mu, rho=rho, elbo=elbo, seed=seed)
