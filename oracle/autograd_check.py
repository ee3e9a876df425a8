"""
TEST INFRASTRUCTURE -- independent second opinion on ``oracle/logp_numpy.py``.

Evaluates the *recorded symbolic graph* of a built model (``sakkara_b200.model.sym.Model``: the
distribution calls with their parameter expressions, exactly what the components handed to the
backend) with torch fp64 tensors and ``torch.distributions`` log-densities, and differentiates
with autograd.  It shares no code with the plan lowering nor with the closed-form gradients of
the NumPy oracle, so agreement between the two pins the lowering (theta layout, gather
composition, linearisation) and the hand-derived derivatives of SURVEY.md §A at once.
"""

import numpy as np
import torch

from sakkara_b200.model import sym


def _evaluate(node, env):
    if isinstance(node, sym.RandomVar):
        return env[id(node)]
    if isinstance(node, sym.DeterministicVar):
        return _evaluate(node.var, env)
    if isinstance(node, sym.DataVar):
        return torch.as_tensor(np.asarray(node.values, dtype=np.float64))
    if isinstance(node, sym.Gather):
        base = _evaluate(node.base, env)
        return base[tuple(torch.as_tensor(np.ascontiguousarray(i), dtype=torch.long) for i in node.indices)]
    if isinstance(node, sym.Masked):
        value = _evaluate(node.value, env)
        fill = _evaluate(node.fill, env)
        return torch.where(torch.as_tensor(node.mask), fill, value)
    if isinstance(node, sym.Concat):
        return torch.cat([torch.as_tensor(_evaluate(piece, env), dtype=torch.float64).reshape(-1) for piece in node.pieces])
    if isinstance(node, sym.BinOp):
        left, right = _evaluate(node.left, env), _evaluate(node.right, env)
        return sym.BinOp.OPS[node.op](left, right)
    if isinstance(node, sym.UnaryOp):
        arg = _evaluate(node.arg, env)
        return {'neg': torch.neg, 'exp': torch.exp, 'sigmoid': torch.sigmoid}[node.op](arg)
    return torch.as_tensor(np.asarray(node, dtype=np.float64))


def logp_dlogp_autograd(recorded: sym.Model, theta: np.ndarray):
    """logp and gradient w.r.t. the raveled unconstrained theta (free variables in creation order)."""
    theta_t = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
    env = {}
    offset = 0
    total = torch.zeros((), dtype=torch.float64)
    for rv in recorded.free_RVs:
        raw = theta_t[offset:offset + rv.size].reshape(rv.shape)
        offset += rv.size
        params = {k: _evaluate(v, env) for k, v in rv.params.items()}
        if rv.family == 'Normal':
            env[id(rv)] = raw
            total = total + torch.distributions.Normal(params['mu'], params['sigma']).log_prob(raw).sum()
        elif rv.family == 'HalfNormal':
            value = torch.exp(raw)
            env[id(rv)] = value
            total = total + (torch.distributions.HalfNormal(params['sigma']).log_prob(value) + raw).sum()
        elif rv.family == 'Exponential':
            value = torch.exp(raw)
            env[id(rv)] = value
            total = total + (torch.distributions.Exponential(params['lam']).log_prob(value) + raw).sum()
        elif rv.family in ('HalfCauchy', 'Gamma', 'InverseGamma', 'LogNormal', 'HalfStudentT'):
            value = torch.exp(raw)
            env[id(rv)] = value
            if rv.family == 'HalfCauchy':
                lp = torch.distributions.HalfCauchy(params['beta']).log_prob(value)
            elif rv.family == 'Gamma':
                lp = torch.distributions.Gamma(params['alpha'], params['beta']).log_prob(value)
            elif rv.family == 'InverseGamma':
                lp = torch.distributions.InverseGamma(params['alpha'], params['beta']).log_prob(value)
            elif rv.family == 'LogNormal':
                lp = torch.distributions.LogNormal(params['mu'], params['sigma']).log_prob(value)
            else:   # a half Student-t is a Student-t folded at 0: twice the density on the positive side
                lp = torch.distributions.StudentT(params['nu'], torch.zeros(()), params['sigma']).log_prob(value) + np.log(2.0)
            total = total + (lp + raw).sum()
        elif rv.family == 'Cauchy':
            env[id(rv)] = raw
            total = total + torch.distributions.Cauchy(params['alpha'], params['beta']).log_prob(raw).sum()
        elif rv.family == 'Laplace':
            env[id(rv)] = raw
            total = total + torch.distributions.Laplace(params['mu'], params['b']).log_prob(raw).sum()
        else:
            raise NotImplementedError(rv.family)
    assert offset == theta_t.numel()
    for rv in recorded.observed_RVs:
        y = _evaluate(rv.observed, env)
        params = {k: _evaluate(v, env) for k, v in rv.params.items()}
        if rv.family == 'Normal':
            total = total + torch.distributions.Normal(params['mu'], params['sigma']).log_prob(y).sum()
        elif rv.family == 'Bernoulli':
            if 'logit_p' in params:
                d = torch.distributions.Bernoulli(logits=params['logit_p'])
            else:
                d = torch.distributions.Bernoulli(probs=params['p'])
            total = total + d.log_prob(y).sum()
        elif rv.family == 'Poisson':
            total = total + torch.distributions.Poisson(params['mu']).log_prob(y).sum()
        elif rv.family == 'StudentT':
            total = total + torch.distributions.StudentT(torch.as_tensor(params['nu'], dtype=torch.float64), params['mu'],
                                                         params['sigma']).log_prob(y).sum()
        elif rv.family == 'NegativeBinomial':
            # torch: total_count failures r and success probability p of the counted event; mean r p / (1 - p)
            mu, alpha = params['mu'], torch.as_tensor(params['alpha'], dtype=torch.float64)
            total = total + torch.distributions.NegativeBinomial(alpha, probs=mu / (mu + alpha)).log_prob(y).sum()
        elif rv.family == 'Gamma':
            alpha = torch.as_tensor(params['alpha'], dtype=torch.float64)
            total = total + torch.distributions.Gamma(alpha, params['beta']).log_prob(y).sum()
        else:
            raise NotImplementedError(rv.family)
    total.backward()
    return float(total.detach()), theta_t.grad.numpy().copy()
