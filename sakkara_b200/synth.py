"""
Synthetic workloads of the shapes named in BASELINE.json, produced directly as
:class:`ModelPlan` objects -- the same plans ``build()`` lowers from the corresponding DSL models
(checked on small sizes in tests/test_synth.py), without materialising a 1e8-row DataFrame.

Row blocks are generated from ``np.random.default_rng([seed, config, block])`` so that every rank
of a row-sharded run can produce exactly its own rows.
"""
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from sakkara_b200.plan.model_plan import Block, LikelihoodSpec, ModelPlan, PriorParam, Term


def _const(v) -> PriorParam:
    return PriorParam(const=np.asarray(float(v), dtype=np.float64))


def _link(block: int, index: np.ndarray) -> PriorParam:
    return PriorParam(block=block, index=np.ascontiguousarray(index, dtype=np.int64))


def _zeros_index(n: int) -> np.ndarray:
    """all-zeros int64 gather index without allocating n elements (read-only broadcast view)"""
    return np.broadcast_to(np.zeros(1, dtype=np.int64), (n,))


def _first_appearance(codes: np.ndarray, fixed: Optional[int] = None) -> Tuple[np.ndarray, int, np.ndarray]:
    """
    codes relabelled in first-appearance order (what Group / pd.factorize produce) + first rows.
    ``fixed``: keep the raw codes as member numbers of a group with that many members -- the
    labelling every rank of a row-sharded run must share (a rank sees only some members).
    """
    if fixed is not None:
        codes = np.asarray(codes, dtype=np.int64)
        return codes, int(fixed), None
    dense, members = pd.factorize(codes, sort=False)
    first = np.flatnonzero(~pd.Series(dense).duplicated(keep='first').values)
    return dense.astype(np.int64, copy=False), len(members), first


def _blocks(spec) -> list:
    out, offset = [], 0
    for name, family, size, dims, params in spec:
        transform = 'log' if family in ('HalfNormal', 'Exponential') else 'identity'
        out.append(Block(name=name, family=family, offset=offset, size=size, shape=(size,), dims=dims,
                         transform=transform, params=params))
        offset += size
    return out


# ---------------------------------------------------------------------------------------------
def linreg_columns(n: int, groups: int, seed: int = 1):
    """C2 data: y = k[group] * x + 0.3 + 0.5 eps, k ~ N(0.7, 0.5)."""
    rng = np.random.default_rng([seed, 2])
    k = rng.normal(0.7, 0.5, groups)
    codes = rng.integers(0, groups, n)
    x = rng.standard_normal(n)
    y = k[codes] * x + 0.3 + 0.5 * rng.standard_normal(n)
    return {'x': x, 'y': y, 'group': codes}, {'k': k}


def linreg_plan(cols, groups: Optional[int] = None) -> ModelPlan:
    """README-shaped hierarchical linear regression; theta = [mu_k, log s_k, k(G), intercept, log s_est]."""
    n = len(cols['y'])
    g, G, _ = _first_appearance(cols['group'], groups)
    zg = np.zeros(G, dtype=np.int64)
    blocks = _blocks([
        ('mu_k', 'Normal', 1, ('global',), {'mu': _const(0), 'sigma': _const(1)}),
        ('sigma_k', 'Exponential', 1, ('global',), {'lam': _const(1)}),
        ('k', 'Normal', G, ('group',), {'mu': _link(0, zg), 'sigma': _link(1, zg)}),
        ('intercept', 'Normal', 1, ('global',), {'mu': _const(0), 'sigma': _const(1)}),
        ('sigma_est', 'Exponential', 1, ('global',), {'lam': _const(1)}),
    ])
    terms = [Term(2, g, np.ascontiguousarray(cols['x'], dtype=np.float64)), Term(3, _zeros_index(n), None)]
    lik = LikelihoodSpec('Normal', 'est', np.ascontiguousarray(cols['y'], dtype=np.float64), sigma_block=4)
    return ModelPlan(blocks, terms, lik, n)


# ---------------------------------------------------------------------------------------------
def logistic_columns(n: int, g1: int, per_g1: int, seed: int = 1, block: int = 0, g2_range=None):
    """C3 data: nested levels g2 = g1 * per_g1 + j; Bernoulli(sigmoid(slope[g1] x + icpt[g2])), |eta| <~ 4."""
    par = np.random.default_rng([seed, 3])
    slope = par.normal(0.8, 0.4, g1)
    icpt = par.normal(0.0, 0.8, g1 * per_g1)
    rng = np.random.default_rng([seed, 3, 1 + block])
    lo, hi = (0, g1 * per_g1) if g2_range is None else g2_range
    g2 = rng.integers(lo, hi, n)
    x = rng.standard_normal(n, dtype=np.float32).astype(np.float64)
    eta = np.clip(slope[g2 // per_g1] * x + icpt[g2], -4, 4)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return {'x': x, 'y': y, 'g1': g2 // per_g1, 'g2': g2}, {'slope': slope, 'icpt': icpt}


def logistic_plan(cols, groups: Optional[Tuple[int, int]] = None) -> ModelPlan:
    """
    theta = [mu_slope, ls_slope, slope(G1), mu_mu_icpt, ls_mu_icpt, mu_icpt(G1), ls_icpt, icpt(G2)].
    ``groups=(g1, per_g1)``: fixed labelling g2 = g1 * per_g1 + j (sharded runs).
    """
    n = len(cols['y'])
    if groups is None:
        c1, G1, _ = _first_appearance(cols['g1'])
        c2, G2, first2 = _first_appearance(cols['g2'])
        parent = c1[first2]                              # g1 code of each g2 member
    else:
        G1, G2 = groups[0], groups[0] * groups[1]
        c1, c2 = np.asarray(cols['g1'], dtype=np.int64), np.asarray(cols['g2'], dtype=np.int64)
        parent = np.arange(G2, dtype=np.int64) // groups[1]
    z1, z2 = np.zeros(G1, dtype=np.int64), np.zeros(G2, dtype=np.int64)
    blocks = _blocks([
        ('mu_slope', 'Normal', 1, ('global',), {'mu': _const(0), 'sigma': _const(1)}),
        ('sigma_slope', 'HalfNormal', 1, ('global',), {'sigma': _const(1)}),
        ('slope', 'Normal', G1, ('g1',), {'mu': _link(0, z1), 'sigma': _link(1, z1)}),
        ('mu_mu_icpt', 'Normal', 1, ('global',), {'mu': _const(0), 'sigma': _const(1)}),
        ('sigma_mu_icpt', 'HalfNormal', 1, ('global',), {'sigma': _const(2)}),
        ('mu_icpt', 'Normal', G1, ('g1',), {'mu': _link(3, z1), 'sigma': _link(4, z1)}),
        ('sigma_icpt', 'HalfNormal', 1, ('global',), {'sigma': _const(1)}),
        ('icpt', 'Normal', G2, ('g2',), {'mu': _link(5, parent), 'sigma': _link(6, z2)}),
    ])
    terms = [Term(2, c1, np.ascontiguousarray(cols['x'], dtype=np.float64)), Term(7, c2, None)]
    lik = LikelihoodSpec('Bernoulli', 'likelihood', np.ascontiguousarray(cols['y'], dtype=np.float64))
    return ModelPlan(blocks, terms, lik, n)


# ---------------------------------------------------------------------------------------------
def poisson_columns(n: int, a: int, b: int, seed: int = 1, block: int = 0, a_range=None):
    """C4 data: crossed factors A, B drawn independently; y ~ Poisson(exp(0.2 x + a[A] + b[B]))."""
    par = np.random.default_rng([seed, 4])
    ea = par.normal(0.0, 0.3, a)
    eb = par.normal(0.0, 0.3, b)
    rng = np.random.default_rng([seed, 4, 1 + block])
    lo, hi = (0, a) if a_range is None else a_range
    A = rng.integers(lo, hi, n, dtype=np.int32)
    B = rng.integers(0, b, n, dtype=np.int32)
    x = rng.standard_normal(n, dtype=np.float32)
    eta = 0.2 * x + ea[A].astype(np.float32) + eb[B].astype(np.float32)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    return {'x': x.astype(np.float64), 'y': y, 'A': A, 'B': B}, {'a': ea, 'b': eb}


def poisson_plan(cols, groups: Optional[Tuple[int, int]] = None) -> ModelPlan:
    """theta = [beta, ls_a, a(A), ls_b, b(B)] (model written obs-term first: beta*x + a + b)"""
    n = len(cols['y'])
    ca, GA, _ = _first_appearance(cols['A'], None if groups is None else groups[0])
    cb, GB, _ = _first_appearance(cols['B'], None if groups is None else groups[1])
    za, zb = np.zeros(GA, dtype=np.int64), np.zeros(GB, dtype=np.int64)
    blocks = _blocks([
        ('beta', 'Normal', 1, ('global',), {'mu': _const(0), 'sigma': _const(1)}),
        ('sigma_a', 'HalfNormal', 1, ('global',), {'sigma': _const(1)}),
        ('a', 'Normal', GA, ('A',), {'mu': _const(0), 'sigma': _link(1, za)}),
        ('sigma_b', 'HalfNormal', 1, ('global',), {'sigma': _const(1)}),
        ('b', 'Normal', GB, ('B',), {'mu': _const(0), 'sigma': _link(3, zb)}),
    ])
    terms = [Term(0, _zeros_index(n), np.ascontiguousarray(cols['x'], dtype=np.float64)),
             Term(2, ca, None), Term(4, cb, None)]
    lik = LikelihoodSpec('Poisson', 'likelihood', np.ascontiguousarray(cols['y'], dtype=np.float64))
    return ModelPlan(blocks, terms, lik, n)


def start_point(plan: ModelPlan, batch: int, seed: int = 0, scale: float = 0.1, dtype='float64') -> np.ndarray:
    """theta points near a plausible posterior mode: N(0, scale^2), log-sigma entries ~ N(log 0.5, scale^2)."""
    rng = np.random.default_rng([seed, 99])
    th = rng.normal(0.0, scale, (batch, plan.n_params))
    for b in plan.blocks:
        if b.transform == 'log':
            th[:, b.offset:b.offset + b.size] += np.log(0.5)
    return th.astype(dtype)
