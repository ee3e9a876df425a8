"""
TEST INFRASTRUCTURE -- ctypes wrapper of oracle/logp_c.c: logp + dlogp of a ModelPlan with the
likelihood evaluated by the fused OpenMP loop and the (tiny) prior part by oracle/logp_numpy.py.
Used by bench.py as the multi-core CPU baseline and by tests as a second CPU implementation.
"""
import ctypes as C
import dataclasses
import os

import numpy as np
from scipy.special import gammaln

from oracle import build_c_oracle
from oracle.logp_numpy import logp_dlogp

MAX_TERMS = 8


class GlmDesc(C.Structure):
    _fields_ = [('family', C.c_int32), ('n_terms', C.c_int32), ('n_rows', C.c_int64), ('n_params', C.c_int64),
                ('index', C.c_void_p * MAX_TERMS), ('x', C.c_void_p * MAX_TERMS), ('base', C.c_int64 * MAX_TERMS),
                ('y', C.c_void_p), ('offset', C.c_void_p), ('sigma', C.c_double), ('sigma_theta_index', C.c_int64)]


_lib = None


def _library():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c_oracle.build())
        for name, real in (('skk_oracle_likelihood_f64', C.c_double), ('skk_oracle_likelihood_f32', C.c_float)):
            fn = getattr(_lib, name)
            fn.argtypes = [C.POINTER(GlmDesc), C.c_void_p, C.c_void_p, C.c_int]
            fn.restype = C.c_double
        _lib.skk_oracle_max_threads.restype = C.c_int
    return _lib


def priors_only(plan):
    """Copy of the plan without observation rows: evaluating it gives the prior part."""
    from sakkara_b200.plan.model_plan import Term
    lik = dataclasses.replace(plan.likelihood, y=plan.likelihood.y[:0],
                              offset=None if plan.likelihood.offset is None else plan.likelihood.offset[:0])
    terms = [Term(t.block, t.index[:0], None if t.x is None else t.x[:0]) for t in plan.terms]
    return dataclasses.replace(plan, terms=terms, likelihood=lik, n_rows=0, logp_const=0.0, n_masked_rows=0)


class COracle:
    """Holds the converted columns of a plan so that repeated evaluations only pay for the loop."""

    def __init__(self, plan, dtype='float64', threads: int = 0):
        self.lib = _library()
        self.plan = plan
        self.prior_plan = priors_only(plan)
        self.dtype = np.dtype(dtype)
        self.threads = threads or (os.cpu_count() or 1)
        lik = plan.likelihood
        self.keep = []
        d = GlmDesc()
        d.family = {'Normal': 0, 'Bernoulli': 1, 'Poisson': 2, 'NegativeBinomial': 3, 'Gamma': 4}[lik.family]
        d.n_terms = len(plan.terms)
        d.n_rows = plan.n_rows
        d.n_params = plan.n_params
        for t, term in enumerate(plan.terms):
            idx = np.ascontiguousarray(term.index, dtype=np.int32)
            self.keep.append(idx)
            d.index[t] = idx.ctypes.data
            if term.x is not None:
                x = np.ascontiguousarray(term.x, dtype=self.dtype)
                self.keep.append(x)
                d.x[t] = x.ctypes.data
            d.base[t] = plan.term_base(term)
        y = np.ascontiguousarray(lik.y, dtype=self.dtype)
        self.keep.append(y)
        d.y = y.ctypes.data
        if lik.offset is not None:
            off = np.ascontiguousarray(lik.offset, dtype=self.dtype)
            self.keep.append(off)
            d.offset = off.ctypes.data
        d.sigma_theta_index = -1
        if lik.family == 'Normal' and lik.sigma_block is not None:
            d.sigma_theta_index = plan.blocks[lik.sigma_block].offset + lik.sigma_index
        self.desc = d
        self.const = -float(np.sum(gammaln(lik.y + 1.0))) if lik.family == 'Poisson' else 0.0
        if lik.family == 'NegativeBinomial':
            a = float(lik.sigma_const)
            d.sigma = a
            self.const = float(np.sum(gammaln(lik.y + a) - gammaln(lik.y + 1.0))) + plan.n_rows * (a * np.log(a) - float(gammaln(a)))
        if lik.family == 'Gamma':
            a = float(lik.sigma_const)
            d.sigma = a
            self.const = (a - 1.0) * float(np.sum(np.log(lik.y))) + plan.n_rows * (a * np.log(a) - float(gammaln(a)))
        self.const += float(getattr(plan, 'logp_const', 0.0))   # rows masked out for NaN observations
        self.fn = self.lib.skk_oracle_likelihood_f64 if self.dtype == np.float64 else self.lib.skk_oracle_likelihood_f32

    def __call__(self, theta):
        theta = np.ascontiguousarray(theta, dtype=self.dtype)
        lik = self.plan.likelihood
        if lik.family == 'Normal':
            self.desc.sigma = float(np.exp(theta[self.desc.sigma_theta_index])) if self.desc.sigma_theta_index >= 0 \
                else lik.sigma_const
        grad = np.zeros(self.plan.n_params, dtype=np.float64)
        logp = self.fn(C.byref(self.desc), theta.ctypes.data, grad.ctypes.data, self.threads)
        lp, gp = logp_dlogp(self.prior_plan, theta.astype(np.float64))
        return logp + self.const + float(lp), grad + gp
