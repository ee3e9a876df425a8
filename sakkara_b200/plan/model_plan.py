"""
Logical plan: the flat, array-only description of a built model.

A :class:`ModelPlan` is what remains of the component graph after lowering: parameter blocks in
creation order (= layout of theta), for every block its prior with constant or
gathered parameters, the terms of the linear predictor as ``theta_block[index[row]] * x[row]``
and the likelihood family.  Index arrays are the reference's gather indices, in the original row
order; no kernel decision (sorting, segments, tiles) is taken at this level, so the CPU oracle
can evaluate a ``ModelPlan`` independently of everything the GPU path does with it.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

LOG_SQRT_2PI = 0.91893853320467274178032973640562
LOG_SQRT_2_OVER_PI = -0.22579135264472743236309761494744

# likelihood / prior family ids shared with the C ABI (include/skk_b200.h)
FAMILY_IDS = {'Normal': 0, 'Bernoulli': 1, 'Poisson': 2, 'HalfNormal': 3, 'Exponential': 4,
              'HalfCauchy': 5, 'Gamma': 6, 'InverseGamma': 7, 'Cauchy': 8, 'Laplace': 9, 'LogNormal': 10,
              'HalfStudentT': 11}
# likelihood families (SKK_LIK_* of include/skk_b200.h; the first three share FAMILY_IDS' numbers)
LIKELIHOOD_IDS = {'Normal': 0, 'Bernoulli': 1, 'Poisson': 2, 'NegativeBinomial': 3, 'NormalLogSigma': 4, 'StudentT': 5,
                  'Gamma': 6}
XCOL_LOG_SIGMA = -2      # xcol of a term of log(sigma) (SKK_XCOL_LOG_SIGMA)
# prior families: names of the parameters that travel as a / b through the C ABI (include/skk_b200.h)
PRIOR_PARAMS = {'Normal': ('mu', 'sigma'), 'HalfNormal': ('sigma', None), 'Exponential': ('lam', None),
                'HalfCauchy': ('beta', None), 'Gamma': ('alpha', 'beta'), 'InverseGamma': ('alpha', 'beta'),
                'Cauchy': ('alpha', 'beta'), 'Laplace': ('mu', 'b'), 'LogNormal': ('mu', 'sigma'),
                'HalfStudentT': ('nu', 'sigma')}


def prior_normaliser(family: str, a, b):
    """
    theta-independent part of the log-density of the families whose device code evaluates the theta-dependent
    part only (``SKK_PRIOR_HALFCAUCHY`` and later, include/skk_b200.h); elementwise, 0 for the basic families.
    """
    from scipy.special import gammaln
    a = np.asarray(a, dtype=np.float64)
    b = None if b is None else np.asarray(b, dtype=np.float64)
    if family == 'HalfCauchy':
        return np.log(2.0 / np.pi) - np.log(a)
    if family in ('Gamma', 'InverseGamma'):
        return a * np.log(b) - gammaln(a)
    if family == 'Cauchy':
        return -np.log(np.pi) - np.log(b) + 0.0 * a
    if family == 'Laplace':
        return -np.log(2.0 * b) + 0.0 * a
    if family == 'LogNormal':
        return -LOG_SQRT_2PI - np.log(b) + 0.0 * a
    if family == 'HalfStudentT':
        return np.log(2.0) + gammaln(0.5 * (a + 1.0)) - gammaln(0.5 * a) - 0.5 * np.log(a * np.pi * b * b)
    return np.zeros_like(a)


@dataclass
class PriorParam:
    """
    One parameter of a block's prior: a constant, or ``theta[block][index]``, or -- member-wise parameters of a
    ``GroupComponent`` -- per element one or the other: ``block == -1``, ``index`` is the absolute element of
    theta or -1, and ``const[size]`` holds the values of the elements with -1.
    """
    const: Optional[np.ndarray] = None        # float64, scalar or shape (size,)
    block: Optional[int] = None               # source block number
    index: Optional[np.ndarray] = None        # int64[size], element of the source block per element

    @property
    def is_const(self) -> bool:
        return self.block is None

    @property
    def is_mixed(self) -> bool:
        return self.block is not None and self.block < 0


@dataclass
class Block:
    """One free random variable = one contiguous slice of theta."""
    name: str
    family: str                                # a key of PRIOR_PARAMS
    offset: int
    size: int
    shape: Tuple[int, ...]
    dims: Optional[Tuple[str, ...]]
    transform: str                             # 'identity' | 'log'
    params: Dict[str, PriorParam] = field(default_factory=dict)

    @property
    def value_name(self) -> str:
        """Name of PyMC's value variable: transformed blocks are sampled as ``<name>_log__``."""
        return self.name + '_log__' if self.transform == 'log' else self.name


@dataclass
class Term:
    """
    ``eta[row] += theta[block][index[row]] * x[row]`` (``x`` omitted = 1).  ``block == -1``: ``index`` holds
    absolute elements of theta (a member-wise coefficient, ``GroupComponent``, reads several blocks).
    """
    block: int
    index: np.ndarray                          # int64[N]
    x: Optional[np.ndarray] = None             # float64[N]


@dataclass
class LikelihoodSpec:
    family: str                                # a key of LIKELIHOOD_IDS
    name: str
    y: np.ndarray                              # float64[N] (counts / 0-1 stored as floats too)
    offset: Optional[np.ndarray] = None        # float64[N] constant part of eta
    sigma_const: Optional[float] = None        # Normal: fixed sigma ... (NegativeBinomial, Gamma: the constant alpha)
    sigma_block: Optional[int] = None          # ... or a positive (log-transformed) block
    sigma_index: int = 0                       # element of that block
    nu: Optional[float] = None                 # StudentT: degrees of freedom (constant); its sigma is 1 (rows
                                               # standardised) or per row in sigma_rows_*
    # Normal whose sigma differs between rows and depends on theta (sigma per group as a parameter): per row the
    # element of theta that holds log(sigma), or -1 and the constant in sigma_rows_const.  Such a plan is evaluated
    # as one standard plan per distinct sigma (ModelPlan.split_by_sigma)
    sigma_rows_element: Optional[np.ndarray] = None   # int64[N]
    sigma_rows_const: Optional[np.ndarray] = None     # float64[N]


@dataclass
class ModelPlan:
    blocks: List[Block]
    terms: List[Term]
    likelihood: LikelihoodSpec
    n_rows: int
    coords: Dict[str, np.ndarray] = field(default_factory=dict)
    logp_const: float = 0.0                    # constant added to logp: the rows masked out for NaN observations
    n_masked_rows: int = 0                     # how many rows those were (they are not in n_rows / the arrays)
    minibatch: Optional[int] = None            # MinibatchLikelihood: rows per evaluation, drawn from all n_rows; the
                                               # likelihood part is scaled by n_rows / minibatch

    @property
    def n_params(self) -> int:
        return sum(b.size for b in self.blocks)

    def term_base(self, term: 'Term') -> int:
        """Offset in theta that ``term.index`` counts from."""
        return 0 if term.block < 0 else self.blocks[term.block].offset

    def take_rows(self, rows) -> 'ModelPlan':
        """The same model over a subset (or multiset) of the observation rows; ``rows=[]`` leaves the priors."""
        import dataclasses
        rows = np.asarray(rows, dtype=np.int64)
        spec = self.likelihood
        lik = dataclasses.replace(spec, y=spec.y[rows], offset=None if spec.offset is None else spec.offset[rows],
                                  sigma_rows_element=None if spec.sigma_rows_element is None else spec.sigma_rows_element[rows],
                                  sigma_rows_const=None if spec.sigma_rows_const is None else spec.sigma_rows_const[rows])
        if rows.size == 0 and lik.sigma_rows_element is not None:
            # no rows, no likelihood: the priors alone (any sigma will do)
            lik = dataclasses.replace(lik, sigma_rows_element=None, sigma_rows_const=None, sigma_const=1.0)
        terms = [Term(t.block, t.index[rows], None if t.x is None else t.x[rows]) for t in self.terms]
        return dataclasses.replace(self, terms=terms, likelihood=lik, n_rows=int(rows.size), logp_const=0.0,
                                   n_masked_rows=0, minibatch=None)

    def block_of_element(self, element: int) -> int:
        for number, b in enumerate(self.blocks):
            if b.offset <= element < b.offset + b.size:
                return number
        raise IndexError(element)

    def split_by_sigma(self, per_element: bool = False) -> List['ModelPlan']:
        """
        A plan whose Normal sigma differs between rows as plans the kernels evaluate, over disjoint row sets:

        * the rows whose sigma is an element of theta: ONE plan whose kernel gathers log(sigma) per row like a term of
          the linear predictor (``SKK_LIK_NORMAL_LOGSIGMA``; it keeps ``sigma_rows_element``, all >= 0) -- or, with
          ``per_element`` (and always when there is a single such element), one standard plan per element;
        * the rows with constant sigmas: a standardised plan (y, offset and data columns divided by sigma_i, unit
          sigma, sum(-log sigma_i) in the constant).

        Their likelihoods add up to this plan's; the priors are this plan's without rows.
        """
        import dataclasses
        spec = self.likelihood
        if spec.sigma_rows_element is None:
            return [self]
        out = []
        elements = np.unique(spec.sigma_rows_element)
        linked = elements[elements >= 0]
        # (StudentT has no standard plan with sigma as one element of theta: its sigma parameters always go through the
        #  log-sigma kernels, whatever ``per_element`` says)
        if linked.size and (spec.family == 'StudentT' or (linked.size > 1 and not per_element)):
            rows = np.flatnonzero(spec.sigma_rows_element >= 0)
            out.append(self if rows.size == self.n_rows else self.take_rows(rows))
            elements = elements[elements < 0]
        for element in elements:
            rows = np.flatnonzero(spec.sigma_rows_element == element)
            part = self.take_rows(rows)
            lik = part.likelihood
            if element >= 0:
                number = self.block_of_element(int(element))
                lik = dataclasses.replace(lik, sigma_block=number, sigma_index=int(element) - self.blocks[number].offset,
                                          sigma_const=None, sigma_rows_element=None, sigma_rows_const=None)
                part = dataclasses.replace(part, likelihood=lik)
            else:
                weight = 1.0 / lik.sigma_rows_const
                terms = [Term(t.block, t.index, weight.copy() if t.x is None else t.x * weight) for t in part.terms]
                lik = dataclasses.replace(lik, y=lik.y * weight, offset=None if lik.offset is None else lik.offset * weight,
                                          sigma_const=1.0, sigma_block=None, sigma_rows_element=None, sigma_rows_const=None)
                part = dataclasses.replace(part, terms=terms, likelihood=lik,
                                           logp_const=-float(np.sum(np.log(spec.sigma_rows_const[rows]), dtype=np.float64)))
            out.append(part)
        return out

    def block_slices(self) -> Dict[str, slice]:
        return {b.name: slice(b.offset, b.offset + b.size) for b in self.blocks}

    def describe(self) -> str:
        lines = [f'ModelPlan: N={self.n_rows} P={self.n_params} likelihood={self.likelihood.family}']
        for i, b in enumerate(self.blocks):
            ps = ', '.join(f'{k}={"const" if p.is_const else ("member-wise" if p.is_mixed else "block%d[idx]" % p.block)}'
                           for k, p in b.params.items())
            lines.append(f'  block {i}: {b.value_name} [{b.offset}:{b.offset + b.size}] {b.family}({ps})')
        for t in self.terms:
            lines.append(f'  term: {"theta" if t.block < 0 else "block %d" % t.block}[idx] * {"x" if t.x is not None else "1"}')
        return '\n'.join(lines)
