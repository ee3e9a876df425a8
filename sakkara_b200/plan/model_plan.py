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

# likelihood / prior family ids shared with the C ABI (include/skk_b200.h)
FAMILY_IDS = {'Normal': 0, 'Bernoulli': 1, 'Poisson': 2, 'HalfNormal': 3, 'Exponential': 4}
LOG_SQRT_2PI = 0.91893853320467274178032973640562
LOG_SQRT_2_OVER_PI = -0.22579135264472743236309761494744


@dataclass
class PriorParam:
    """One parameter of a block's prior: a constant, or ``theta[block][index]``."""
    const: Optional[np.ndarray] = None        # float64, scalar or shape (size,)
    block: Optional[int] = None               # source block number
    index: Optional[np.ndarray] = None        # int64[size], element of the source block per element

    @property
    def is_const(self) -> bool:
        return self.block is None


@dataclass
class Block:
    """One free random variable = one contiguous slice of theta."""
    name: str
    family: str                                # Normal | HalfNormal | Exponential
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
    """``eta[row] += theta[block][index[row]] * x[row]`` (``x`` omitted = 1)."""
    block: int
    index: np.ndarray                          # int64[N]
    x: Optional[np.ndarray] = None             # float64[N]


@dataclass
class LikelihoodSpec:
    family: str                                # Normal | Bernoulli | Poisson
    name: str
    y: np.ndarray                              # float64[N] (counts / 0-1 stored as floats too)
    offset: Optional[np.ndarray] = None        # float64[N] constant part of eta
    sigma_const: Optional[float] = None        # Normal only: fixed sigma ...
    sigma_block: Optional[int] = None          # ... or a positive (log-transformed) block
    sigma_index: int = 0                       # element of that block


@dataclass
class ModelPlan:
    blocks: List[Block]
    terms: List[Term]
    likelihood: LikelihoodSpec
    n_rows: int
    coords: Dict[str, np.ndarray] = field(default_factory=dict)
    logp_const: float = 0.0                    # constant added to logp: the rows masked out for NaN observations
    n_masked_rows: int = 0                     # how many rows those were (they are not in n_rows / the arrays)

    @property
    def n_params(self) -> int:
        return sum(b.size for b in self.blocks)

    def block_slices(self) -> Dict[str, slice]:
        return {b.name: slice(b.offset, b.offset + b.size) for b in self.blocks}

    def describe(self) -> str:
        lines = [f'ModelPlan: N={self.n_rows} P={self.n_params} likelihood={self.likelihood.family}']
        for i, b in enumerate(self.blocks):
            ps = ', '.join(f'{k}={"const" if p.is_const else "block%d[idx]" % p.block}' for k, p in b.params.items())
            lines.append(f'  block {i}: {b.value_name} [{b.offset}:{b.offset + b.size}] {b.family}({ps})')
        for t in self.terms:
            lines.append(f'  term: block {t.block}[idx] * {"x" if t.x is not None else "1"}')
        return '\n'.join(lines)
