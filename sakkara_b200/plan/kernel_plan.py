"""
Kernel plan: what the CUDA library needs, derived from a :class:`ModelPlan` (SURVEY.md §7 step 1/2,
"index builder").  All integer work here must be exact; it is checked bit-for-bit in
``tests/test_kernel_plan.py`` (permutation, segment offsets, slot tables).

Decisions taken here:

* **levels** -- every term of the linear predictor indexes its block through a per-row code
  column.  Constant columns are *global* terms.  The remaining distinct columns are ordered by
  functional dependence (column c determines column d when every code of c occurs with one code of
  d -- the reference's parent test, ``sakkara/relation/groupset.py:42-44``).  The finest column
  becomes the **primary** level: rows are sorted by it, its groups are contiguous segments and the
  kernel needs no code column for it.  Columns it determines (nested parents) ride on the same
  segments through a slot -> theta table.  A second, crossed factor becomes the **secondary** level
  with a per-row code column (uint16 when it fits); rows are sorted by it inside every primary
  segment so that equal codes are contiguous.
* **slot order** -- primary groups are numbered so that the groups of one parent are adjacent
  (lexicographic in coarse-to-fine ancestor codes), which keeps parent-level reductions local.
* **row order** -- stable sort by (primary slot, secondary slot); LSD radix passes over 16-bit
  digits (NumPy's stable sort of uint16 is a radix sort), skipped when the rows are already sorted.
* **storage** -- data columns in the plan's floating type; 0/1 and small-count observations as
  uint8; identical data columns are stored once.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
from scipy.special import gammaln

from sakkara_b200.plan.lower import UnsupportedModelError
from sakkara_b200.relation.groupset import first_rows
from sakkara_b200.plan.model_plan import (FAMILY_IDS, LIKELIHOOD_IDS, PRIOR_PARAMS, XCOL_LOG_SIGMA, ModelPlan, Term,
                                          prior_normaliser)

LEVEL_GLOBAL, LEVEL_PRIMARY, LEVEL_SECONDARY = 0, 1, 2


class CrossedFactorsError(UnsupportedModelError):
    """
    More crossed grouping factors than one plan sorts by.  Carries the smallest of them -- ``codes[N]``, the factor's
    group of every row of the plan, and its ``size`` -- because inside one group of a factor its term is a single
    element of theta: rows split by that factor are plans with one crossed factor less
    (``valuegrad.PartitionedValueGrad``).
    """

    def __init__(self, message: str, codes: np.ndarray, size: int):
        super().__init__(message)
        self.codes, self.size = codes, int(size)
MAX_TERMS_PER_LEVEL = 2
MAX_SECONDARY_TERMS = 1
MAX_XCOLS = 6


@dataclass
class KernelTerm:
    level: int
    xcol: int                       # -1: no data column
    theta_index: np.ndarray         # int32[n_slots]
    block: int = -1                 # for diagnostics


@dataclass
class KernelBlock:
    family: int
    offset: int
    size: int
    a_const: Optional[np.ndarray] = None    # float64
    a_index: Optional[np.ndarray] = None    # int32, absolute theta index per element
    b_const: Optional[np.ndarray] = None
    b_index: Optional[np.ndarray] = None


@dataclass
class KernelPlan:
    dtype: np.dtype
    family: int
    n_rows: int
    n_rows_total: int
    n_params: int
    perm: Optional[np.ndarray]              # kernel row -> row of the (sharded) input; None = identity
    xcols: List[np.ndarray]
    y: np.ndarray
    y_kind: int                             # 0 float, 1 uint8
    eta_offset: Optional[np.ndarray]
    n_primary: int
    seg_offsets: np.ndarray                 # int64[n_primary + 1]
    n_secondary: int
    sec_codes: Optional[np.ndarray]         # uint16 / int32 [n_rows]
    terms: List[KernelTerm]
    blocks: List[KernelBlock]
    sigma_const: float
    sigma_theta_index: int
    logp_const: float
    include_priors: bool = True
    prior_logp_const: float = 0.0           # the part of logp_const that is normalisers of prior families

    @property
    def stored_bytes_per_row(self) -> int:
        s = self.dtype.itemsize
        n = len(self.xcols) * s + (1 if self.y_kind == 1 else s)
        if self.eta_offset is not None:
            n += s
        if self.sec_codes is not None:
            n += self.sec_codes.dtype.itemsize
        return n

    def describe(self) -> str:
        lv = {0: 'global', 1: 'primary', 2: 'secondary'}
        lines = [f'KernelPlan: N={self.n_rows} P={self.n_params} dtype={self.dtype} family={self.family} '
                 f'primary={self.n_primary} secondary={self.n_secondary} bytes/row={self.stored_bytes_per_row}']
        for t in self.terms:
            lines.append(f'  term[{lv[t.level]}] block {t.block} xcol {t.xcol} slots {len(t.theta_index)}')
        return '\n'.join(lines)


# ---------------------------------------------------------------------------------------------
# integer helpers
# ---------------------------------------------------------------------------------------------
def dense_codes(index: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """First-appearance factorisation: ``(codes[N] int64, elements[G] int64, first_row[G])``."""
    codes, elements = pd.factorize(index, sort=False)
    codes = codes.astype(np.int64, copy=False)
    return codes, np.asarray(elements, dtype=np.int64), first_rows(codes, len(elements))


def determines(codes_c: np.ndarray, first_row_c: np.ndarray, codes_d: np.ndarray) -> Optional[np.ndarray]:
    """Lookup ``code of c -> code of d`` if c determines d, else None."""
    table = codes_d[first_row_c]
    return table if np.array_equal(table[codes_c], codes_d) else None


def stable_argsort_small(keys: np.ndarray, n_keys: int) -> np.ndarray:
    """Stable argsort of non-negative integer keys < n_keys with 16-bit radix passes."""
    if n_keys <= 1 << 16:
        return np.argsort(keys.astype(np.uint16, copy=False), kind='stable')
    low = (keys & 0xFFFF).astype(np.uint16)
    order = np.argsort(low, kind='stable')
    high = (keys >> 16)[order]
    if n_keys <= 1 << 32:
        order = order[np.argsort(high.astype(np.uint16), kind='stable')]
        return order
    return np.argsort(keys, kind='stable')


def sort_rows(primary: Optional[np.ndarray], n_primary: int, secondary: Optional[np.ndarray],
              n_secondary: int) -> Optional[np.ndarray]:
    """Permutation that sorts rows by (primary, secondary), stable; None when already sorted."""
    n = len(primary) if primary is not None else (len(secondary) if secondary is not None else 0)
    if n == 0 or (primary is None and secondary is None):
        return None

    def is_sorted() -> bool:
        if primary is not None and secondary is not None:
            if (primary[1:] < primary[:-1]).any():
                return False
            return not ((primary[1:] == primary[:-1]) & (secondary[1:] < secondary[:-1])).any()
        keys = primary if primary is not None else secondary
        return not (keys[1:] < keys[:-1]).any()

    if is_sorted():
        return None
    perm = None
    if secondary is not None:
        perm = stable_argsort_small(secondary, n_secondary)
    if primary is not None:
        keys = primary if perm is None else primary[perm]
        order = stable_argsort_small(keys, n_primary)
        perm = order if perm is None else perm[order]
    return perm.astype(np.int64, copy=False)


# ---------------------------------------------------------------------------------------------
@dataclass
class _Column:
    codes: np.ndarray
    elements: np.ndarray       # block element (flat index) of every code
    first_row: np.ndarray
    terms: List[int] = field(default_factory=list)

    @property
    def size(self) -> int:
        return len(self.elements)


def make_kernel_plan(plan: ModelPlan, dtype='float64', rows: Optional[np.ndarray] = None,
                     n_rows_total: Optional[int] = None, narrow_y: bool = True,
                     include_priors: bool = True, structure_only: bool = False) -> KernelPlan:
    """
    :param rows: optional subset of the rows (this rank's shard); all ranks pass the same ``plan``
        and disjoint ``rows`` whose union is all rows.
    :param n_rows_total: number of rows over all ranks (defaults to this plan's rows).
    :param structure_only: compute levels, permutation and segments only (no data columns); used to
        cut sorted rows into per-rank shards.
    """
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype('float32'), np.dtype('float64')):
        raise ValueError('dtype must be float32 or float64')
    lik = plan.likelihood
    log_sigma = lik.sigma_rows_element is not None and plan.n_rows > 0
    if log_sigma and (lik.sigma_rows_element < 0).any():
        raise UnsupportedModelError('rows with a constant sigma next to rows whose sigma is a parameter: evaluate the '
                                    'parts of ModelPlan.split_by_sigma() (model.logp_dlogp_function() does that)')

    def take(a):
        if a is None or rows is None:
            return a
        return a[rows]

    if rows is None:
        n = plan.n_rows
    elif isinstance(rows, slice):
        n = len(range(*rows.indices(plan.n_rows)))
    else:
        n = len(rows)
    total = n if n_rows_total is None else int(n_rows_total)

    # ---- classify the terms ------------------------------------------------------------------------
    global_terms: Dict[int, int] = {}          # term number -> theta index
    columns: List[_Column] = []
    column_of_term: Dict[int, int] = {}
    elem_of_term: Dict[int, np.ndarray] = {}    # block element per code, per term
    seen_arrays: Dict[int, int] = {}
    plan_terms = list(plan.terms) if n else []     # a plan without rows (the priors alone) has no linear predictor
    sigma_term = -1
    if log_sigma and n:
        # sigma per row = exp(theta[element]): log(sigma) is gathered like a term of the predictor (absolute elements)
        sigma_term = len(plan_terms)
        plan_terms.append(Term(block=-1, index=lik.sigma_rows_element, x=None))
    positive = np.zeros(plan.n_params, dtype=bool)
    for b in plan.blocks:
        positive[b.offset:b.offset + b.size] = b.transform != 'identity'
    for number, term in enumerate(plan_terms):
        base = plan.term_base(term)
        index = take(term.index)
        if term.block >= 0:
            block = plan.blocks[term.block]
            if block.transform != 'identity':
                raise UnsupportedModelError(f'{block.name}: positive variables cannot enter the linear predictor')
        elif number == sigma_term:
            if n and not positive[np.unique(index)].all():
                raise UnsupportedModelError('a likelihood sigma must be a positive (log-transformed) variable')
        elif n and positive[np.unique(index)].any():
            raise UnsupportedModelError('positive variables cannot enter the linear predictor')
        if n == 0 or (term.block >= 0 and block.size == 1) or index.min() == index.max():
            global_terms[number] = base + (int(index[0]) if n else 0)
            continue
        key = id(term.index)
        if key in seen_arrays:
            col = seen_arrays[key]
            column_of_term[number] = col
            elem_of_term[number] = columns[col].elements
            columns[col].terms.append(number)
            continue
        codes, elements, first_row = dense_codes(index)
        match = None
        for c, other in enumerate(columns):
            if other.size == len(elements) and np.array_equal(other.codes, codes):
                match = c
                break
        if match is None:
            columns.append(_Column(codes, elements, first_row))
            match = len(columns) - 1
        seen_arrays[key] = match
        column_of_term[number] = match
        elem_of_term[number] = elements       # same partition, but this term's own block elements
        columns[match].terms.append(number)

    # ---- order the columns by functional dependence --------------------------------------------------
    k = len(columns)
    table: Dict[Tuple[int, int], np.ndarray] = {}
    for c in range(k):
        for d in range(k):
            if c != d and columns[c].size >= columns[d].size:
                t = determines(columns[c].codes, columns[c].first_row, columns[d].codes)
                if t is not None:
                    table[(c, d)] = t
    finest = [c for c in range(k)
              if not any((d, c) in table and (c, d) not in table for d in range(k) if d != c)]
    # twins cannot occur here (identical partitions were merged above)
    if len(finest) > 2:
        smallest = min(finest, key=lambda c: columns[c].size)
        raise CrossedFactorsError(f'{len(finest)} crossed grouping factors; the fused kernel handles at most two',
                                  codes=columns[smallest].codes, size=columns[smallest].size)
    primary = secondary = None
    if len(finest) == 1:
        primary = finest[0]
    elif len(finest) == 2:
        a, b = finest
        # the smaller factor goes to the shared-memory table, the larger one is sorted
        primary, secondary = (a, b) if columns[a].size >= columns[b].size else (b, a)

    def assign_owners() -> Dict[int, int]:
        out: Dict[int, int] = {}
        for c in range(k):
            if c == primary or c == secondary:
                out[c] = c
            elif primary is not None and (primary, c) in table:
                out[c] = primary
            elif secondary is not None and (secondary, c) in table:
                out[c] = secondary
            else:
                raise UnsupportedModelError('a grouping column is determined by neither sorted level')
        return out

    def level_terms(fine) -> int:
        return sum(1 for c in column_of_term.values() if owner[c] == fine)

    owner = assign_owners()
    if secondary is not None and level_terms(secondary) > MAX_SECONDARY_TERMS and \
            level_terms(primary) <= MAX_SECONDARY_TERMS:
        # the smaller factor carries more terms than the secondary level takes and the larger one few enough: they
        # change places (the secondary level's tables are then the larger ones; valuegrad.PartitionedValueGrad splits a
        # level that does not fit)
        primary, secondary = secondary, primary
        owner = assign_owners()
    if primary is not None and secondary is None and level_terms(primary) > MAX_TERMS_PER_LEVEL:
        # One grouping factor (with its parents) carries more terms than the primary level takes -- an intercept and two
        # random slopes by the same group -- and there is no crossed factor: the secondary level is free.  The last of the
        # terms moves there, on a copy of its own column (the rows are sorted by the primary level, so a level it determines
        # is constant inside every segment: trivially sorted inside it)
        number = max(t for t, c in column_of_term.items() if owner[c] == primary)
        c = column_of_term[number]
        src = columns[c]
        src.terms.remove(number)
        columns.append(_Column(src.codes, src.elements, src.first_row, [number]))
        secondary = len(columns) - 1
        column_of_term[number] = secondary
        for (a, b), t in list(table.items()):
            if b == c:
                table[(a, secondary)] = t
            if a == c:
                table[(secondary, b)] = t
        table[(c, secondary)] = table[(secondary, c)] = np.arange(src.size)
        k = len(columns)
        owner = assign_owners()

    def slot_order(fine: int) -> np.ndarray:
        """codes of `fine` ordered so that the groups of one ancestor are adjacent"""
        col = columns[fine]
        ancestors = sorted((c for c in range(k) if c != fine and owner.get(c) == fine),
                           key=lambda c: columns[c].size)
        keys = [np.arange(col.size)] + [table[(fine, a)] for a in reversed(ancestors)]
        return np.lexsort(tuple(keys)) if ancestors else np.arange(col.size)

    # ---- more global terms than the kernels take -------------------------------------------------------------------
    # A global term is a term on the ancestor of every level: like any term on a parent of the sorted level it can ride on
    # the primary segments with a slot -> theta table that names its one element for every slot (the finish kernel adds
    # the slots of such an element in fixed order).  The surplus moves there while the primary level has room; a model
    # without any grouping level gets a primary level of one segment.
    promoted = set()
    primary_terms = sum(1 for c in column_of_term.values() if primary is not None and owner[c] == primary)
    surplus, room = len(global_terms) - MAX_TERMS_PER_LEVEL, MAX_TERMS_PER_LEVEL - primary_terms
    if n and surplus > 0 and room > 0:
        promoted = set(sorted(global_terms)[len(global_terms) - min(surplus, room):])

    n_primary = n_secondary = 0
    p_slot_of_row = s_slot_of_row = None
    p_code_of_slot = s_code_of_slot = None
    if primary is None and promoted:
        n_primary = 1
        p_slot_of_row = np.zeros(n, dtype=np.int64)
    if primary is not None:
        p_code_of_slot = slot_order(primary)
        n_primary = len(p_code_of_slot)
        inv = np.empty(n_primary, dtype=np.int64)
        inv[p_code_of_slot] = np.arange(n_primary)
        p_slot_of_row = inv[columns[primary].codes]
    if secondary is not None:
        s_code_of_slot = slot_order(secondary)
        own = [t for t in columns[secondary].terms]
        if len(own) == 1 and not any(c != secondary and owner.get(c) == secondary for c in range(k)):
            # the level's only term sits on the level itself: number the slots in the order of the term's theta
            # elements, so that slot -> theta is `base + slot` and the kernel fills its value table with one
            # coalesced read of theta instead of a gather through the slot table
            s_code_of_slot = np.argsort(elem_of_term[own[0]], kind='stable')
        n_secondary = len(s_code_of_slot)
        inv = np.empty(n_secondary, dtype=np.int64)
        inv[s_code_of_slot] = np.arange(n_secondary)
        s_slot_of_row = inv[columns[secondary].codes]

    perm = sort_rows(p_slot_of_row, n_primary, s_slot_of_row, n_secondary)

    def arrange(a):
        return a if perm is None else a[perm]

    seg_offsets = np.zeros(n_primary + 1, dtype=np.int64)
    if p_slot_of_row is not None:
        np.cumsum(np.bincount(p_slot_of_row, minlength=n_primary), out=seg_offsets[1:])
    sec_codes = None
    if secondary is not None:
        sec_codes = arrange(s_slot_of_row).astype(np.uint16 if n_secondary <= 1 << 16 else np.int32)

    if structure_only:
        return KernelPlan(dtype=dtype, family=LIKELIHOOD_IDS[lik.family], n_rows=n, n_rows_total=total,
                          n_params=plan.n_params, perm=perm, xcols=[], y=np.empty(0), y_kind=0, eta_offset=None,
                          n_primary=n_primary, seg_offsets=seg_offsets, n_secondary=n_secondary, sec_codes=sec_codes,
                          terms=[], blocks=[], sigma_const=1.0, sigma_theta_index=-1, logp_const=0.0)

    # ---- data columns ------------------------------------------------------------------------------------
    xcols: List[np.ndarray] = []
    x_sources: List[np.ndarray] = []

    def xcol_of(x: Optional[np.ndarray]) -> int:
        if x is None:
            return -1
        for c, src in enumerate(x_sources):
            if src is x or (src.shape == x.shape and np.array_equal(src, x)):
                return c
        if len(xcols) >= MAX_XCOLS:
            raise UnsupportedModelError(f'more than {MAX_XCOLS} distinct data columns in the linear predictor')
        x_sources.append(x)
        xcols.append(np.ascontiguousarray(arrange(take(x)), dtype=dtype))
        return len(xcols) - 1

    # ---- terms --------------------------------------------------------------------------------------------------
    terms: List[KernelTerm] = []
    counts = {LEVEL_GLOBAL: 0, LEVEL_PRIMARY: 0, LEVEL_SECONDARY: 0}
    for number, term in enumerate(plan_terms):
        xc = XCOL_LOG_SIGMA if number == sigma_term else xcol_of(term.x)
        if number in promoted:
            level = LEVEL_PRIMARY
            theta_index = np.full(n_primary, global_terms[number], dtype=np.int32)
        elif number in global_terms:
            level = LEVEL_GLOBAL
            theta_index = np.array([global_terms[number]], dtype=np.int32)
        else:
            c = column_of_term[number]
            fine = owner[c]
            level = LEVEL_PRIMARY if fine == primary else LEVEL_SECONDARY
            code_of_slot = p_code_of_slot if fine == primary else s_code_of_slot
            own_codes = code_of_slot if c == fine else table[(fine, c)][code_of_slot]
            theta_index = (plan.term_base(term) + elem_of_term[number][own_codes]).astype(np.int32)
        counts[level] += 1
        terms.append(KernelTerm(level=level, xcol=xc, theta_index=theta_index, block=term.block))
    if counts[LEVEL_GLOBAL] > MAX_TERMS_PER_LEVEL or counts[LEVEL_PRIMARY] > MAX_TERMS_PER_LEVEL \
            or counts[LEVEL_SECONDARY] > MAX_SECONDARY_TERMS:
        raise UnsupportedModelError(
            f'{counts[LEVEL_GLOBAL]} global / {counts[LEVEL_PRIMARY]} primary / {counts[LEVEL_SECONDARY]} secondary '
            f'terms; the compiled kernels cover {MAX_TERMS_PER_LEVEL}/{MAX_TERMS_PER_LEVEL}/{MAX_SECONDARY_TERMS}')

    # ---- observed column, offset, constants ---------------------------------------------------------------
    y_rows = take(lik.y)
    y_sorted = arrange(y_rows)
    y_kind = 0
    if narrow_y and lik.family in ('Bernoulli', 'Poisson', 'NegativeBinomial') and (n == 0 or y_rows.max() <= 255):
        y_store = y_sorted.astype(np.uint8)
        y_kind = 1
    else:
        y_store = np.ascontiguousarray(y_sorted, dtype=dtype)
    eta_offset = None
    if lik.offset is not None:
        eta_offset = np.ascontiguousarray(arrange(take(lik.offset)), dtype=dtype)
    logp_const = 0.0
    if lik.family == 'Poisson':
        logp_const = -float(np.sum(gammaln(y_rows + 1.0), dtype=np.float64))
    elif lik.family == 'NegativeBinomial':
        # per row: lgamma(y + alpha) - lgamma(alpha) - lgamma(y + 1) + alpha log(alpha); the kernel adds
        # y eta - (y + alpha) log(alpha + exp(eta))
        a = float(lik.sigma_const)
        logp_const = float(np.sum(gammaln(y_rows + a) - gammaln(y_rows + 1.0), dtype=np.float64)) + \
            n * (a * np.log(a) - float(gammaln(a)))
    elif lik.family == 'Gamma':
        # per row: alpha log(alpha) - lgamma(alpha) + (alpha - 1) log(y); the kernel adds -alpha (eta + y exp(-eta))
        a = float(lik.sigma_const)
        logp_const = (a - 1.0) * float(np.sum(np.log(y_rows), dtype=np.float64)) + n * (a * np.log(a) - float(gammaln(a)))
    if plan.logp_const:
        # rows masked out for NaN observations (ModelPlan.logp_const): every rank adds its share, so that
        # the all-reduced constant of a sharded run is the whole
        logp_const += float(plan.logp_const) * (n / total if total else 1.0)

    sigma_const, sigma_theta_index = 1.0, -1
    family = LIKELIHOOD_IDS[lik.family]
    if lik.family == 'StudentT':
        # the kernel adds -(nu + 1)/2 log(1 + r^2/nu) - log(sigma) per row (sigma = 1 without a zeta term)
        nu = float(lik.nu)
        sigma_const = nu                                             # (nu travels in the sigma field of the C ABI)
        logp_const += n * float(gammaln(0.5 * (nu + 1.0)) - gammaln(0.5 * nu) - 0.5 * np.log(nu * np.pi))
    elif log_sigma:
        family = LIKELIHOOD_IDS['NormalLogSigma']
        logp_const -= n * 0.91893853320467274178032973640562        # the kernel adds -r^2 / 2 - log(sigma) per row
    if lik.family in ('NegativeBinomial', 'Gamma'):
        sigma_const = float(lik.sigma_const)             # (its alpha travels in the same field of the C ABI)
    if lik.family == 'Normal' and not log_sigma:
        if lik.sigma_block is None:
            sigma_const = float(lik.sigma_const)
        else:
            sigma_theta_index = plan.blocks[lik.sigma_block].offset + lik.sigma_index

    # ---- priors ---------------------------------------------------------------------------------------------------
    blocks: List[KernelBlock] = []
    prior_logp_const = 0.0
    for b in plan.blocks:
        kb = KernelBlock(family=FAMILY_IDS[b.family], offset=b.offset, size=b.size)
        names = PRIOR_PARAMS[b.family]
        for slot, pname in zip('ab', names):
            if pname is None:
                continue
            prm = b.params[pname]
            if prm.is_const:
                setattr(kb, slot + '_const', np.ascontiguousarray(np.atleast_1d(prm.const), dtype=np.float64))
            elif prm.is_mixed:
                # per element a constant or an element of theta (index -1: the constant)
                setattr(kb, slot + '_const', np.ascontiguousarray(np.broadcast_to(prm.const, (b.size,)), dtype=np.float64))
                setattr(kb, slot + '_index', np.ascontiguousarray(prm.index, dtype=np.int32))
            else:
                src = plan.blocks[prm.block]
                setattr(kb, slot + '_index', (src.offset + prm.index).astype(np.int32))
        blocks.append(kb)
        if include_priors and FAMILY_IDS[b.family] >= FAMILY_IDS['HalfCauchy']:
            # the device evaluates these families without their normaliser; like the constant of the masked
            # rows every rank adds its share
            norm = np.broadcast_to(prior_normaliser(b.family, kb.a_const, kb.b_const), (b.size,))
            prior_logp_const += float(np.sum(norm, dtype=np.float64)) * (n / total if total else 1.0)

    return KernelPlan(dtype=dtype, family=family, n_rows=n, n_rows_total=total,
                      n_params=plan.n_params, perm=perm, xcols=xcols, y=np.ascontiguousarray(y_store), y_kind=y_kind,
                      eta_offset=eta_offset, n_primary=n_primary, seg_offsets=seg_offsets, n_secondary=n_secondary,
                      sec_codes=None if sec_codes is None else np.ascontiguousarray(sec_codes), terms=terms,
                      blocks=blocks, sigma_const=sigma_const, sigma_theta_index=sigma_theta_index,
                      logp_const=logp_const + prior_logp_const, include_priors=include_priors,
                      prior_logp_const=prior_logp_const)
