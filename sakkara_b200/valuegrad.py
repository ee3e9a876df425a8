"""
``ValueGradFunction``: the callable PyMC's samplers invoke once per leapfrog step
(``pymc.Model.logp_dlogp_function()`` of the reference stack), evaluated by
libskk_b200.so.

    f = model.logp_dlogp_function(dtype='float64')
    logp, dlogp = f(theta)            # theta: raveled unconstrained parameters, creation order

A batch ``theta[B, P]`` evaluates B chains / leapfrog points in one call so that the row data is
read from HBM once per batch tile.  The device plan is created lazily in the calling process
(CUDA contexts do not survive ``fork``; PyMC runs one worker process per chain) and is dropped
when the object is pickled.
"""
import os
from typing import Optional

import numpy as np

from sakkara_b200.plan.kernel_plan import KernelPlan, make_kernel_plan
from sakkara_b200.plan.lower import UnsupportedModelError
from sakkara_b200.plan.model_plan import ModelPlan


class ValueGradFunction:
    def __init__(self, plan: ModelPlan, dtype: str = 'float64', max_batch: int = 1, device: Optional[int] = None,
                 rows=None, n_rows_total: Optional[int] = None, narrow_y: bool = True, include_priors: bool = True):
        self._include_priors = include_priors
        self.model_plan = plan
        self.dtype = np.dtype(dtype)
        self.max_batch = int(max_batch)
        self.device = device
        self._rows = rows
        self._n_rows_total = n_rows_total
        self._narrow_y = narrow_y
        self._kernel_plan: Optional[KernelPlan] = None
        self._skk = None
        self._pid = None

    # -- PyMC ValueGradFunction surface -------------------------------------------------------
    @property
    def size(self) -> int:
        return self.model_plan.n_params

    def set_extra_values(self, extra_vars) -> None:
        """All variables of the supported models are continuous; nothing to set."""

    def get_extra_values(self):
        return {}

    def __call__(self, theta, grad_out=None, extra_vars=None):
        theta = np.asarray(theta, dtype=self.dtype)
        logp, grad = self.skk.eval(theta)
        if grad_out is not None:
            grad_out[...] = grad
            return logp, grad_out
        return logp, grad

    # -- plan management ------------------------------------------------------------------------
    @property
    def kernel_plan(self) -> KernelPlan:
        if self._kernel_plan is None:
            self._kernel_plan = make_kernel_plan(self.model_plan, dtype=self.dtype, rows=self._rows,
                                                 n_rows_total=self._n_rows_total, narrow_y=self._narrow_y,
                                                 include_priors=self._include_priors)
        return self._kernel_plan

    @property
    def skk(self):
        if self._skk is None or self._pid != os.getpid():
            from sakkara_b200.runtime import SkkPlan
            device = self.device
            if device is None:
                device = int(os.environ.get('LOCAL_RANK', '0')) if 'LOCAL_RANK' in os.environ else 0
            self._skk = SkkPlan(self.kernel_plan, max_batch=self.max_batch, device=device)
            self._pid = os.getpid()
        return self._skk

    def stats(self) -> dict:
        return self.skk.stats()

    def close(self) -> None:
        if self._skk is not None and self._pid == os.getpid():
            self._skk.close()
        self._skk = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_skk'] = None
        state['_pid'] = None
        return state


class MinibatchValueGrad:
    """
    Value and gradient of a model with a ``MinibatchLikelihood`` (reference
    ``composable/hierarchical/likelihood.py:65-89``): every call draws ``plan.minibatch`` rows uniformly with
    replacement (``pymc.data.minibatch_index`` semantics, ``sakkara/relation/group.py:59``), evaluates the
    likelihood of those rows with the fused kernels and scales it by ``n_rows / minibatch`` (PyMC's ``total_size``);
    the prior terms are evaluated once per call by a plan without rows.  The estimate is unbiased for the full
    logp and gradient, which is what stochastic ADVI consumes.

    The rows of a call form a new (tiny) kernel plan -- sorted, uploaded, evaluated, dropped -- so a call costs a
    plan creation (about a millisecond), not a pass over the data set; nothing of the full data is touched.

    :param seed: seed of the row draws (``numpy.random.default_rng``); ``last_rows`` holds the rows of the last call.
    """

    def __init__(self, plan: ModelPlan, dtype: str = 'float64', device: Optional[int] = None, seed=None):
        if plan.minibatch is None:
            raise ValueError('the model has no MinibatchLikelihood')
        self.model_plan = plan
        self.dtype = np.dtype(dtype)
        self.device = device
        self.batch_size = int(plan.minibatch)
        self.scale = plan.n_rows / self.batch_size
        self.rng = np.random.default_rng(seed)
        self.last_rows: Optional[np.ndarray] = None
        self._priors = ValueGradFunction(plan.take_rows([]), dtype=dtype, device=device)

    @property
    def size(self) -> int:
        return self.model_plan.n_params

    def set_extra_values(self, extra_vars) -> None:
        pass

    def draw_rows(self) -> np.ndarray:
        return np.sort(self.rng.integers(0, self.model_plan.n_rows, size=self.batch_size))

    def __call__(self, theta, grad_out=None, extra_vars=None, rows=None):
        theta = np.asarray(theta, dtype=self.dtype)
        if theta.ndim != 1:
            raise ValueError('a mini-batched model evaluates one theta per call')
        rows = self.draw_rows() if rows is None else np.asarray(rows, dtype=np.int64)
        self.last_rows = rows
        batch = ValueGradFunction(self.model_plan, dtype=self.dtype, device=self.device, rows=rows,
                                  n_rows_total=len(rows), include_priors=False)
        try:
            lik_logp, lik_grad = batch(theta)
        finally:
            batch.close()
        prior_logp, prior_grad = self._priors(theta)
        scale = self.dtype.type(self.model_plan.n_rows / len(rows))
        logp = self.dtype.type(prior_logp + scale * lik_logp)
        grad = (prior_grad + scale * lik_grad).astype(self.dtype)
        if grad_out is not None:
            grad_out[...] = grad
            return logp, grad_out
        return logp, grad

    def close(self) -> None:
        self._priors.close()


class SplitSigmaValueGrad:
    """
    Value and gradient of a model whose Normal likelihood has a sigma that differs between rows and depends on theta
    -- sigma per group as a parameter, e.g. a ``GroupComponent`` of positive variables and constants (reference
    ``tests/workflows/test_state_space.py:64-82``) or a positive block gathered by group.  The rows whose sigma is a
    parameter form ONE plan whose row kernel gathers ``log(sigma)`` per row like a term of the linear predictor
    (``SKK_LIK_NORMAL_LOGSIGMA``: lp = -r^2/2 - zeta, d/d eta = r / sigma, d/d zeta = r^2 - 1), so any number of
    groups costs one pass; rows with constant sigmas, if any, form a standardised standard plan; the priors come from
    a plan without rows; the parts are added on the host in a fixed order.  ``per_element=True`` evaluates one
    standard plan per sigma parameter instead (the cross-check of the kernel family in the tests).
    """

    def __init__(self, plan: ModelPlan, dtype: str = 'float64', max_batch: int = 1, device: Optional[int] = None,
                 per_element: bool = False):
        self.model_plan = plan
        self.dtype = np.dtype(dtype)
        self.parts = [ValueGradFunction(part, dtype=dtype, max_batch=max_batch, device=device, include_priors=False)
                      for part in plan.split_by_sigma(per_element=per_element)]
        self._priors = ValueGradFunction(plan.take_rows([]), dtype=dtype, max_batch=max_batch, device=device)

    @property
    def size(self) -> int:
        return self.model_plan.n_params

    def set_extra_values(self, extra_vars) -> None:
        pass

    def __call__(self, theta, grad_out=None, extra_vars=None):
        theta = np.asarray(theta, dtype=self.dtype)
        logp, grad = self._priors(theta)
        # (the parts are taken with ModelPlan.take_rows, which leaves the constant of the rows masked for NaN behind)
        logp = np.array(logp, dtype=np.float64) + float(self.model_plan.logp_const)
        grad = np.array(grad, dtype=np.float64)
        for part in self.parts:
            lp, g = part(theta)
            logp = logp + lp
            grad = grad + g
        logp = logp.astype(self.dtype) if logp.ndim else self.dtype.type(logp)
        grad = grad.astype(self.dtype)
        if grad_out is not None:
            grad_out[...] = grad
            return logp, grad_out
        return logp, grad

    def stats(self) -> dict:
        return self.parts[0].stats()

    def close(self) -> None:
        for f in self.parts + [self._priors]:
            f.close()


def secondary_chunk_groups(dtype, max_batch: int) -> int:
    """
    Largest secondary level one plan should carry: the row kernel keeps one table of the level's values and one
    accumulator table per warp in shared memory -- ``(warps + 1) * groups * batch tile * sizeof(T)`` next to the warps'
    staging rings -- and gives up pipeline stages and then warps to make a larger level fit, until it cannot
    (``skk_plan_create``: "does not fit in shared memory").  With tables of up to 16 KB a CTA of eight warps still
    fits an SM (9 x 16 KB + rings < 227 KB); beyond that the level is split over several plans instead.
    """
    tile = 4 if max_batch > 1 else 1
    return max(256, (16 * 1024) // (np.dtype(dtype).itemsize * tile))


class PartitionedValueGrad:
    """
    Value and gradient of a model that one plan of the fused kernels cannot hold, as a sum of plans over disjoint
    sets of rows (the likelihood is a sum over rows; every row is still read once per evaluation):

    * a crossed model whose smaller (secondary) factor is too large for the row kernel's shared-memory tables is
      split by ranges of secondary groups into parts with at most ``chunk_groups`` of them each;
    * a model with more than two crossed factors is split by the groups of its smallest crossed factor -- inside one
      group that factor's term is a single element of theta, a global term of the part -- until two are left;
    * a model (or part) with more terms on a level than the kernels take is evaluated with the terms that have no data
      column folded into the finest of them (:func:`fold_terms`) -- possibly as a single part.

    Every part is an ordinary plan over its rows (same theta, its own sort, segments and tables) evaluated by the
    same kernels; the priors come from a plan without rows; the parts are added on the host in a fixed order.
    """

    def __init__(self, plan: ModelPlan, dtype: str = 'float64', max_batch: int = 1, device: Optional[int] = None,
                 chunk_groups: Optional[int] = None, max_parts: int = 256):
        self.model_plan = plan
        self.dtype = np.dtype(dtype)
        self.chunk_groups = int(chunk_groups or secondary_chunk_groups(dtype, max_batch))
        self.part_rows = partition_rows(plan, self.chunk_groups, max_parts)
        self.part_plans, self.aliases, self.parts = [], [], []
        for rows in self.part_rows:
            whole = plan.take_rows(rows)
            part, folds = fold_terms(whole, only_constant=True)
            f = ValueGradFunction(part, dtype=dtype, max_batch=max_batch, device=device, include_priors=False)
            try:
                f.kernel_plan
            except UnsupportedModelError as err:
                if 'compiled kernels cover' not in str(err):
                    raise
                # more terms on a level than the kernels take: fold every term the finest one determines
                part, folds = fold_terms(whole, only_constant=False)
                f = ValueGradFunction(part, dtype=dtype, max_batch=max_batch, device=device, include_priors=False)
                f.kernel_plan
            self.part_plans.append(part)
            self.aliases.append(folds)
            self.parts.append(f)
        self._priors = ValueGradFunction(plan.take_rows([]), dtype=dtype, max_batch=max_batch, device=device)

    @property
    def size(self) -> int:
        return self.model_plan.n_params

    def set_extra_values(self, extra_vars) -> None:
        pass

    def __call__(self, theta, grad_out=None, extra_vars=None):
        theta = np.asarray(theta, dtype=self.dtype)
        logp, grad = self._priors(theta)
        logp = np.array(logp, dtype=np.float64) + float(self.model_plan.logp_const)     # (rows masked for NaN)
        grad = np.array(grad, dtype=np.float64)
        for part, alias in zip(self.parts, self.aliases):
            lp, g = part(alias_theta(theta, alias))
            logp = logp + lp
            grad = grad + alias_grad(g, alias)
        logp = logp.astype(self.dtype) if logp.ndim else self.dtype.type(logp)
        grad = grad.astype(self.dtype)
        if grad_out is not None:
            grad_out[...] = grad
            return logp, grad_out
        return logp, grad

    def stats(self) -> dict:
        return self.parts[0].stats()

    def close(self) -> None:
        for f in self.parts + [self._priors]:
            f.close()


def fold_terms(part: ModelPlan, only_constant: bool = True):
    """
    Terms without a data column add elements of theta to the predictor; when the gather of one is a function of the
    gather of another -- a global intercept or a peeled crossed factor (constant inside the part) next to anything, a
    term on a parent level next to a term on its child level -- the two are one term of the finer level evaluated at
    ``theta'[child element] = theta[child element] + theta[parent element]``, and the coarser term's gradient is the
    sum of the finer one's over its children.  The kernels take 2 global, 2 primary-level and 1 secondary-level terms;
    folding evaluates models beyond that with the same kernels.  It needs the part's priors evaluated elsewhere
    (nothing else may read the modified elements): parts are plans without priors.

    ``only_constant``: fold only terms that are constant inside the part, into the first of them (what a part of a
    peeled crossed factor needs); otherwise every term that the finest one determines is folded into that.

    Returns ``(part, None)`` when nothing folds, else ``(part without the folded terms, folds)`` with ``folds`` a list
    of ``(host elements[G], folded elements[G], ratio of the two terms' constant coefficients)``, absolute positions in theta (:func:`alias_theta`, :func:`alias_grad`).
    """
    import dataclasses
    import pandas as pd
    from sakkara_b200.relation.groupset import first_rows
    # (a term whose data column is one constant c counts as well: c * theta_a + c' * theta_p = c * (theta_a + c'/c * theta_p))
    coef = {}
    for i, t in enumerate(part.terms if part.n_rows else []):
        if t.x is None:
            coef[i] = 1.0
        elif t.x.min() == t.x.max() and t.x[0] != 0.0:
            coef[i] = float(t.x[0])
    free = sorted(coef)
    if len(free) < 2:
        return part, None
    absolute = {i: part.term_base(part.terms[i]) + np.asarray(part.terms[i].index, dtype=np.int64) for i in free}
    distinct = {i: 1 if a.min() == a.max() else len(np.unique(a)) for i, a in absolute.items()}
    if only_constant:
        free = [i for i in free if distinct[i] == 1]
        if len(free) < 2:
            return part, None
        host = free[0]
    else:
        host = max(free, key=lambda i: distinct[i])            # the finest (the first among equals)
    codes, elements = pd.factorize(absolute[host], sort=False)
    elements = np.asarray(elements, dtype=np.int64)
    first = first_rows(codes.astype(np.int64, copy=False), len(elements))
    folds, dropped = [], set()
    for i in free:
        if i == host:
            continue
        table = absolute[i][first]
        if distinct[i] == 1 or np.array_equal(table[codes], absolute[i]):
            folds.append((elements, table, coef[i] / coef[host]))
            dropped.add(i)
    if not folds:
        return part, None
    for j, t in enumerate(part.terms):
        # (the host's elements must be read by the host term only: the part is evaluated with other values in them)
        if j != host and j not in dropped and np.isin(part.term_base(t) + t.index, elements).any():
            return part, None
    return dataclasses.replace(part, terms=[t for j, t in enumerate(part.terms) if j not in dropped]), folds


def merge_constant_terms(part: ModelPlan):
    """:func:`fold_terms` restricted to the terms that are constant inside the part."""
    return fold_terms(part, only_constant=True)


def alias_theta(theta: np.ndarray, folds) -> np.ndarray:
    """The point a folded part is evaluated at: ``theta[host] += ratio * theta[folded]`` for every fold."""
    if folds is None:
        return theta
    out = np.array(theta, copy=True)
    for host, folded, ratio in folds:
        out[..., host] += ratio * theta[..., folded]
    return out


def alias_grad(grad: np.ndarray, folds) -> np.ndarray:
    """The gradient of the unfolded model: every folded element receives the sum of its hosts' gradients."""
    if folds is None:
        return grad
    out = np.array(grad, dtype=np.float64, copy=True)
    flat = out.reshape(-1, out.shape[-1])
    source = np.array(grad, dtype=np.float64).reshape(-1, out.shape[-1])
    for host, folded, ratio in folds:
        for row in range(flat.shape[0]):
            np.add.at(flat[row], folded, ratio * source[row, host])
    return out


ChunkedSecondaryValueGrad = PartitionedValueGrad      # (the first of the two cases, under its first name)


def split_rows_by_secondary(plan: ModelPlan, chunk_groups: int):
    """
    Rows of a plannable model split so that no part's secondary level exceeds ``chunk_groups``: part k holds the rows
    whose secondary group (in the slot order of the whole model's kernel plan) lies in ``[k * size, (k + 1) * size)``,
    ``size`` = the level's groups spread evenly over the fewest parts.  One part (all rows) when the level fits.
    """
    structure = make_kernel_plan(plan, dtype='float64', structure_only=True)
    if structure.n_secondary <= chunk_groups:
        return [np.arange(plan.n_rows)]
    n_parts = -(-structure.n_secondary // chunk_groups)
    size = -(-structure.n_secondary // n_parts)
    part_of_sorted_row = structure.sec_codes.astype(np.int64) // size
    original = np.arange(plan.n_rows) if structure.perm is None else structure.perm
    order = np.argsort(part_of_sorted_row, kind='stable')
    bounds = np.searchsorted(part_of_sorted_row[order], np.arange(n_parts + 1))
    return [np.sort(original[order[bounds[k]:bounds[k + 1]]]) for k in range(n_parts)]


def partition_rows(plan: ModelPlan, chunk_groups: int, max_parts: int = 256):
    """
    Row numbers of the parts of :class:`PartitionedValueGrad` (ascending inside a part; one part = all rows when a
    single plan does).  More than two crossed factors: split by the smallest one, part by part, until the planner
    accepts; then every part by ranges of its secondary level.
    """
    from sakkara_b200.plan.kernel_plan import CrossedFactorsError

    def recurse(rows):
        sub = plan if rows is None else plan.take_rows(rows)
        base = np.arange(plan.n_rows) if rows is None else rows
        try:
            return [base[p] for p in split_rows_by_secondary(sub, chunk_groups)]
        except CrossedFactorsError as err:
            if err.size > max_parts:
                raise
            order = np.argsort(err.codes, kind='stable')
            bounds = np.searchsorted(err.codes[order], np.arange(err.size + 1))
            out = []
            for k in range(err.size):
                out += recurse(base[np.sort(order[bounds[k]:bounds[k + 1]])])
            return out

    parts = recurse(None)
    if len(parts) > max_parts:
        raise UnsupportedModelError(f'the model would take {len(parts)} plans (at most {max_parts})')
    return parts
