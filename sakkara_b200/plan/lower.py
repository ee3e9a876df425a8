"""
Lowering of a recorded model (:mod:`sakkara_b200.model.sym`) to a :class:`ModelPlan`.

The symbolic graph the components build contains exactly what the reference hands to PyTensor:
gathers ``element[index arrays]`` (reference ``sakkara/relation/representation.py:253``) and
elementwise arithmetic (``sakkara/model/function/base.py:88-92``).  The lowering brings every
parameter expression into the normal form

    sum_k  coef_k (*) theta_{block_k}[index_k]  +  const

by pushing gathers down to the leaves (index composition) and distributing products over sums.
Both spellings of a crossed model (``a + b + beta*x`` which builds a dense [A, B] intermediate in
the reference, and ``beta*x + a + b``) therefore end up as the same per-row
gathers.  Anything that is not affine in the parameters -- products of two blocks, ``pow``,
arbitrary ``f_`` callables -- raises ``NotImplementedError``: there is no CPU fallback on this
path, the caller is told to use the stock PyMC graph.
"""
import threading
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Union

import numpy as np

from sakkara_b200.model import sym
from sakkara_b200.model.dist import FAMILIES
from sakkara_b200.plan.model_plan import Block, LikelihoodSpec, ModelPlan, PriorParam, Term


class UnsupportedModelError(NotImplementedError):
    """The model has a structure the fused GLM kernel cannot express."""


@dataclass
class Monomial:
    coef: Union[float, np.ndarray]             # scalar or array broadcastable to the form's shape
    var: Optional[sym.RandomVar] = None        # free variable, None for the constant part
    index: Optional[np.ndarray] = None         # int64, flat element of `var` per cell, form-shaped
    tag: Optional[int] = None                  # monomials that stem from the pieces of one Concat node: their
                                               # supports (coef != 0) are disjoint, so they can be read as ONE
                                               # gather over several blocks (_merge_pieces)


class AffineForm:
    def __init__(self, shape, monomials: List[Monomial]):
        self.shape = tuple(shape)
        self.monomials = monomials

    @property
    def is_const(self) -> bool:
        return all(m.var is None for m in self.monomials)

    def const_value(self) -> np.ndarray:
        total = np.zeros(self.shape, dtype=np.float64)
        for m in self.monomials:
            assert m.var is None
            total = total + m.coef
        return total


class _Seen(threading.local):
    """MinibatchIndex nodes met by _lower since lower_model() started (per thread)."""

    def __init__(self):
        self.nodes: List[Any] = []


_seen = _Seen()


def _scale(coef, factor):
    return coef * factor


def _lower(node: Any) -> AffineForm:
    if isinstance(node, sym.DeterministicVar):
        return _lower(node.var)
    if isinstance(node, sym.RandomVar):
        if node.is_observed:
            raise UnsupportedModelError('Observed variables cannot be parameters of other variables')
        index = np.arange(node.size, dtype=np.int64).reshape(node.shape)
        return AffineForm(node.shape, [Monomial(1.0, node, index)])
    if isinstance(node, sym.DataVar):
        return AffineForm(node.shape, [Monomial(np.asarray(node.values, dtype=np.float64))])
    if isinstance(node, sym.Gather):
        base = _lower(node.base)
        out = []
        for m in base.monomials:
            coef = m.coef
            if isinstance(coef, np.ndarray) and coef.ndim > 0:
                coef = np.broadcast_to(coef, base.shape)[node.indices]
            index = None if m.var is None else np.broadcast_to(m.index, base.shape)[node.indices]
            out.append(Monomial(coef, m.var, index, m.tag))
        return AffineForm(node.shape, out)
    if isinstance(node, sym.MinibatchGather):
        # lowered over ALL rows: which rows an evaluation uses is decided when it runs (ModelPlan.minibatch)
        _seen.nodes.append(node.minibatch)
        return _lower(node.base)
    if isinstance(node, sym.Concat):
        # piece by piece: the piece's monomials, zero outside the piece's range
        total = node.shape[0]
        out, start = [], 0
        for piece, length in zip(node.pieces, node.lengths):
            form = _lower(piece)
            for m in form.monomials:
                coef = np.zeros(total, dtype=np.float64)
                coef[start:start + length] = np.broadcast_to(m.coef, (length,))
                index = None
                if m.var is not None:
                    index = np.zeros(total, dtype=np.int64)
                    index[start:start + length] = np.broadcast_to(m.index, (length,))
                out.append(Monomial(coef, m.var, index, id(node) if m.var is not None else None))
            start += length
        return AffineForm(node.shape, out)
    if isinstance(node, sym.UnaryOp):
        if node.op == 'neg':
            inner = _lower(node.arg)
            return AffineForm(inner.shape, [Monomial(_scale(m.coef, -1.0), m.var, m.index, m.tag) for m in inner.monomials])
        raise UnsupportedModelError(f'{node.op}() may only wrap the whole linear predictor of the likelihood')
    if isinstance(node, sym.BinOp):
        left, right = _lower(node.left), _lower(node.right)
        shape = np.broadcast_shapes(left.shape, right.shape)
        if node.op in ('add', 'sub'):
            sign = 1.0 if node.op == 'add' else -1.0
            monos = [_broadcast(m, left.shape, shape) for m in left.monomials]
            monos += [_broadcast(Monomial(_scale(m.coef, sign), m.var, m.index, m.tag), right.shape, shape)
                      for m in right.monomials]
            return AffineForm(shape, monos)
        if node.op == 'mul':
            if not (left.is_const or right.is_const):
                raise UnsupportedModelError('Products of two random variables are not affine in theta')
            const, other, other_shape = (left, right, right.shape) if left.is_const else (right, left, left.shape)
            factor = const.const_value()
            factor = float(factor) if factor.ndim == 0 else factor
            return AffineForm(shape, [_broadcast(Monomial(_scale(m.coef, factor), m.var, m.index, m.tag), other_shape, shape)
                                      for m in other.monomials])
        if node.op == 'truediv':
            if not right.is_const:
                raise UnsupportedModelError('Division by a random variable is not affine in theta')
            factor = 1.0 / right.const_value()
            factor = float(factor) if factor.ndim == 0 else factor
            return AffineForm(shape, [_broadcast(Monomial(_scale(m.coef, factor), m.var, m.index, m.tag), left.shape, shape)
                                      for m in left.monomials])
        raise UnsupportedModelError(f'Operator {node.op} cannot be lowered to the fused GLM kernel')
    if isinstance(node, sym.Var):
        raise UnsupportedModelError(f'Cannot lower node of type {type(node).__name__}')
    value = np.asarray(node, dtype=np.float64)
    return AffineForm(value.shape, [Monomial(float(value) if value.ndim == 0 else value)])


def _broadcast(m: Monomial, from_shape, to_shape) -> Monomial:
    """Bring the index array of a monomial to the (broadcast) shape of the enclosing form."""
    if m.var is None or tuple(from_shape) == tuple(to_shape):
        return m
    return Monomial(m.coef, m.var, np.broadcast_to(m.index, to_shape), m.tag)


def _merge_constants(form: AffineForm) -> Optional[np.ndarray]:
    consts = [m for m in form.monomials if m.var is None]
    if not consts:
        return None
    total = 0.0
    for m in consts:
        total = total + m.coef
    return total


@dataclass
class _Gather:
    """A merged monomial: ``coef * theta[absolute index]`` per cell, reading one block or several."""
    coef: np.ndarray                           # float64, form-shaped
    absolute: np.ndarray                       # int64, form-shaped: element of theta
    blocks: List[int]                          # the blocks it reads, in order of appearance


def _merge_pieces(form: AffineForm, block_of, blocks) -> List[_Gather]:
    """
    The non-constant monomials of ``form`` as gathers from theta.  Monomials that stem from the pieces of one
    ``Concat`` (a ``GroupComponent``) and whose supports are disjoint are read as one gather with absolute
    indices: ``c1 * a[i1] + c2 * b[i2] == (c1 + c2) * theta[where(c1 != 0, off_a + i1, off_b + i2)]``.
    """
    out: List[_Gather] = []
    groups: Dict[int, List[int]] = {}
    supports: Dict[int, np.ndarray] = {}
    for m in form.monomials:
        if m.var is None:
            continue
        number = block_of[id(m.var)]
        coef = np.array(np.broadcast_to(np.asarray(m.coef, dtype=np.float64), form.shape))
        absolute = blocks[number].offset + np.broadcast_to(m.index, form.shape).astype(np.int64)
        if m.tag is not None:
            support = coef != 0.0
            target = None
            for k in groups.get(m.tag, []):
                if not (supports[k] & support).any():
                    target = k
                    break
            if target is not None:
                g = out[target]
                g.absolute = np.where(support, absolute, g.absolute)
                g.coef = g.coef + coef
                supports[target] |= support
                if number not in g.blocks:
                    g.blocks.append(number)
                continue
            groups.setdefault(m.tag, []).append(len(out))
            supports[len(out)] = support
        out.append(_Gather(coef, np.array(absolute), [number]))
    return out


def _prior_param(value: Any, block_of, size: int, what: str, blocks=None) -> PriorParam:
    form = _lower(value)
    if form.is_const:
        const = np.asarray(form.const_value(), dtype=np.float64)
        if const.size not in (1, size):
            raise UnsupportedModelError(f'{what}: constant of size {const.size} for a block of size {size}')
        return PriorParam(const=const.reshape(-1) if const.size > 1 else const.reshape(()))
    monos = form.monomials
    if any(m.tag is not None for m in monos):
        # member-wise parameter (GroupComponent): per element either a constant or one element of theta
        gathers = _merge_pieces(form, block_of, blocks)
        const = _merge_constants(form)
        const = np.zeros(form.shape) if const is None else np.broadcast_to(np.asarray(const, dtype=np.float64), form.shape)
        if len(gathers) != 1 or not np.isin(gathers[0].coef, (0.0, 1.0)).all() or \
                np.any((gathers[0].coef == 1.0) & (const != 0.0)):
            raise UnsupportedModelError(f'{what}: every element must be a constant or a single variable')
        g = gathers[0]
        absolute = np.where(g.coef == 1.0, g.absolute, -1).reshape(-1)
        const = np.array(const, dtype=np.float64).reshape(-1)
        if absolute.size == 1 and size > 1:
            absolute, const = np.repeat(absolute, size), np.repeat(const, size)
        if absolute.size != size:
            raise UnsupportedModelError(f'{what}: parameter of size {absolute.size} for a block of size {size}')
        return PriorParam(const=const, block=-1, index=np.ascontiguousarray(absolute, dtype=np.int64))
    if len(monos) != 1 or not np.all(np.asarray(monos[0].coef) == 1.0):
        raise UnsupportedModelError(f'{what}: prior parameters must be constants or a single gathered variable')
    m = monos[0]
    index = np.ascontiguousarray(np.broadcast_to(m.index, form.shape), dtype=np.int64).reshape(-1)
    if index.size == 1 and size > 1:
        index = np.repeat(index, size)
    if index.size != size:
        raise UnsupportedModelError(f'{what}: parameter of size {index.size} for a block of size {size}')
    return PriorParam(block=block_of[id(m.var)], index=index)


def _masked_likelihood(obs: sym.RandomVar, y: np.ndarray):
    """
    NaN-masked likelihood (reference ``hierarchical/likelihood.py:32-57``): every parameter of the
    observed variable is a ``sym.Masked`` node over the same rows.  Returns the parameters without
    the wrappers, the row mask and the constant the masked rows add to logp (their parameters and
    their observation are constants), or ``(params, None, 0.0)`` for an ordinary likelihood.
    """
    wrapped = {k: v for k, v in obs.params.items() if isinstance(v, sym.Masked)}
    if not wrapped:
        return dict(obs.params), None, 0.0
    if len(wrapped) != len(obs.params):
        raise UnsupportedModelError('Either all or none of the likelihood parameters carry a NaN mask')
    mask = None
    for node in wrapped.values():
        rows = np.broadcast_to(node.mask, obs.shape).reshape(-1)     # (never mini-batched: components.py)
        if mask is not None and not np.array_equal(mask, rows):
            raise UnsupportedModelError('Likelihood parameters are masked on different rows')
        mask = rows
    fill = {}
    for key, node in wrapped.items():
        form = _lower(node.fill)
        if not form.is_const:
            raise UnsupportedModelError(f'nan_param_mask[{key}] must be a constant')
        fill[key] = np.broadcast_to(np.asarray(form.const_value(), dtype=np.float64), obs.shape).reshape(-1)[mask]
    y0 = y[mask]
    if obs.family == 'Normal':
        z = (y0 - fill['mu']) / fill['sigma']
        terms = -0.5 * z * z - np.log(fill['sigma']) - 0.91893853320467274178032973640562
    elif obs.family == 'Bernoulli':
        if 'logit_p' in fill:
            terms = y0 * fill['logit_p'] - np.logaddexp(0.0, fill['logit_p'])
        else:
            with np.errstate(divide='ignore', invalid='ignore'):   # 0 * log(0) counts as 0
                terms = np.where(y0 > 0, y0 * np.log(fill['p']), 0.0) + \
                    np.where(y0 < 1, (1.0 - y0) * np.log1p(-fill['p']), 0.0)
    elif obs.family == 'Poisson':
        from scipy.special import gammaln
        with np.errstate(divide='ignore', invalid='ignore'):
            terms = np.where(y0 > 0, y0 * np.log(fill['mu']), 0.0) - fill['mu'] - gammaln(y0 + 1.0)
    elif obs.family == 'NegativeBinomial':
        from scipy.stats import nbinom
        terms = nbinom.logpmf(y0, fill['alpha'], fill['alpha'] / (fill['alpha'] + fill['mu']))
    elif obs.family == 'StudentT':
        from scipy.stats import t as student_t
        terms = student_t.logpdf(y0, fill['nu'], loc=fill['mu'], scale=fill['sigma'])
    elif obs.family == 'Gamma':
        from scipy.stats import gamma as gamma_dist
        terms = gamma_dist.logpdf(y0, fill['alpha'], scale=1.0 / fill['beta'])
    else:
        raise UnsupportedModelError(f'Likelihood family {obs.family} is not supported')
    return {k: v.value for k, v in wrapped.items()}, mask, float(np.sum(terms, dtype=np.float64))


def lower_model(model: sym.Model) -> ModelPlan:
    """
    Turn a recorded model into the flat plan.  theta is the concatenation of the free variables in
    creation order, each on PyMC's unconstrained scale (SURVEY.md §A "theta layouts").
    """
    if len(model.observed_RVs) != 1:
        raise UnsupportedModelError(f'Exactly one observed likelihood is required, found {len(model.observed_RVs)}')

    blocks: List[Block] = []
    block_of = {}
    offset = 0
    for number, rv in enumerate(model.free_RVs):
        family = FAMILIES[rv.family]
        if family.discrete:
            raise UnsupportedModelError(f'{rv.name}: discrete free variables have no gradient')
        block_of[id(rv)] = number
        blocks.append(Block(name=rv.name, family=rv.family, offset=offset, size=rv.size, shape=rv.shape,
                            dims=rv.dims, transform='log' if family.positive else 'identity'))
        offset += rv.size

    for rv, block in zip(model.free_RVs, blocks):
        for key, value in rv.params.items():
            param = _prior_param(value, block_of, block.size, f'{rv.name}.{key}', blocks)
            if not param.is_const:
                if param.is_mixed:
                    absolute = param.index[param.index >= 0]
                    sources = sorted({plan_block_of(blocks, int(j)) for j in np.unique(absolute)})
                else:
                    sources = [param.block]
                if any(src >= block_of[id(rv)] for src in sources):
                    raise UnsupportedModelError(f'{rv.name}.{key}: parameter refers to a later variable')
                if block.family != 'Normal':
                    raise UnsupportedModelError(f'{rv.name}.{key}: only Normal blocks may have random parameters')
                if key == 'sigma' and any(blocks[src].transform != 'log' for src in sources):
                    raise UnsupportedModelError(f'{rv.name}.sigma must be a positive (HalfNormal/Exponential) variable')
            block.params[key] = param

    obs = model.observed_RVs[0]
    _seen.nodes = []
    y_form = _lower(obs.observed)
    if not y_form.is_const:
        raise UnsupportedModelError('Observed data must be constant')
    # a mini-batched likelihood (MinibatchLikelihood) is lowered over the whole data set
    obs_shape = obs.shape if obs.total_size is None else (obs.total_size,)
    y = np.ascontiguousarray(np.broadcast_to(y_form.const_value(), obs_shape), dtype=np.float64).reshape(-1)
    n_rows = y.size
    # rows masked for NaN observations: lowered like all others (so that the indices stay the reference's),
    # then dropped from the plan; what they add to logp does not depend on theta
    obs_params, masked, masked_logp = _masked_likelihood(obs, y)

    spec = LikelihoodSpec(family=obs.family, name=obs.name, y=y)
    row_sigma = None
    if obs.family in ('Normal', 'StudentT'):
        eta_node = obs_params['mu']
        sigma = _lower(obs_params['sigma'])
        student = obs.family == 'StudentT'
        if student:
            nu = _lower(obs_params['nu'])
            value = np.unique(np.asarray(nu.const_value(), dtype=np.float64)) if nu.is_const else np.empty(0)
            if value.size != 1 or not value[0] > 0:
                raise UnsupportedModelError('StudentT nu must be one positive constant')
            spec.nu = float(value[0])
        if sigma.is_const:
            value = np.asarray(sigma.const_value(), dtype=np.float64)
            if masked is not None and value.size == n_rows:
                value = value.reshape(-1)[~masked]
            if np.unique(value).size != 1 or student:   # (StudentT: always standardised, the kernel's sigma is exp(zeta))
                # known measurement errors, one per row (also: a GroupComponent of constants): the rows are
                # standardised -- y, the offset and every data column divided by sigma_i -- which leaves a unit
                # sigma, and sum(-log sigma_i) joins the plan's constant (applied below, once the terms exist)
                if (value <= 0).any():
                    raise ValueError('Normal sigma must be positive')
                row_sigma = np.ascontiguousarray(np.broadcast_to(sigma.const_value(), obs_shape), dtype=np.float64).reshape(-1)
                spec.sigma_const = 1.0
            else:
                spec.sigma_const = float(value.reshape(-1)[0])
        else:
            # per row: a constant, or one positive (log-transformed) element of theta
            gathers = _merge_pieces(AffineForm(sigma.shape, [m for m in sigma.monomials if m.var is not None]),
                                    block_of, blocks)
            const = _merge_constants(sigma)
            const = np.zeros(obs_shape) if const is None else \
                np.array(np.broadcast_to(np.asarray(const, dtype=np.float64), sigma.shape))
            if len(gathers) != 1 or not np.isin(gathers[0].coef, (0.0, 1.0)).all() or \
                    np.any((gathers[0].coef == 1.0) & (const != 0.0)):
                raise UnsupportedModelError('Likelihood sigma must be, per row, a constant or a single positive variable')
            g = gathers[0]
            element = np.broadcast_to(np.where(g.coef == 1.0, g.absolute, -1), obs_shape).reshape(-1)
            const = np.broadcast_to(const, obs_shape).reshape(-1)
            used = np.unique(element[~masked] if masked is not None else element)
            if any(blocks[plan_block_of(blocks, int(e))].transform != 'log' for e in used if e >= 0):
                raise UnsupportedModelError('Likelihood sigma must be a positive (HalfNormal/Exponential ...) variable')
            if ((element < 0) & (const <= 0))[~masked if masked is not None else slice(None)].any():
                raise ValueError('Normal sigma must be positive')
            if used.size == 1 and used[0] >= 0 and not student:
                source = plan_block_of(blocks, int(used[0]))
                spec.sigma_block, spec.sigma_index = source, int(used[0]) - blocks[source].offset
            else:
                # sigma differs between rows (GroupComponent of parameters and constants, or a block gathered by
                # group): kept per row; the kernels gather log(sigma) like a term of the predictor, rows with constant
                # sigmas form a standardised plan of their own (ModelPlan.split_by_sigma, valuegrad.SplitSigmaValueGrad)
                spec.sigma_rows_element = np.ascontiguousarray(element, dtype=np.int64)
                spec.sigma_rows_const = np.ascontiguousarray(const, dtype=np.float64)
    elif obs.family == 'Bernoulli':
        if 'logit_p' in obs_params:
            eta_node = obs_params['logit_p']
        else:
            eta_node = _unwrap_link(obs_params['p'], 'sigmoid', 'Bernoulli p')
        if not np.isin(y, (0.0, 1.0)).all():
            raise ValueError('Bernoulli observations must be 0 or 1')
    elif obs.family in ('Poisson', 'NegativeBinomial'):
        eta_node = _unwrap_link(obs_params['mu'], 'exp', f'{obs.family} mu')
        if (y < 0).any() or (y != np.floor(y)).any():
            raise ValueError(f'{obs.family} observations must be non-negative integers')
        if obs.family == 'NegativeBinomial':
            alpha = _lower(obs_params['alpha'])
            value = np.unique(np.asarray(alpha.const_value(), dtype=np.float64)) if alpha.is_const else np.empty(0)
            if value.size != 1 or not value[0] > 0:
                raise UnsupportedModelError('NegativeBinomial alpha must be one positive constant')
            spec.sigma_const = float(value[0])
    elif obs.family == 'Gamma':
        # the Gamma GLM with a log link: a constant shape alpha and the rate beta = alpha / mu, mu = exp(eta)
        alpha = _lower(obs_params['alpha'])
        value = np.unique(np.asarray(alpha.const_value(), dtype=np.float64)) if alpha.is_const else np.empty(0)
        if value.size != 1 or not value[0] > 0:
            raise UnsupportedModelError('Gamma alpha must be one positive constant')
        spec.sigma_const = float(value[0])
        eta_node = _gamma_rate_predictor(obs_params['beta'], spec.sigma_const)
        if not (y if masked is None else y[~masked] > 0).all() or not (y if masked is None else y[~masked]).min(initial=1.0) > 0:
            raise ValueError('Gamma observations must be positive')
    else:
        raise UnsupportedModelError(f'Likelihood family {obs.family} is not supported')

    eta = _lower(eta_node)
    if np.prod(eta.shape, dtype=np.int64) not in (1, n_rows):
        raise UnsupportedModelError(f'Linear predictor of shape {eta.shape} for {n_rows} observations')
    terms: List[Term] = []
    for m in eta.monomials:
        if m.var is None or m.tag is not None:
            continue
        index = np.ascontiguousarray(np.broadcast_to(m.index, eta.shape), dtype=np.int64).reshape(-1)
        if index.size == 1 and n_rows > 1:
            index = np.repeat(index, n_rows)
        coef = np.asarray(m.coef, dtype=np.float64)
        if coef.ndim == 0 and float(coef) == 1.0:
            x = None
        else:
            x = np.ascontiguousarray(np.broadcast_to(coef, eta.shape), dtype=np.float64).reshape(-1)
            if x.size == 1 and n_rows > 1:
                x = np.repeat(x, n_rows)
        terms.append(Term(block=block_of[id(m.var)], index=index, x=x))
    # member-wise coefficients (GroupComponent): the pieces of one component become one term
    tagged = AffineForm(eta.shape, [m for m in eta.monomials if m.var is not None and m.tag is not None])
    for g in _merge_pieces(tagged, block_of, blocks):
        absolute = np.ascontiguousarray(g.absolute, dtype=np.int64).reshape(-1)
        coef = g.coef.reshape(-1)
        if absolute.size == 1 and n_rows > 1:
            absolute, coef = np.repeat(absolute, n_rows), np.repeat(coef, n_rows)
        x = None if np.all(coef == 1.0) else np.ascontiguousarray(coef)
        if len(g.blocks) == 1:
            terms.append(Term(block=g.blocks[0], index=absolute - blocks[g.blocks[0]].offset, x=x))
        else:
            terms.append(Term(block=-1, index=absolute, x=x))     # different blocks in different rows
    const = _merge_constants(eta)
    if const is not None and np.any(np.asarray(const) != 0.0):
        spec.offset = np.ascontiguousarray(np.broadcast_to(const, (n_rows,)), dtype=np.float64)

    plan = ModelPlan(blocks=blocks, terms=terms, likelihood=spec, n_rows=n_rows,
                     coords={k: np.asarray(v) for k, v in model.coords.items()})
    if obs.total_size is not None:
        # (the observed variable itself keeps the shape of the whole group, reference test_minibatch.py:22; the size
        # of the batch is that of the index variable its parameters and observations are gathered with)
        if len({id(mb) for mb in _seen.nodes}) != 1 or _seen.nodes[0].n != n_rows:
            raise UnsupportedModelError('A mini-batched likelihood needs one mini-batch variable over all observations')
        plan.minibatch = int(_seen.nodes[0].batch_size)
    elif _seen.nodes:
        raise UnsupportedModelError('Mini-batched components outside a MinibatchLikelihood')
    if row_sigma is not None:
        weight = 1.0 / row_sigma
        spec.y = spec.y * weight
        spec.offset = None if spec.offset is None else spec.offset * weight
        for t in terms:
            t.x = weight.copy() if t.x is None else t.x * weight
        plan.logp_const -= float(np.sum(np.log(row_sigma if masked is None else row_sigma[~masked]), dtype=np.float64))
    if masked is not None:
        keep = ~masked
        spec.y = np.ascontiguousarray(spec.y[keep])
        if spec.offset is not None:
            spec.offset = np.ascontiguousarray(spec.offset[keep])
        for t in terms:
            t.index = np.ascontiguousarray(t.index[keep])
            if t.x is not None:
                t.x = np.ascontiguousarray(t.x[keep])
        if spec.sigma_rows_element is not None:
            spec.sigma_rows_element = np.ascontiguousarray(spec.sigma_rows_element[keep])
            spec.sigma_rows_const = np.ascontiguousarray(spec.sigma_rows_const[keep])
        plan.n_rows = int(keep.sum())
        plan.logp_const += masked_logp
        plan.n_masked_rows = int(masked.sum())
    return plan


def plan_block_of(blocks: List[Block], element: int) -> int:
    """Number of the block that holds element ``element`` of theta."""
    for number, b in enumerate(blocks):
        if b.offset <= element < b.offset + b.size:
            return number
    raise IndexError(element)


def _gamma_rate_predictor(node: Any, alpha: float):
    """
    The linear predictor eta of a Gamma likelihood's rate ``beta = alpha / exp(eta)``: accepts ``c / exp(L)`` and
    ``c * exp(L)`` (either order) with a positive constant ``c``; a ``c`` other than alpha becomes a constant of eta
    (``beta = alpha / exp(eta)`` with ``eta = L + log(alpha / c)``, resp. ``-L + log(alpha / c)``).
    """
    while isinstance(node, sym.DeterministicVar):
        node = node.var
    if isinstance(node, sym.Gather):
        # (a rate without any observation-level term is formed on the dense tensor of its factors and gathered to the
        #  observations afterwards, see _unwrap_link: the elementwise expression commutes with the gather)
        return sym.Gather(_gamma_rate_predictor(node.base, alpha), node.indices)
    what = 'Gamma beta must be alpha / exp(linear predictor) for the fused kernel'
    if not isinstance(node, sym.BinOp) or node.op not in ('truediv', 'mul'):
        raise UnsupportedModelError(what)

    def exp_arg(n):
        while isinstance(n, sym.DeterministicVar):
            n = n.var
        return n.arg if isinstance(n, sym.UnaryOp) and n.op == 'exp' else None

    def constant(n):
        try:
            form = _lower(n)
        except UnsupportedModelError:
            return None
        if not form.is_const:
            return None
        value = np.unique(np.asarray(form.const_value(), dtype=np.float64))
        return float(value[0]) if value.size == 1 and value[0] > 0 else None

    if node.op == 'truediv':
        c, inner, sign = constant(node.left), exp_arg(node.right), 1.0
    else:
        c, inner, sign = constant(node.left), exp_arg(node.right), -1.0
        if c is None or inner is None:
            c, inner = constant(node.right), exp_arg(node.left)
    if c is None or inner is None:
        raise UnsupportedModelError(what)
    eta = inner if sign > 0 else sym.UnaryOp('neg', inner)
    shift = float(np.log(alpha / c))
    return eta if shift == 0.0 else sym.BinOp('add', eta, shift)


def _unwrap_link(node: Any, link: str, what: str):
    while isinstance(node, sym.DeterministicVar):
        node = node.var
    if isinstance(node, sym.UnaryOp) and node.op == link:
        return node.arg
    if isinstance(node, sym.Gather):
        # A predictor without any observation-level term (``f_(exp)(a + b)`` with a, b on crossed factors) is evaluated
        # on the dense [A, B] tensor and only then gathered to the observations (the reference's minimal representations,
        # sakkara/relation/representation.py:277-290): an elementwise link commutes with the gather
        base = node.base
        while isinstance(base, sym.DeterministicVar):
            base = base.var
        if isinstance(base, sym.UnaryOp) and base.op == link:
            return sym.Gather(base.arg, node.indices)
    raise UnsupportedModelError(f'{what} must be {link}(linear predictor) for the fused kernel')
