"""
Minimal symbolic backend used by ``build()`` in place of a PyMC/PyTensor graph.

The reference hands every component to PyMC, which records random variables and builds a
PyTensor graph out of ``element[index arrays]`` gathers (``sakkara/relation/representation.py:253``)
and elementwise arithmetic (``sakkara/model/function/base.py:88-92``).  The B200 path needs exactly
that information -- creation order of the free variables (= layout of theta), their priors, the
gather indices and the arithmetic of the linear predictor -- and nothing else of PyMC, so this
module records it in a few node classes that :mod:`sakkara_b200.plan.lower` pattern-matches into
the flat kernel plan.
"""
import operator
import threading
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class Var:
    """Base class of symbolic tensors.  Only shape-preserving elementwise ops and gathers exist."""
    shape: Tuple[int, ...] = ()

    def __getitem__(self, index) -> 'Var':
        if not isinstance(index, tuple):
            index = (index,)
        if len(index) == 1 and isinstance(index[0], MinibatchIndex):
            return MinibatchGather(self, index[0])
        arrays = tuple(np.asarray(i) for i in index)
        if len(arrays) == 1 and arrays[0].dtype == bool:
            # boolean mask over the whole tensor (reference ``composable/group.py:93``: ``variable[mask]``)
            if arrays[0].shape != tuple(self.shape):
                raise IndexError(f'mask of shape {arrays[0].shape} for a tensor of shape {tuple(self.shape)}')
            arrays = np.nonzero(arrays[0])
        return Gather(self, tuple(np.asarray(i) for i in arrays))

    # shape changes are gathers with the identity permutation (row-major, like PyTensor's)
    def ravel(self) -> 'Var':
        if len(self.shape) == 1:
            return self
        size = int(np.prod(self.shape)) if len(self.shape) else 1
        return Gather(self, tuple(np.asarray(i) for i in np.unravel_index(np.arange(size), self.shape)))

    def flatten(self) -> 'Var':
        return self.ravel()

    def reshape(self, *shape) -> 'Var':
        if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)):
            shape = tuple(shape[0])
        flat = self.ravel()
        size = flat.shape[0]
        return Gather(flat, (np.arange(size).reshape(shape),))

    # arithmetic -----------------------------------------------------------------------------
    def __add__(self, other): return BinOp('add', self, other)
    def __radd__(self, other): return BinOp('add', other, self)
    def __sub__(self, other): return BinOp('sub', self, other)
    def __rsub__(self, other): return BinOp('sub', other, self)
    def __mul__(self, other): return BinOp('mul', self, other)
    def __rmul__(self, other): return BinOp('mul', other, self)
    def __truediv__(self, other): return BinOp('truediv', self, other)
    def __rtruediv__(self, other): return BinOp('truediv', other, self)
    def __pow__(self, other, modulo=None): return BinOp('pow', self, other)
    def __rpow__(self, other, modulo=None): return BinOp('pow', other, self)
    def __mod__(self, other): return BinOp('mod', self, other)
    def __rmod__(self, other): return BinOp('mod', other, self)
    def __floordiv__(self, other): return BinOp('floordiv', self, other)
    def __rfloordiv__(self, other): return BinOp('floordiv', other, self)
    def __neg__(self): return UnaryOp('neg', self)

    # numpy must not try to broadcast over a Var when it is the left operand
    __array_ufunc__ = None


def shape_of(x: Any) -> Tuple[int, ...]:
    if isinstance(x, Var):
        return x.shape
    return np.shape(x)


def broadcast_shape(a: Any, b: Any) -> Tuple[int, ...]:
    return np.broadcast_shapes(shape_of(a), shape_of(b))


class Gather(Var):
    """``base[index_0, index_1, ...]`` with integer index arrays of one common shape."""

    def __init__(self, base: Var, indices: Tuple[np.ndarray, ...]):
        if len(indices) != len(base.shape):
            raise IndexError(f'{len(indices)} index arrays for a tensor of rank {len(base.shape)}')
        self.base = base
        self.indices = indices
        self.shape = np.broadcast_shapes(*[i.shape for i in indices])


class BinOp(Var):
    OPS = {'add': operator.add, 'sub': operator.sub, 'mul': operator.mul, 'truediv': operator.truediv,
           'pow': operator.pow, 'mod': operator.mod, 'floordiv': operator.floordiv}

    def __init__(self, op: str, left: Any, right: Any):
        self.op = op
        self.left = left
        self.right = right
        self.shape = broadcast_shape(left, right)


class UnaryOp(Var):
    def __init__(self, op: str, arg: Any):
        self.op = op
        self.arg = arg
        self.shape = shape_of(arg)


class Masked(Var):
    """
    ``fill`` in the cells where ``mask`` is true, ``value`` elsewhere: what the reference's
    per-row ``GroupComponent`` composes for the parameters of a likelihood whose observations
    contain NaN (``composable/hierarchical/likelihood.py:41-53``: the rows with NaN get the
    constant ``nan_param_mask[param]``, all other rows the parameter's own component).
    """

    def __init__(self, value: Any, mask: np.ndarray, fill: Any):
        self.value = value
        self.mask = np.asarray(mask, dtype=bool)
        self.fill = fill
        self.shape = np.broadcast_shapes(shape_of(value), self.mask.shape)


class Concat(Var):
    """
    One-dimensional concatenation of one-dimensional pieces: what ``pt.stack(pieces).ravel()`` is for
    the equally long member-wise pieces of a ``GroupComponent`` (reference ``composable/group.py:95-100``).
    A piece may be a ``Var`` or plain numbers.
    """

    def __init__(self, pieces):
        self.pieces = list(pieces)
        lengths = []
        for piece in self.pieces:
            shape = shape_of(piece)
            if len(shape) != 1:
                raise ValueError(f'Concat takes one-dimensional pieces, got shape {shape}')
            lengths.append(int(shape[0]))
        self.lengths = lengths
        self.shape = (int(sum(lengths)),)


def stack(pieces) -> Var:
    """``pt.stack`` of equally long one-dimensional pieces: a ``[len(pieces), length]`` tensor."""
    pieces = list(pieces)
    cat = Concat(pieces)
    if len(set(cat.lengths)) > 1:
        raise ValueError(f'stack needs pieces of equal length, got {cat.lengths}')
    return cat.reshape(len(pieces), cat.lengths[0] if cat.lengths else 0)


class MinibatchIndex(Var):
    """
    ``batch_size`` row numbers drawn uniformly (with replacement) from ``range(n)``, redrawn at every evaluation:
    what ``pymc.data.minibatch_index(0, n, size=(batch_size,))`` is in the reference
    (``sakkara/relation/group.py:54-60``).
    """

    def __init__(self, n: int, batch_size: int):
        self.n = int(n)
        self.batch_size = int(batch_size)
        self.shape = (self.batch_size,)


class MinibatchGather(Var):
    """``base[minibatch]``: the rows of the current minibatch (reference ``sakkara/model/minibatch.py:31-33``)."""

    def __init__(self, base: Var, minibatch: MinibatchIndex):
        if len(shape_of(base)) < 1 or shape_of(base)[0] != minibatch.n:
            raise IndexError(f'minibatch over {minibatch.n} members for a tensor of shape {shape_of(base)}')
        self.base = base
        self.minibatch = minibatch
        self.shape = (minibatch.batch_size,) + tuple(shape_of(base)[1:])


class DataVar(Var):
    """Constant data registered in the model (reference: ``pm.ConstantData``, ``fixed/data.py:41``)."""

    def __init__(self, name: str, values: np.ndarray):
        self.name = name
        self.values = np.asarray(values)
        self.shape = self.values.shape


class RandomVar(Var):
    """
    One distribution call recorded in a :class:`Model`: a free parameter block, or the observed
    likelihood when ``observed`` is given (reference: ``hierarchical/distribution.py:49-51``).
    """

    def __init__(self, family: str, name: str, params: Dict[str, Any], shape: Tuple[int, ...],
                 dims: Optional[Tuple[str, ...]], observed: Any = None, total_size: Optional[int] = None):
        self.total_size = None if total_size is None else int(total_size)   # minibatch: rows of the whole data set
        self.family = family
        self.name = name
        self.params = params
        self.shape = tuple(int(s) for s in shape)
        self.dims = dims
        self.observed = observed

    @property
    def is_observed(self) -> bool:
        return self.observed is not None

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if len(self.shape) else 1


class DeterministicVar(Var):
    """Named alias of another variable (reference: ``pm.Deterministic``); adds no logp term."""

    def __init__(self, name: str, var: Any, dims=None):
        self.name = name
        self.var = var
        self.dims = dims
        self.shape = shape_of(var)


# ---------------------------------------------------------------------------------------------
# elementwise functions usable inside f_(...) in place of pytensor.tensor.exp etc.
# ---------------------------------------------------------------------------------------------
class ElementwiseFunction:
    def __init__(self, name: str):
        self.name = name
        self.__name__ = name

    def __call__(self, x):
        return UnaryOp(self.name, x)

    def __repr__(self):
        return self.name


exp = ElementwiseFunction('exp')
sigmoid = ElementwiseFunction('sigmoid')
invlogit = ElementwiseFunction('sigmoid')


# ---------------------------------------------------------------------------------------------
# model context
# ---------------------------------------------------------------------------------------------
_context = threading.local()


class Model:
    """
    Recording model context, the counterpart of ``pm.Model(coords=...)`` in
    ``sakkara/model/utils.py:31``.  Variables register themselves in creation order.
    """

    def __init__(self, coords: Optional[Dict[str, Any]] = None):
        self.coords = dict(coords or {})
        self.free_RVs: List[RandomVar] = []
        self.observed_RVs: List[RandomVar] = []
        self.data_vars: List[DataVar] = []
        self.deterministics: List[DeterministicVar] = []
        self.named_vars: Dict[str, Var] = {}

    def __enter__(self) -> 'Model':
        stack = getattr(_context, 'stack', None)
        if stack is None:
            stack = _context.stack = []
        stack.append(self)
        return self

    def __exit__(self, *exc) -> None:
        _context.stack.pop()

    @staticmethod
    def current() -> 'Model':
        stack = getattr(_context, 'stack', None)
        if not stack:
            raise TypeError('No model on context stack; distributions must be created inside build()')
        return stack[-1]

    def register(self, var: Var) -> Var:
        name = var.name
        if name in self.named_vars:
            # same rule as PyMC: names are unique within a model
            raise ValueError(f'Variable name {name} already exists.')
        self.named_vars[name] = var
        if isinstance(var, RandomVar):
            (self.observed_RVs if var.is_observed else self.free_RVs).append(var)
        elif isinstance(var, DataVar):
            self.data_vars.append(var)
        elif isinstance(var, DeterministicVar):
            self.deterministics.append(var)
        return var


def ConstantData(name: str, values) -> DataVar:
    return Model.current().register(DataVar(name, values))


def Deterministic(name: str, var, dims=None) -> DeterministicVar:
    return Model.current().register(DeterministicVar(name, var, dims))
