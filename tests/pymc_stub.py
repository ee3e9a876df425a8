"""
TEST INFRASTRUCTURE -- a small executable stand-in for the parts of PyTensor / PyMC that
``sakkara_b200/pytensor_op.py`` talks to (neither is installable in this image).

It is *not* a recording fake: graphs built against it can be evaluated and differentiated, so the
tests can check what ``pm.sample()`` would see -- the raveled order of the value variables, the
``<name>_log__`` value variables of positive blocks, the log-Jacobian PyMC adds for ``HalfFlat`` (and
that the Potential subtracts again), and that logp and dlogp of one point cost ONE call of the
evaluator (PyTensor's merge rewrite joins the two applications of the Op; here applications with
the same op and inputs are evaluated once).

Implements, with the documented PyMC 5 / PyTensor 2 semantics:
  pytensor.graph.basic.Apply, pytensor.graph.op.Op (make_node / perform / grad / __call__),
  pytensor.tensor: as_tensor_variable, scalar, vector, log, exp, concatenate; Variable methods
  astype / ravel / sum / + - * ; pytensor.grad;
  pymc: Model(coords) context, Flat, HalfFlat (log transform), Potential, ConstantData,
  Model.logp_dlogp_function() -> callable(raveled value vars) -> (logp, dlogp).

install() puts the modules into sys.modules, uninstall() removes them again.
"""
import sys
import types

import numpy as np


# ---------------------------------------------------------------------------------------------
# graph
# ---------------------------------------------------------------------------------------------
class Variable:
    def __init__(self, dtype='float64', ndim=0, name=None, value=None):
        self.dtype, self.ndim, self.name = str(dtype), ndim, name
        self.owner, self.index, self.value = None, 0, value     # value: constants only

    def astype(self, dtype): return self if str(dtype) == self.dtype else Cast(dtype)(self)   # (a no-op cast returns the variable itself, as in PyTensor)
    def ravel(self): return Ravel()(self)
    def sum(self): return Sum()(self)
    def __add__(self, o): return Elemwise('add')(self, as_tensor_variable(o))
    def __radd__(self, o): return Elemwise('add')(as_tensor_variable(o), self)
    def __sub__(self, o): return Elemwise('sub')(self, as_tensor_variable(o))
    def __rsub__(self, o): return Elemwise('sub')(as_tensor_variable(o), self)
    def __mul__(self, o): return Elemwise('mul')(self, as_tensor_variable(o))
    def __rmul__(self, o): return Elemwise('mul')(as_tensor_variable(o), self)


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, out in enumerate(self.outputs):
            out.owner, out.index = self, i


class Op:
    __props__ = ()

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else node.outputs

    def make_node(self, *inputs):
        raise NotImplementedError

    def perform(self, node, inputs, outputs):
        raise NotImplementedError

    def grad(self, inputs, output_grads):
        raise NotImplementedError

    def _key(self):
        return (type(self),) + tuple(getattr(self, p) for p in self.__props__)


def as_tensor_variable(x):
    if isinstance(x, Variable):
        return x
    value = np.asarray(x, dtype=np.float64)
    return Variable('float64', value.ndim, value=value)


def scalar(name=None, dtype='float64'): return Variable(dtype, 0, name)
def vector(name=None, dtype='float64'): return Variable(dtype, 1, name)


class Elemwise(Op):
    __props__ = ('fn',)
    FN = {'add': np.add, 'sub': np.subtract, 'mul': np.multiply, 'log': np.log, 'exp': np.exp}

    def __init__(self, fn): self.fn = fn

    def make_node(self, *inputs):
        return Apply(self, inputs, [Variable(inputs[0].dtype, max(i.ndim for i in inputs))])

    def perform(self, node, inputs, outputs):
        outputs[0][0] = self.FN[self.fn](*inputs)

    def grad(self, inputs, gouts):
        g = gouts[0]
        if self.fn == 'add': return [Unbroadcast()(g, inputs[0]), Unbroadcast()(g, inputs[1])]
        if self.fn == 'sub': return [Unbroadcast()(g, inputs[0]), Unbroadcast()(g * -1.0, inputs[1])]
        if self.fn == 'mul': return [Unbroadcast()(g * inputs[1], inputs[0]), Unbroadcast()(g * inputs[0], inputs[1])]
        if self.fn == 'log': return [g * Elemwise('exp')(Elemwise('log')(inputs[0]) * -1.0)]
        if self.fn == 'exp': return [g * Elemwise('exp')(inputs[0])]


class Unbroadcast(Op):
    """sum a gradient back to the shape of the (possibly broadcast) operand"""
    def make_node(self, g, like): return Apply(self, [g, like], [Variable(g.dtype, like.ndim)])

    def perform(self, node, inputs, outputs):
        g, like = (np.asarray(v) for v in inputs)
        while g.ndim > like.ndim:
            g = g.sum(axis=0)
        for ax, n in enumerate(like.shape):
            if n == 1 and g.shape[ax] != 1:
                g = g.sum(axis=ax, keepdims=True)
        outputs[0][0] = np.broadcast_to(g, like.shape).copy()


class Cast(Op):
    __props__ = ('dtype',)
    def __init__(self, dtype): self.dtype = str(dtype)
    def make_node(self, x): return Apply(self, [x], [Variable(self.dtype, x.ndim)])
    def perform(self, node, inputs, outputs): outputs[0][0] = np.asarray(inputs[0], dtype=self.dtype)
    def grad(self, inputs, gouts): return [gouts[0]]


class Ravel(Op):
    def make_node(self, x): return Apply(self, [x], [Variable(x.dtype, 1)])
    def perform(self, node, inputs, outputs): outputs[0][0] = np.ravel(inputs[0])
    def grad(self, inputs, gouts): return [ReshapeLike()(gouts[0], inputs[0])]


class ReshapeLike(Op):
    def make_node(self, g, like): return Apply(self, [g, like], [Variable(g.dtype, like.ndim)])
    def perform(self, node, inputs, outputs): outputs[0][0] = np.reshape(inputs[0], np.shape(inputs[1]))


class Sum(Op):
    def make_node(self, x): return Apply(self, [x], [Variable(x.dtype, 0)])
    def perform(self, node, inputs, outputs): outputs[0][0] = np.sum(inputs[0])
    def grad(self, inputs, gouts): return [Unbroadcast()(gouts[0] * Elemwise('exp')(inputs[0] * 0.0), inputs[0])]


class Join(Op):
    def make_node(self, *xs): return Apply(self, xs, [Variable(xs[0].dtype, 1)])
    def perform(self, node, inputs, outputs): outputs[0][0] = np.concatenate([np.atleast_1d(v) for v in inputs])
    def grad(self, inputs, gouts): return [Split(i)(gouts[0], *inputs) for i in range(len(inputs))]


class Split(Op):
    __props__ = ('which',)
    def __init__(self, which): self.which = which
    def make_node(self, g, *xs): return Apply(self, [g, *xs], [Variable(g.dtype, 1)])

    def perform(self, node, inputs, outputs):
        sizes = [np.size(v) for v in inputs[1:]]
        lo = int(np.sum(sizes[:self.which]))
        outputs[0][0] = np.asarray(inputs[0])[lo:lo + sizes[self.which]]


def log(x): return Elemwise('log')(as_tensor_variable(x))
def exp(x): return Elemwise('exp')(as_tensor_variable(x))
def concatenate(xs): return Join()(*[as_tensor_variable(x) for x in xs])


PERFORM_CALLS = {'count': 0}


def evaluate(outputs, givens):
    """Values of `outputs` given {input variable: value}; applications with the same op and inputs run once."""
    values = {id(k): np.asarray(v) for k, v in givens.items()}
    merged = {}

    def value_of(var):
        if id(var) in values:
            return values[id(var)]
        if var.owner is None:
            if var.value is None:
                raise KeyError(f'no value for input {var.name}')
            return var.value
        node = var.owner
        ins = [value_of(i) for i in node.inputs]
        key = (node.op._key() if node.op.__props__ or type(node.op).__module__ == __name__ else id(node.op),
               tuple(id(i) for i in node.inputs))
        if key not in merged:
            storage = [[None] for _ in node.outputs]
            node.op.perform(node, ins, storage)
            PERFORM_CALLS['count'] += 1
            merged[key] = [s[0] for s in storage]
        for out, val in zip(node.outputs, merged[key]):
            values[id(out)] = val
        return values[id(var)]

    return [value_of(o) for o in outputs]


def grad(cost, wrt):
    """Reverse mode over the graph of `cost`; returns d cost / d w for each w in wrt."""
    order, seen = [], set()

    def visit(var):
        if var.owner is None or id(var.owner) in seen:
            return
        seen.add(id(var.owner))
        for i in var.owner.inputs:
            visit(i)
        order.append(var.owner)

    visit(cost)
    acc = {id(cost): as_tensor_variable(1.0)}
    for node in reversed(order):
        gouts = [acc.get(id(o)) for o in node.outputs]
        if all(g is None for g in gouts):
            continue
        gouts = [g if g is not None else as_tensor_variable(0.0) for g in gouts]
        if isinstance(node.op, (Unbroadcast, ReshapeLike, Split)):
            gins = [None] * len(node.inputs)         # helper ops of the backward pass are not differentiated
        else:
            gins = node.op.grad(node.inputs, gouts)
        for i, g in zip(node.inputs, gins):
            if g is not None:
                acc[id(i)] = g if id(i) not in acc else acc[id(i)] + g
    return [acc[id(w)] for w in wrt]


# ---------------------------------------------------------------------------------------------
# pymc
# ---------------------------------------------------------------------------------------------
class Model:
    _stack = []

    def __init__(self, coords=None):
        self.coords = dict(coords or {})
        self.free_RVs, self.value_vars, self.potentials, self.jacobians = [], [], [], []
        self.named_vars, self.rvs_to_values, self.named_vars_to_dims, self.data = {}, {}, {}, {}

    def __enter__(self):
        Model._stack.append(self)
        return self

    def __exit__(self, *exc):
        Model._stack.pop()

    @classmethod
    def get_context(cls):
        if not cls._stack:
            raise TypeError('No model on context stack.')
        return cls._stack[-1]

    def _register(self, name, rv, value_var, shape, dims):
        if name in self.named_vars:
            raise ValueError(f'Variable name {name} already exists.')
        if dims is not None:
            want = tuple(len(self.coords[d]) for d in dims)
            if shape is not None and tuple(shape) != want:
                raise ValueError(f'{name}: shape {shape} does not match dims {dims} = {want}')
            shape = want
        value_var.shape_hint = tuple(shape if shape is not None else ())
        rv.name = name
        self.named_vars[name] = rv
        self.named_vars_to_dims[name] = dims
        self.free_RVs.append(rv)
        self.value_vars.append(value_var)
        self.rvs_to_values[name] = value_var
        return rv

    def logp_dlogp_function(self):
        cost = as_tensor_variable(0.0)
        for term in self.potentials + self.jacobians:
            cost = cost + term
        grads = grad(cost, self.value_vars)
        sizes = [int(np.prod(v.shape_hint)) if v.shape_hint else 1 for v in self.value_vars]

        def f(q):
            q = np.asarray(q, dtype=np.float64)
            givens, lo = {}, 0
            for v, n in zip(self.value_vars, sizes):
                givens[v] = q[lo:lo + n].reshape(v.shape_hint)
                lo += n
            out = evaluate([cost] + grads, givens)
            return float(out[0]), np.concatenate([np.ravel(np.broadcast_to(g, v.shape_hint))
                                                  for g, v in zip(out[1:], self.value_vars)])
        f.size = sum(sizes)
        f.value_var_names = [v.name for v in self.value_vars]
        return f


def Flat(name, shape=None, dims=None):
    value = Variable('float64', len(shape or dims or ()), name)
    return Model.get_context()._register(name, value, value, shape, dims)


def HalfFlat(name, shape=None, dims=None):
    """positive support: sampled through the log transform, value variable <name>_log__, log-Jacobian = u"""
    model = Model.get_context()
    value = Variable('float64', len(shape or dims or ()), name + '_log__')
    rv = exp(value)
    model.jacobians.append(value.sum())
    return model._register(name, rv, value, shape, dims)


def Potential(name, var):
    model = Model.get_context()
    model.potentials.append(var)
    model.named_vars[name] = var
    return var


def ConstantData(name, values, dims=None):
    model = Model.get_context()
    model.data[name] = np.asarray(values)
    return as_tensor_variable(np.asarray(values, dtype=np.float64))


_NAMES = ('pymc', 'pytensor', 'pytensor.tensor', 'pytensor.graph', 'pytensor.graph.basic', 'pytensor.graph.op')


def install():
    """Register the stand-ins as `pymc` / `pytensor` (a real installation takes precedence)."""
    this = sys.modules[__name__]
    pm = types.ModuleType('pymc')
    for n in ('Model', 'Flat', 'HalfFlat', 'Potential', 'ConstantData'):
        setattr(pm, n, getattr(this, n))
    pm.__version__ = '5.stub'
    pt = types.ModuleType('pytensor.tensor')
    for n in ('as_tensor_variable', 'scalar', 'vector', 'log', 'exp', 'concatenate'):
        setattr(pt, n, getattr(this, n))
    pt.TensorVariable = Variable
    pytensor = types.ModuleType('pytensor')
    pytensor.tensor, pytensor.grad = pt, grad
    graph = types.ModuleType('pytensor.graph')
    basic = types.ModuleType('pytensor.graph.basic')
    basic.Apply = Apply
    opmod = types.ModuleType('pytensor.graph.op')
    opmod.Op = Op
    graph.basic, graph.op = basic, opmod
    pytensor.graph = graph
    for name, mod in zip(_NAMES, (pm, pytensor, pt, graph, basic, opmod)):
        mod.__skk_stub__ = True
        sys.modules[name] = mod
    return pm


def uninstall():
    for name in _NAMES:
        if getattr(sys.modules.get(name), '__skk_stub__', False):
            del sys.modules[name]
