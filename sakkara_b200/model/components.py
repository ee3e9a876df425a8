"""
Component DSL: the user-facing classes of ``sakkara.model``
re-implemented against the recording backend in :mod:`sakkara_b200.model.sym`.

The public surface -- constructor signatures, the ``retrieve_groups -> prebuild ->
build_representation -> build_variable`` life cycle (reference ``sakkara/model/base.py:72-80``),
automatic naming (``composable/base.py:59-60``, ``function/base.py:45-52``), depth-first build
order in kwargs / operand order (= layout of theta) and ``clear()``
semantics (``tests/workflows/test_linreg.py:101-113``) -- is what model code written for the
reference relies on, so it is kept.  What differs is what a built ``variable`` is: a node of the
small symbolic graph that the plan lowering consumes, not a PyTensor variable.
"""
import operator
import warnings
from itertools import chain
from typing import Any, Callable, Dict, Optional, Set, Tuple, Union

import numpy as np
import pandas as pd

from sakkara_b200.model import dist as _dist
from sakkara_b200.model import sym
from sakkara_b200.relation.groupset import GroupSet
from sakkara_b200.relation.representation import MinimalTensorRepresentation, UnrepeatableRepresentation

GroupSpec = Optional[Union[str, Tuple[str, ...]]]


def _as_group_tuple(group: GroupSpec) -> Tuple[str, ...]:
    if group is None:
        return tuple()
    if isinstance(group, str):
        return (group,)
    return tuple(group)


class ModelComponent:
    """
    Node of the component graph.  Arithmetic between components (and between a component and a
    plain number) yields a :class:`FunctionComponent` (reference ``math_op.py:9-57``).
    """

    def __init__(self):
        self.representation = None
        self.variable = None

    # -- to be provided by subclasses --------------------------------------------------------
    def get_name(self) -> Optional[str]:
        raise NotImplementedError

    def set_name(self, name: str) -> None:
        raise NotImplementedError

    def clear(self) -> None:
        raise NotImplementedError

    def retrieve_groups(self) -> Set[str]:
        raise NotImplementedError

    def prebuild(self, groupset: GroupSet) -> None:
        raise NotImplementedError

    def build_representation(self, groupset: GroupSet) -> None:
        raise NotImplementedError

    def build_variable(self) -> None:
        raise NotImplementedError

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        """The component's counterpart for mini-batches of ``group`` (reference ``base.py:65-70``)."""
        raise NotImplementedError

    # -- life cycle ----------------------------------------------------------------------------
    def build(self, groupset: GroupSet) -> None:
        self.prebuild(groupset)
        self.build_representation(groupset)
        self.build_variable()

    # -- arithmetic ----------------------------------------------------------------------------
    def __add__(self, other): return FunctionComponent.math_op(operator.add, self, other)
    def __radd__(self, other): return FunctionComponent.math_op(operator.add, other, self)
    def __sub__(self, other): return FunctionComponent.math_op(operator.sub, self, other)
    def __rsub__(self, other): return FunctionComponent.math_op(operator.sub, other, self)
    def __mul__(self, other): return FunctionComponent.math_op(operator.mul, self, other)
    def __rmul__(self, other): return FunctionComponent.math_op(operator.mul, other, self)
    def __pow__(self, power, modulo=None): return FunctionComponent.math_op(operator.pow, self, power)
    def __rpow__(self, other, modulo=None): return FunctionComponent.math_op(operator.pow, other, self)
    def __mod__(self, other): return FunctionComponent.math_op(operator.mod, self, other)
    def __rmod__(self, other): return FunctionComponent.math_op(operator.mod, other, self)
    def __floordiv__(self, other): return FunctionComponent.math_op(operator.floordiv, self, other)
    def __rfloordiv__(self, other): return FunctionComponent.math_op(operator.floordiv, other, self)
    def __truediv__(self, other): return FunctionComponent.math_op(operator.truediv, self, other)
    def __rtruediv__(self, other): return FunctionComponent.math_op(operator.truediv, other, self)
    def __neg__(self): return FunctionComponent(operator.neg, None, self)


# =============================================================================================
# expression nodes
# =============================================================================================
class FunctionComponent(ModelComponent):
    """
    Intermediate state of an operation between components (reference ``function/base.py:11-157``).
    Arguments are mapped to the minimal common representation, then ``fct`` is applied.
    """

    def __init__(self, fct: Callable, output_group: Optional[Tuple[str, ...]], *args: ModelComponent,
                 **kwargs: ModelComponent):
        super().__init__()
        self.fct = fct
        self.output_group = output_group
        self.args = args
        self.kwargs = kwargs
        self.input_representation = None

    def _operands(self):
        return chain(self.args, self.kwargs.values())

    def get_name(self) -> Optional[str]:
        if any(c.get_name() is None for c in self._operands()):
            return None
        return str(self.fct)

    def set_name(self, name: str) -> None:
        stem = name + '_' + str(self.fct)
        for position, comp in enumerate(self.args):
            if comp.get_name() is None:
                comp.set_name(f'{stem}_arg{position}')
        for key, comp in self.kwargs.items():
            if comp.get_name() is None:
                comp.set_name(f'{stem}_{key}')

    def clear(self) -> None:
        self.variable = None
        self.representation = None
        for comp in self._operands():
            comp.clear()

    def retrieve_groups(self) -> Set[str]:
        groups: Set[str] = set()
        for comp in self._operands():
            groups |= comp.retrieve_groups()
        return groups

    def prebuild(self, groupset: GroupSet) -> None:
        # Operands without a name are postponed: building a sibling first may name them
        # (reference function/base.py:60-76, same retry budget).
        pending = list(self._operands())
        n_operands = len(pending)
        attempts = 0
        while pending:
            comp = pending.pop(0)
            if comp.get_name() is None:
                if attempts < n_operands * (1 + n_operands) / 2:
                    pending.append(comp)
                else:
                    raise ValueError('All arguments to must be named')
            elif comp.variable is None:
                comp.build(groupset)
            attempts += 1

    def build_representation(self, groupset: GroupSet) -> None:
        joint = MinimalTensorRepresentation()
        for comp in self._operands():
            joint = MinimalTensorRepresentation(*joint.get_groups(), *comp.representation.get_groups())
        self.input_representation = joint
        if self.output_group is None:
            self.representation = joint
        else:
            self.representation = MinimalTensorRepresentation(*[groupset[g] for g in self.output_group])

    def build_variable(self) -> None:
        target = self.input_representation
        args = tuple(c.representation.map(c.variable, target) for c in self.args)
        kwargs = {k: c.representation.map(c.variable, target) for k, c in self.kwargs.items()}
        self.variable = self.fct(*args, **kwargs)

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        # the same function over the mini-batched operands (reference function/base.py:100-104)
        return FunctionComponent(self.fct, self.output_group, *[c.to_minibatch(batch_size, group) for c in self.args],
                                 **{k: c.to_minibatch(batch_size, group) for k, c in self.kwargs.items()})

    @staticmethod
    def math_op(fct: Callable, left: Any, right: Any) -> 'FunctionComponent':
        left_is_comp = isinstance(left, ModelComponent)
        if left_is_comp and isinstance(right, ModelComponent):
            return FunctionComponent(fct, None, left, right)
        if left_is_comp:
            return FunctionComponent(lambda x: fct(x, right), None, left)
        return FunctionComponent(lambda x: fct(left, x), None, right)


class f_:
    """
    Wrap a generic function so that it accepts components (reference ``function/wrapper.py:7-40``).
    For the fused kernel only elementwise links are lowerable: use ``sakkara_b200.model.sym.exp``
    (or ``pytensor.tensor.exp``) around a linear predictor.
    """

    def __init__(self, fct: Callable, output_group: GroupSpec = None):
        self.fct = fct
        self.output_group = (output_group,) if isinstance(output_group, str) else output_group

    def __call__(self, *args, **kwargs) -> FunctionComponent:
        wrapped_args = [a if isinstance(a, ModelComponent) else UnrepeatableComponent(a) for a in args]
        wrapped_kwargs = {k: a if isinstance(a, ModelComponent) else UnrepeatableComponent(a)
                          for k, a in kwargs.items()}
        return FunctionComponent(_link_function(self.fct), self.output_group, *wrapped_args, **wrapped_kwargs)


def _link_function(fct: Callable) -> Callable:
    """Replace pytensor elementwise functions by their recording counterparts (by name)."""
    if isinstance(fct, sym.ElementwiseFunction):
        return fct
    name = getattr(fct, '__name__', None) or getattr(fct, 'name', None) or str(fct)
    known = {'exp': sym.exp, 'sigmoid': sym.sigmoid, 'invlogit': sym.sigmoid, 'expit': sym.sigmoid}
    if name in known:
        return known[name]
    return fct


# =============================================================================================
# fixed values
# =============================================================================================
class FixedValueComponent(ModelComponent):
    """Non-random value assigned to a group (reference ``fixed/base.py:12-39``)."""

    def __init__(self, value: Any, group: Union[str, Tuple[str, ...]], name: Optional[str] = None):
        super().__init__()
        self.group = _as_group_tuple(group)
        self.name = name
        self.values = value

    def set_name(self, name: str) -> None:
        self.name = name

    def clear(self) -> None:
        self.variable = None
        self.representation = None

    def prebuild(self, groupset: GroupSet) -> None:
        pass

    def retrieve_groups(self) -> Set[str]:
        return set(self.group)


class UnrepeatableComponent(FixedValueComponent):
    """A constant passed straight to the distribution (reference ``fixed/base.py:42-60``)."""

    def __init__(self, value: Any):
        super().__init__(value, 'global')

    def get_name(self) -> Optional[str]:
        return self.name if self.name is not None else 'fixed'

    def build_representation(self, groupset: GroupSet) -> None:
        self.representation = UnrepeatableRepresentation()

    def build_variable(self) -> None:
        self.variable = self.values

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        return self


class DataComponent(FixedValueComponent):
    """
    Data column wrapped as a component (reference ``fixed/data.py:16-44``).

    :param data: array with one element per member (combination) of ``group``; scalars are
        promoted to a length-1 array.
    """

    def __init__(self, data: Any, group: Union[str, Tuple[str, ...]], name: Optional[str] = None):
        values = data if isinstance(data, np.ndarray) else np.array([data])
        super().__init__(values, group, name)

    def get_name(self) -> Optional[str]:
        return self.name

    def build_representation(self, groupset: GroupSet) -> None:
        self.representation = MinimalTensorRepresentation()
        for g in self.group:
            self.representation.add_group(groupset[g])

    def build_variable(self) -> None:
        self.variable = sym.ConstantData(self.name, self.values)

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        return MinibatchComponent(self, batch_size, group)


def data_components(df: pd.DataFrame, group: Union[str, Tuple[str, ...]] = 'obs') -> Dict[str, DataComponent]:
    """One :class:`DataComponent` per column of ``df`` (reference ``fixed/data.py:47-58``)."""
    return {column: DataComponent(np.asarray(df[column].values), group, column) for column in df}


# =============================================================================================
# composables
# =============================================================================================
class Composable(ModelComponent):
    """Component built from named sub-components (reference ``composable/base.py:15-82``)."""

    def __init__(self, name: Optional[str], group: GroupSpec, subcomponents: Dict[Any, ModelComponent]):
        super().__init__()
        self.name = name
        self.group = _as_group_tuple(group)
        self.subcomponents = subcomponents
        self._groups_cache: Optional[Set[str]] = None

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        return MinibatchComponent(self, batch_size, group)

    def get_name(self) -> Optional[str]:
        return self.name

    def set_name(self, name: str) -> None:
        self.name = name

    def build_components(self, groupset: GroupSet) -> None:
        for param, comp in self.subcomponents.items():
            if comp.get_name() is None:
                comp.set_name(f'{param}_{self.get_name()}')
            if comp.variable is None:
                comp.build(groupset)

    def clear(self) -> None:
        self.variable = None
        self.representation = None
        for comp in self.subcomponents.values():
            comp.clear()

    def retrieve_groups(self) -> Set[str]:
        # cached per instance like the reference (@cache, quirk D.6)
        if self._groups_cache is None:
            groups: Set[str] = set(self.group)
            for comp in self.subcomponents.values():
                groups |= comp.retrieve_groups()
            self._groups_cache = groups
        return self._groups_cache

    def dims(self) -> Tuple[str, ...]:
        return tuple(str(g) for g in self.representation.groups)


class HierarchicalComponent(Composable):
    """
    Component defined on one or several groups whose parameters are other components
    (reference ``composable/hierarchical/base.py:12-55``).
    """

    def __getitem__(self, item: Any) -> ModelComponent:
        return self.subcomponents[item]

    def prebuild(self, groupset: GroupSet) -> None:
        self.build_components(groupset)

    def build_representation(self, groupset: GroupSet) -> None:
        rep = MinimalTensorRepresentation(groupset['global'])
        if len(self.group) == 0:
            for comp in self.subcomponents.values():
                for g in comp.representation.get_groups():
                    rep.add_group(g)
        else:
            for g in self.group:
                rep.add_group(groupset[g])
            if tuple(str(g) for g in rep.groups) != self.group:
                raise ValueError('Groups are not a minimal')
        self.representation = rep

    def get_built_components(self) -> Dict[str, Any]:
        """Sub-component variables gathered onto this component's shape (the prior-level parent
        gather)."""
        built = {}
        for key, comp in self.subcomponents.items():
            joint = MinimalTensorRepresentation(*self.representation.get_groups(),
                                                *comp.representation.get_groups())
            on_joint = comp.representation.map(comp.variable, joint)
            built[key] = joint.map(on_joint, self.representation)
        return built


class DistributionComponent(HierarchicalComponent):
    """
    Component whose variable is drawn from a distribution
    (reference ``composable/hierarchical/distribution.py:10-51``).

    :param generator: ``sakkara_b200.dist`` family or the PyMC distribution of the same name.
    :param name: variable name; derived from the parent (``'<param>_<parent>'``) when omitted.
    :param group: group(s) the variable is defined on; inferred from the parameters when omitted.
    :param subcomponents: distribution parameters -- components or plain values.
    """

    def __init__(self, generator: Callable, name: Optional[str] = None, group: GroupSpec = None,
                 **subcomponents: Any):
        wrapped = {k: v if isinstance(v, ModelComponent) else UnrepeatableComponent(v)
                   for k, v in subcomponents.items()}
        super().__init__(name, group, wrapped)
        self.generator = generator

    def build_variable(self) -> None:
        family = _dist.resolve_generator(self.generator)
        self.variable = family(self.name, **self.get_built_components(),
                               shape=self.representation.get_shape(), dims=self.dims())


class Likelihood(DistributionComponent):
    """
    Observed distribution (reference ``composable/hierarchical/likelihood.py:12-62``).

    Rows whose observation is NaN are masked like in the reference: their distribution parameters
    are replaced by the constants of ``nan_param_mask`` and their observation by ``nan_data_mask``,
    so they add a constant to logp and nothing to the gradient.  The reference composes this with a
    per-row ``GroupComponent`` (one Python object per row); here the parameters are wrapped in
    :class:`sakkara_b200.model.sym.Masked` nodes and the lowering drops the rows from the kernel plan
    and keeps their constant.
    """

    def __init__(self, generator: Callable, observed: DataComponent, name: str = 'likelihood',
                 group: Union[str, Tuple[str, ...]] = 'obs', nan_param_mask: Dict[str, Any] = None,
                 nan_data_mask: Any = None, **kwargs: Any):
        values = np.asarray(observed.values)
        self.nan_rows: Optional[np.ndarray] = None
        self.nan_param_mask: Optional[Dict[str, Any]] = None
        params = dict(kwargs)
        if values.dtype.kind == 'f' and np.isnan(values).any():
            # (same condition as the reference, likelihood.py:33: a mask for the parameters is required
            #  and the data mask has to be falsy -- in practice 0)
            if nan_param_mask is None or nan_data_mask:
                raise ValueError('Both nan_param_mask and nan_data_mask must be defined when there is '
                                 'Nan in the observed data.')
            warnings.warn('Observables contains NaN values, these will be ignored in likelihood computation.',
                          UserWarning)
            for k in kwargs:
                if k not in nan_param_mask:
                    raise ValueError('Mask values for all parameters must be defined in nan_param_mask')
            self.nan_rows = np.isnan(values)
            self.nan_param_mask = {k: nan_param_mask[k] for k in kwargs}
            fill = 0.0 if nan_data_mask is None else nan_data_mask
            params['observed'] = DataComponent(np.where(self.nan_rows, fill, values), group,
                                               name='masked_' + str(observed.name))
        else:
            params['observed'] = observed
        super().__init__(generator, name, group, **params)

    def build_variable(self) -> None:
        if self.nan_rows is None:
            return super().build_variable()
        family = _dist.resolve_generator(self.generator)
        built = self.get_built_components()
        shape = self.representation.get_shape()
        if tuple(shape) != self.nan_rows.shape:
            raise ValueError(f'NaN mask of shape {self.nan_rows.shape} for a likelihood of shape {tuple(shape)}')
        for key, fill in self.nan_param_mask.items():
            built[key] = sym.Masked(built[key], self.nan_rows, fill)
        self.variable = family(self.name, **built, shape=shape, dims=self.dims())


class MinibatchLikelihood(Likelihood):
    """
    Likelihood over mini-batches of iid rows (reference ``composable/hierarchical/likelihood.py:65-89``): every
    parameter and the observations are replaced by their mini-batch counterparts -- ``batch_size`` rows redrawn at
    every evaluation -- and the distribution is told the size of the whole data set (``total_size``), so that its
    logp is scaled by ``total_size / batch_size``.

    :param batch_size: observations per evaluation.
    """

    def __init__(self, generator: Callable, observed: DataComponent, batch_size: int, name: str = 'likelihood',
                 group: str = 'obs', nan_param_mask: Dict[str, Any] = None, nan_data_mask: Any = None, **kwargs: Any):
        kwargs['total_size'] = len(observed.values)
        super().__init__(generator, observed, name, group, nan_param_mask, nan_data_mask, **kwargs)
        if self.nan_rows is not None:
            raise NotImplementedError('NaN-masked observations in a MinibatchLikelihood')
        self.batch_size = batch_size
        for key, comp in self.subcomponents.items():
            self.subcomponents[key] = comp.to_minibatch(batch_size, group)


class Reshaper(HierarchicalComponent):
    """Force a component onto another group (reference ``composable/hierarchical/reshaper.py:9-21``)."""

    def __init__(self, component: ModelComponent, group: GroupSpec):
        super().__init__(None, group, subcomponents={'var': component})

    def build_variable(self) -> None:
        self.variable = self.get_built_components()['var']


class GroupComponent(Composable):
    """
    A component given member by member of a group (reference ``composable/group.py:15-103``): every member
    (value of the data-frame column; a tuple for several groups) gets its own component or plain value.

    The variable is the concatenation, in member order, of every member's component mapped to the joint
    representation and restricted to the member's cells -- a :class:`sakkara_b200.model.sym.Concat` node, which the
    lowering turns into gathers that read different parameter blocks (or constants) in different cells.

    :param group: group(s) the component is defined for.
    :param name: name of the resulting (deterministic) variable.
    :param membercomponents: ``{member: component or value}``.
    """

    def __init__(self, group: Union[str, Tuple[str, ...]], name: Optional[str] = None,
                 membercomponents: Optional[Dict[Any, Any]] = None):
        super().__init__(name, group, dict())
        self.base_representation = None
        self.components_representation = None
        if membercomponents is not None:
            for member, component in membercomponents.items():
                self.add(member, component)

    def __getitem__(self, item: Any) -> ModelComponent:
        return self.subcomponents[item if isinstance(item, tuple) else (item,)]

    def add(self, member: Any, component: Any) -> None:
        """Set the component (or plain value) of one member."""
        if not isinstance(component, ModelComponent):
            component = DataComponent(component, 'global')
        self.subcomponents[member if isinstance(member, tuple) else (member,)] = component
        self._groups_cache = None

    def prebuild(self, groupset: GroupSet) -> None:
        self.build_components(groupset)

    def build_representation(self, groupset: GroupSet) -> None:
        base = MinimalTensorRepresentation()
        for g in self.group:
            base.add_group(groupset[g])
        parts = MinimalTensorRepresentation(groupset['global'])
        for comp in self.subcomponents.values():
            parts = MinimalTensorRepresentation(*comp.representation.get_groups(), *parts.get_groups())
        joint = MinimalTensorRepresentation(*base.get_groups(), *parts.get_groups())
        # the member-wise pieces are laid out member-major, so the component's own groups must stay the leading ones
        if base.get_groups() != joint.get_groups()[:len(base.get_groups())]:
            raise ValueError('The group value specified for this GroupComponent have children among the '
                             'component groups')
        self.base_representation, self.components_representation, self.representation = base, parts, joint

    def build_variable(self) -> None:
        base, joint = self.base_representation, self.representation
        members = base.get_member_tuples()
        if not all(m in self.subcomponents for m in members):
            raise ValueError('All member of group component must be specified.')
        shape = joint.get_shape()
        # every base group's member per cell of the joint representation
        member_of_cell = [np.asarray(base.map(m, joint)).reshape(-1) for m in base.get_members()]
        pieces = []
        for member in members:
            comp = self.subcomponents[member]
            on_joint = comp.representation.map(comp.variable, joint)
            if not isinstance(on_joint, sym.Var):
                on_joint = np.broadcast_to(np.asarray(on_joint, dtype=np.float64), shape)
            mask = np.ones(int(np.prod(shape)), dtype=bool)
            for axis, value in enumerate(member):
                mask &= member_of_cell[axis] == value
            pieces.append(on_joint.reshape(-1)[mask])
        full = sym.stack(pieces).ravel().reshape(shape)
        self.variable = sym.Deterministic(self.name, full, dims=self.dims())


# =============================================================================================
# wrappers
# =============================================================================================
class WrapperComponent(ModelComponent):
    """Component wrapping exactly one other component (reference ``wrapper.py:10-35``)."""

    def __init__(self, component: ModelComponent):
        super().__init__()
        self.component = component

    def get_name(self) -> Optional[str]:
        return self.component.get_name()

    def set_name(self, name: str) -> None:
        self.component.set_name(name)

    def clear(self) -> None:
        self.component.clear()
        self.representation = None
        self.variable = None

    def retrieve_groups(self) -> Set[str]:
        return self.component.retrieve_groups()

    def prebuild(self, groupset: GroupSet) -> None:
        if self.component.variable is None:
            self.component.build(groupset)


class MinibatchComponent(WrapperComponent):
    """
    Mini-batch view of a component (reference ``minibatch.py:9-33``): the wrapped component is built in full, mapped
    to ``group`` and indexed with the group's mini-batch variable.
    """

    def __init__(self, component: ModelComponent, batch_size: int, group: str):
        super().__init__(component)
        self.batch_size = batch_size
        self.group = group
        self.transformed_variable = None

    def build_representation(self, groupset: GroupSet) -> None:
        self.representation = MinimalTensorRepresentation(groupset[self.group])
        self.transformed_variable = self.component.representation.map(self.component.variable, self.representation)

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        return self.component.to_minibatch(batch_size, group)

    def build_variable(self) -> None:
        minibatch = self.representation.get_groups()[0].get_minibatch(self.batch_size)
        value = self.transformed_variable
        if not isinstance(value, sym.Var):
            value = sym.DataVar(None, np.asarray(value))
        self.variable = value[minibatch]


class MinibatchDeterministic(WrapperComponent):
    """Named alias of a mini-batched variable, without dims (reference ``deterministic.py:23-32``)."""

    def __init__(self, name: str, component: ModelComponent):
        super().__init__(component)
        self.name = name

    def get_name(self) -> Optional[str]:
        return self.name

    def build_representation(self, groupset: GroupSet) -> None:
        self.representation = self.component.representation

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        return self

    def build_variable(self) -> None:
        self.variable = sym.Deterministic(self.name, self.component.variable)


class DeterministicComponent(WrapperComponent):
    """Named alias of a component's variable (reference ``deterministic.py:36-51``)."""

    def __init__(self, name: str, component: ModelComponent):
        super().__init__(component)
        self.name = name

    def build_representation(self, groupset: GroupSet) -> None:
        self.representation = self.component.representation

    def to_minibatch(self, batch_size: int, group: str) -> 'ModelComponent':
        return MinibatchDeterministic(self.name, self.component.to_minibatch(batch_size, group))

    def build_variable(self) -> None:
        self.variable = sym.Deterministic(self.name, self.component.variable,
                                          dims=tuple(str(g) for g in self.representation.groups))
