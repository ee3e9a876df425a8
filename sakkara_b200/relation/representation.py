"""
Shape algebra of components.

A representation is the ordered list of groups that spans the tensor of a component.  Mapping a
tensor from one representation to another is a gather with integer index arrays; those index
arrays are the only thing the relation layer contributes to the logp+dlogp hot path, and they
must be bit-identical to what the reference computes (``sakkara/relation/representation.py:222-253``).

The reference finds every index with a Python loop of ``DataFrame.loc`` calls over the target
members (``representation.py:249``, ~16-21 us per row).  Here the same indices
come from integer code arrays and the groups' lookup vectors, O(size of the target) vectorised:

* source group is a twin/parent of a target group ``t``  ->  ``t.maps[source][codes of t in target]``
* source group is a child of several target groups whose sizes multiply to its own size
  ->  dense look-up table over the parent-code tuples (``representation.py:75-123``)
"""
import abc
from abc import ABC
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from sakkara_b200.relation.group import Group


class Representation:
    """Abstract shape of a component's variable (reference ``representation.py:11-63``)."""

    @abc.abstractmethod
    def get_shape(self) -> Tuple[int, ...]:
        raise NotImplementedError

    @abc.abstractmethod
    def get_groups(self) -> List[Group]:
        raise NotImplementedError

    @abc.abstractmethod
    def get_member_array(self, group: Group) -> np.ndarray:
        raise NotImplementedError

    @abc.abstractmethod
    def get_members(self) -> Tuple[np.ndarray, ...]:
        raise NotImplementedError

    @abc.abstractmethod
    def get_member_tuples(self) -> List[Tuple[Any, ...]]:
        raise NotImplementedError

    @abc.abstractmethod
    def map(self, element: Any, target_representation: 'Representation') -> Any:
        raise NotImplementedError


def one_to_one_mapping(group: Group, target: Representation) -> Optional[Group]:
    """First target group that has ``group`` as a twin or parent (reference ``:66-72``)."""
    for target_group in target.get_groups():
        if group in target_group.twins or group in target_group.parents:
            return target_group
    return None


def subset_prod(remain: int, candidates: List[Group]) -> Tuple[bool, List[Group]]:
    """
    Choose candidates whose sizes multiply to ``remain``; same search order and result order as
    the reference (``representation.py:75-102``): greedy on the first divisor, back-tracking
    over the tail, result listed innermost choice first.
    """
    pick = None
    for position, cand in enumerate(candidates):
        if len(cand) <= remain and remain % len(cand) == 0:
            pick = position
            break
    if pick is None:
        return False, []
    chosen = candidates[pick]
    rest = candidates[pick + 1:]
    if len(chosen) == remain:
        return True, [chosen]
    ok, found = subset_prod(remain // len(chosen), rest)
    if ok:
        found.append(chosen)
        return True, found
    return subset_prod(remain, rest)


def one_to_combination_mapping(group: Group, target: Representation) -> Tuple[Group, ...]:
    """Parent groups in ``target`` that jointly identify the members of ``group`` (reference ``:105-123``)."""
    candidates = [t for t in target.get_groups() if group in t.children and len(group) % len(t) == 0]
    size = int(np.prod([len(t) for t in candidates])) if candidates else 1
    if len(group) % size != 0:
        raise ValueError('Group is not mappable to target representation')
    _, mapped = subset_prod(len(group), candidates)
    return tuple(mapped)


class UnrepeatableRepresentation(Representation, ABC):
    """Representation of plain constants: one value, never repeated (reference ``:126-150``)."""

    def get_shape(self) -> Tuple[int, ...]:
        return 1,

    def get_groups(self) -> List[Group]:
        return []

    def get_member_array(self, group: Group):
        raise ValueError('This representation does not hold any groups')

    def get_members(self):
        raise ValueError('This representation does not hold any groups')

    def get_member_tuples(self):
        raise ValueError('This representation does not hold any groups')

    def map(self, element: object, target_representation: 'Representation'):
        return element

    def __eq__(self, other):
        return isinstance(other, UnrepeatableRepresentation)

    __hash__ = None


class TensorRepresentation(Representation, ABC):
    """Representation spanned by a sequence of groups (reference ``:153-266``)."""

    def __init__(self, *groups: Group):
        self.groups: List[Group] = []
        for g in groups:
            self.add_group(g)

    @abc.abstractmethod
    def add_group(self, group: Group):
        raise NotImplementedError

    def get_groups(self) -> List[Group]:
        return self.groups

    def get_shape(self) -> Tuple[int, ...]:
        return tuple(len(g) for g in self.groups)

    def __eq__(self, other: 'Representation'):
        theirs = other.get_groups()
        return len(self.groups) == len(theirs) and all(a in b.twins for a, b in zip(self.groups, theirs))

    __hash__ = None

    # -- code/member arrays shaped like this representation ----------------------------------
    def get_code_array(self, group: Group) -> np.ndarray:
        """int64 array shaped like the representation holding the code of ``group`` per cell."""
        if group not in self.groups:
            raise ValueError('Group is not found in this representation')
        axis = self.groups.index(group)
        shape = [1] * len(self.groups)
        shape[axis] = len(group)
        return np.broadcast_to(np.arange(len(group), dtype=np.int64).reshape(shape), self.get_shape())

    def get_member_array(self, group: Group) -> np.ndarray:
        members = group.members if isinstance(group.members, np.ndarray) else np.asarray(group.members)
        return members[self.get_code_array(group)]

    def get_members(self) -> Tuple[np.ndarray, ...]:
        if len(self.groups) == 0:
            raise ValueError('This representation does not hold any groups')
        return tuple(self.get_member_array(g) for g in self.groups)

    def get_member_tuples(self) -> List[Tuple[Any, ...]]:
        if len(self.groups) == 0:
            raise ValueError('This representation does not hold any groups')
        return list(zip(*[np.ravel(m) for m in self.get_members()]))

    # -- mapping -----------------------------------------------------------------------------
    def get_group_mapping(self, target: 'Representation') -> Dict[Group, Tuple[Group, ...]]:
        """For each own group the target group(s) that determine it (reference ``:175-197``)."""
        out = {}
        for group in self.groups:
            single = one_to_one_mapping(group, target)
            if single is not None:
                out[group] = (single,)
                continue
            combo = one_to_combination_mapping(group, target)
            if len(combo) == 0:
                raise ValueError('Group could not be mapped to any other group or combination of groups')
            out[group] = combo
        return out

    def index_arrays(self, target: 'Representation') -> Tuple[np.ndarray, ...]:
        """
        The gather indices of :meth:`map`: one int64 array per own group, each shaped like
        ``target`` (bit-exact with reference ``representation.py:229-253``).
        """
        assignment = self.get_group_mapping(target)
        target_shape = target.get_shape()
        arrays = []
        for group in self.groups:
            via = assignment[group]
            if len(via) == 1:
                carrier = via[0]
                idx = carrier.code_of(group)[target.get_code_array(carrier)]
            else:
                # members of `group` are identified by the tuple of their parents' codes
                table = np.full(tuple(len(p) for p in via), -1, dtype=np.int64)
                table[tuple(group.code_of(p) for p in via)] = np.arange(len(group), dtype=np.int64)
                idx = table[tuple(target.get_code_array(p) for p in via)]
                if (idx < 0).any():
                    raise KeyError(f'Some member combinations of {[str(p) for p in via]} have no member in {group}')
            arrays.append(np.ascontiguousarray(idx, dtype=np.int64).reshape(target_shape))
        return tuple(arrays)

    def map(self, element: Any, target: Representation) -> Any:
        if self == target:
            return element
        return element[self.index_arrays(target)]


class MinimalTensorRepresentation(TensorRepresentation, ABC):
    """
    Keeps the smallest set of groups that still determines every group added so far: a group is
    skipped when a twin or child is already present, and adding a child evicts its parents
    (reference ``representation.py:269-290``).
    """

    def add_group(self, group: Group) -> None:
        if any(present in group.twins or present in group.children for present in self.groups):
            return
        self.groups.append(group)
        for parent in group.parents:
            if parent in self.groups:
                self.groups.remove(parent)
