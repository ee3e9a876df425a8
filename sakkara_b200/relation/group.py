"""
Group nodes of the relation graph (integer half of the hot path).

A :class:`Group` is one DataFrame column: its *members* are the unique values of the column in
first-appearance order and every row of the frame carries the integer *code* of its member.
Relations to other groups (parent / child / twin) are stored as plain ``int64`` lookup vectors
``maps[other] : member code -> code in other`` instead of the pandas tables the reference keeps
(reference: ``sakkara/relation/group.py:17-52``), so that every gather index the model needs can
be produced with vectorised NumPy at 1e8 rows.  The pandas view (``Group.mapping``) is still
available for callers that used it.
"""
from typing import Any, Dict, Optional, Set

import numpy as np
import pandas as pd


class Group:
    """
    One column of the data frame seen as a node of the group graph.

    :param name: name of the group, equal to the column name.
    :param members: unique values of the column ordered by first appearance
        (reference semantics: ``sakkara/relation/groupset.py:76``).
    :param codes: optional per-row member code (``int64``, same order as the frame's rows).
    """

    def __init__(self, name: str, members, codes: Optional[np.ndarray] = None):
        self.name = name
        self.members = members
        self.codes = codes
        self.parents: Set['Group'] = set()
        self.children: Set['Group'] = set()
        self.twins: Set['Group'] = {self}
        # member code -> code of the related member in group <key>; identity for the group itself
        self.maps: Dict[str, np.ndarray] = {name: np.arange(len(members), dtype=np.int64)}
        self.minibatch = None

    # -- relation wiring -------------------------------------------------------------------
    def add_child(self, child: 'Group') -> None:
        self.children.add(child)

    def add_parent(self, parent: 'Group', parent_mapping) -> None:
        """
        :param parent_mapping: for every member of this group (in member order) the code of its
            parent member (reference: ``sakkara/relation/group.py:34-43``).
        """
        self.parents.add(parent)
        self.maps[parent.name] = np.asarray(parent_mapping, dtype=np.int64)

    def add_twin(self, twin: 'Group') -> None:
        # twins enumerate their members in the same first-appearance order, hence identity
        self.twins.add(twin)
        self.maps[twin.name] = np.arange(len(self), dtype=np.int64)

    # -- views -----------------------------------------------------------------------------
    @property
    def mapping(self) -> pd.DataFrame:
        """pandas view of the lookup vectors, indexed by member (what the reference stores)."""
        return pd.DataFrame(index=self.members, data=self.maps)

    def code_of(self, other: 'Group') -> np.ndarray:
        """Lookup vector member-code -> code in ``other`` (a parent, a twin or the group itself)."""
        try:
            return self.maps[other.name]
        except KeyError:
            raise ValueError(f'Group {other.name} is not a parent or twin of {self.name}')

    def get_minibatch(self, batch_size):
        """The group's minibatch index variable, created at the first call (reference ``group.py:54-60``)."""
        if getattr(self, 'minibatch', None) is None:
            from sakkara_b200.model.sym import MinibatchIndex
            self.minibatch = MinibatchIndex(len(self), batch_size)
        return self.minibatch

    def clear_minibatch(self) -> None:
        self.minibatch = None

    def __str__(self) -> str:
        return self.name

    def __repr__(self) -> str:
        return f'Group({self.name!r}, n={len(self)})'

    def __len__(self) -> int:
        return len(self.members)

    def __lt__(self, other: Any) -> bool:
        return str(self) < str(other)
