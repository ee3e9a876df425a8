"""
Construction of the group graph from a data frame.

Semantics follow the reference (``sakkara/relation/groupset.py:27-93``) but every step is an
O(N) vectorised pass over integer codes instead of ``groupby().nunique()`` and per-member
``DataFrame.loc`` look-ups, so the same relations can be inferred at 1e8 rows:

* members      = unique column values in first-appearance order  (``groupset.py:76``)
* j parent of i iff every member of i co-occurs with exactly one member of j **and** i has
  strictly more members than j                                    (``groupset.py:42-44``)
* twins        = mutual one-to-one co-occurrence                  (``groupset.py:52-65``)
* parent code of a child member = code of the parent value on the row where the child member
  appears first                                                   (``groupset.py:78-86``)
"""
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import pandas as pd

from sakkara_b200.relation.group import Group


@dataclass(frozen=True)
class GroupSet:
    """Set of group nodes used for model creation (reference ``groupset.py:10-24``)."""
    groups: Dict[str, Group]

    def __getitem__(self, item: str) -> Group:
        return self.groups[item]

    def __contains__(self, item: str) -> bool:
        return item in self.groups

    def coords(self) -> Dict[str, np.ndarray]:
        return {name: group.members for name, group in self.groups.items()}


def first_rows(codes: np.ndarray, n_members: int) -> np.ndarray:
    """Row of the first appearance of every code: one reversed scatter (the last write of a slot wins)."""
    first_row = np.empty(n_members, dtype=np.int64)
    first_row[codes[::-1]] = np.arange(len(codes) - 1, -1, -1, dtype=np.int64)
    return first_row


def factorize_column(values) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """
    First-appearance factorisation of one column.

    :return: ``(codes[N] int64, members[G], first_row[G] int64)`` where ``first_row[g]`` is the
        row on which member ``g`` appears for the first time.
    """
    n = len(values)
    if n and isinstance(values, np.ndarray) and values.dtype.kind in 'iu' and values[0] == 0 and values[-1] == n - 1:
        # the ``obs`` column build() adds (and any other row counter): its own factorisation, no hashing
        counter = np.arange(n, dtype=np.int64)
        if np.array_equal(values, counter):
            return counter, values, counter
    codes, members = pd.factorize(values, sort=False)
    if (codes < 0).any():
        raise ValueError('Group columns must not contain missing values')
    codes = codes.astype(np.int64, copy=False)
    members = np.asarray(members) if not isinstance(members, np.ndarray) else members
    return codes, members, first_rows(codes, len(members))


def _determines(codes_i: np.ndarray, first_row_i: np.ndarray, codes_j: np.ndarray) -> bool:
    """True iff each member of column i is seen together with exactly one member of column j."""
    representative = codes_j[first_row_i]
    return bool(np.array_equal(representative[codes_i], codes_j))


def functional_matrix(columns: List[str], codes: Dict[str, np.ndarray],
                      first_row: Dict[str, np.ndarray]) -> np.ndarray:
    """
    ``F[i, j]`` = column i determines column j (diagonal False).  Decided from the member counts where they
    suffice -- a function from the members of i onto those of j needs at least as many members in i; a column with one
    member per row determines every column; every column determines a constant one -- and by one pass over the rows
    otherwise.
    """
    n = len(columns)
    n_rows = len(codes[columns[0]]) if columns else 0
    sizes = [len(first_row[c]) for c in columns]
    out = np.zeros((n, n), dtype=bool)
    for i, ci in enumerate(columns):
        for j, cj in enumerate(columns):
            if i == j:
                continue
            if sizes[i] < sizes[j]:
                out[i, j] = False
            elif sizes[i] == n_rows or sizes[j] <= 1:
                out[i, j] = True
            else:
                out[i, j] = _determines(codes[ci], first_row[ci], codes[cj])
    return out


def get_parent_df(df: pd.DataFrame) -> pd.DataFrame:
    """
    CxC boolean frame, element (i, j) True when column j is a parent of column i; rows ordered
    from the group with fewest parents to the one with most (reference ``groupset.py:27-49``).
    """
    columns = list(df.columns)
    fact = {c: factorize_column(df[c].values) for c in columns}
    codes = {c: f[0] for c, f in fact.items()}
    first_row = {c: f[2] for c, f in fact.items()}
    sizes = np.array([len(fact[c][1]) for c in columns])
    determines = functional_matrix(columns, codes, first_row)
    is_parent = determines & (sizes[:, None] > sizes[None, :])
    order = np.argsort(is_parent.sum(axis=1), kind='stable')
    return pd.DataFrame(is_parent[order], index=[columns[k] for k in order], columns=columns)


def get_twin_df(df: pd.DataFrame) -> pd.DataFrame:
    """CxC boolean frame of twin relations (reference ``groupset.py:52-65``)."""
    columns = list(df.columns)
    fact = {c: factorize_column(df[c].values) for c in columns}
    determines = functional_matrix(columns, {c: f[0] for c, f in fact.items()},
                                   {c: f[2] for c, f in fact.items()})
    return pd.DataFrame(determines & determines.T, index=columns, columns=columns)


def init(df: pd.DataFrame) -> GroupSet:
    """
    Build the :class:`GroupSet` of a frame that holds only group columns
    (reference ``groupset.py:68-93``).
    """
    columns = list(df.columns)
    fact = {c: factorize_column(df[c].values) for c in columns}
    codes = {c: f[0] for c, f in fact.items()}
    first_row = {c: f[2] for c, f in fact.items()}
    groups = {c: Group(c, fact[c][1], codes[c]) for c in columns}

    sizes = np.array([len(groups[c]) for c in columns])
    determines = functional_matrix(columns, codes, first_row)
    is_parent = determines & (sizes[:, None] > sizes[None, :])
    is_twin = determines & determines.T

    for i, child in enumerate(columns):
        for j, parent in enumerate(columns):
            if is_parent[i, j]:
                groups[child].add_parent(groups[parent], codes[parent][first_row[child]])
                groups[parent].add_child(groups[child])
    for i, name in enumerate(columns):
        for j, other in enumerate(columns):
            if is_twin[i, j]:
                groups[name].add_twin(groups[other])

    return GroupSet(groups)
