r"""
Conformance of the vectorised group algebra with the reference's relation semantics, on the
graph the reference itself tests (``/root/reference/tests/components/test_relation.py:9-42``):

      g (global)
      |\
      a  e
      |\ |
      b  d          'o' (observation) is a child of everything
      |
      c
"""
import numpy as np
import pandas as pd
import pytest

from sakkara_b200.relation import groupset
from sakkara_b200.relation.representation import MinimalTensorRepresentation as TR


@pytest.fixture
def df():
    e = np.array([0] * 5 + [1] * 11 + [0] * 2 + [1] * 14)
    return pd.DataFrame({
        'g': list(map(str, np.repeat('global', 32))),
        'a': list(map(str, np.repeat(np.arange(2), 16))),
        'b': list(map(str, np.repeat(np.arange(4), 8))),
        'c': list(map(str, np.repeat(np.arange(8), 4))),
        'd': list(map(str, e + 2 * np.repeat(np.arange(2), 16))),
        'e': e,
        'o': np.arange(32),
    })


@pytest.fixture
def gs(df):
    return groupset.init(df)


GRAPH = {
    'g': {'children': list('abcdeo'), 'parents': []},
    'a': {'children': list('bcdo'), 'parents': ['g']},
    'b': {'children': list('co'), 'parents': list('ga')},
    'c': {'children': ['o'], 'parents': list('gab')},
    'd': {'children': ['o'], 'parents': list('gae')},
    'e': {'children': list('od'), 'parents': ['g']},
    'o': {'children': [], 'parents': list('gabcde')},
}


def test_parent_matrix(df):
    df['o_copy'] = df['o'].copy()
    parents = groupset.get_parent_df(df)
    order = list(parents.index)
    assert order[0] == 'g' and set(order[1:3]) == set('ae') and set(order[3:6]) == set('bdc')
    for name in df.columns:
        assert not parents.loc[name, name]
    for name in 'abcdeo':
        assert parents.loc[name, 'g']
        assert not parents.loc['g', name]
    assert parents.loc['b', 'a'] and not parents.loc['b', list('cdeo')].any()
    assert parents.loc['c', list('ab')].all() and not parents.loc['c', list('de')].any()
    assert parents.loc['d', list('ae')].all() and not parents.loc['d', list('bc')].any()
    assert not parents.loc['e', list('abcd')].any()
    # equal cardinality: twins, not parents (strict ">" of groupset.py:44)
    assert not parents.loc['o_copy', 'o'] and not parents.loc['o', 'o_copy']
    twins = groupset.get_twin_df(df)
    assert twins.loc['o', 'o_copy'] and twins.loc['o_copy', 'o'] and not twins.loc['a', 'e']


def test_sizes_members_and_coords(gs, df):
    assert {k: len(v) for k, v in gs.groups.items()} == {'g': 1, 'a': 2, 'b': 4, 'c': 8, 'd': 4, 'e': 2, 'o': 32}
    coords = gs.coords()
    for name in df.columns:
        assert list(coords[name]) == list(pd.unique(df[name]))      # first-appearance order


def test_first_appearance_order():
    frame = pd.DataFrame({'k': [5, 3, 5, 9, 3, 9, 1]})
    g = groupset.init(frame)['k']
    assert list(g.members) == [5, 3, 9, 1]
    assert list(g.codes) == [0, 1, 0, 2, 1, 2, 3]
    assert g.codes.dtype == np.int64


def test_relations(gs):
    for name, rel in GRAPH.items():
        assert {str(x) for x in gs[name].children} == set(rel['children'])
        assert {str(x) for x in gs[name].parents} == set(rel['parents'])
        assert gs[name].twins == {gs[name]}


def test_parent_codes(gs, df):
    # code of the parent member of each child member, taken at the child's first appearance
    assert list(gs['c'].maps['b']) == [0, 0, 1, 1, 2, 2, 3, 3]
    assert list(gs['c'].maps['a']) == [0, 0, 0, 0, 1, 1, 1, 1]
    assert list(gs['d'].maps['e']) == [0, 1, 0, 1]
    assert list(gs['d'].maps['a']) == [0, 0, 1, 1]
    table = gs['c'].mapping
    assert list(table.index) == list(gs['c'].members) and set(table.columns) == {'c', 'b', 'a', 'g'}


def test_minimal_representation(gs):
    def names(*groups):
        return [str(g) for g in TR(*[gs[k] for k in groups]).get_groups()]
    assert names('g', 'a') == ['a']
    assert names('a', 'g') == ['a']
    assert names('a', 'b', 'c') == ['c']
    assert names('c', 'a') == ['c']
    assert names('b', 'e') == ['b', 'e']
    assert names('b', 'd') == ['b', 'd']
    assert names('a', 'e', 'd') == ['d']
    assert names('g', 'o', 'c') == ['o']


def test_map_shapes_identity_and_errors(gs):
    class Recorder:
        def __getitem__(self, index):
            return index

    rec = Recorder()
    # identity: equal representations return the element untouched
    assert TR(gs['b']).map(rec, TR(gs['b'])) is rec
    # parent -> child
    (idx,) = TR(gs['a']).map(rec, TR(gs['c']))
    assert idx.dtype == np.int64 and list(idx) == [0, 0, 0, 0, 1, 1, 1, 1]
    # global -> anything is a gather with zeros, not a broadcast
    (idx,) = TR(gs['g']).map(rec, TR(gs['o']))
    assert idx.shape == (32,) and not idx.any()
    # block -> observations == factorize codes
    (idx,) = TR(gs['d']).map(rec, TR(gs['o']))
    assert np.array_equal(idx, gs['d'].codes)
    # onto a two-group target
    (idx,) = TR(gs['a']).map(rec, TR(gs['b'], gs['e']))
    assert idx.shape == (4, 2) and np.array_equal(idx, np.repeat([[0], [0], [1], [1]], 2, axis=1))
    # child -> tuple of parents: d (4 members) is identified by (a, e)
    (idx,) = TR(gs['d']).map(rec, TR(gs['a'], gs['e']))
    assert idx.shape == (2, 2) and np.array_equal(idx, [[0, 1], [2, 3]])
    # not mappable: child information cannot be recovered from a parent
    with pytest.raises(ValueError):
        TR(gs['c']).map(rec, TR(gs['a']))
    with pytest.raises(ValueError):
        TR(gs['b']).map(rec, TR(gs['e']))


def test_member_arrays(gs):
    rep = TR(gs['b'], gs['e'])
    assert rep.get_shape() == (4, 2)
    mb, me = rep.get_members()
    assert mb.shape == (4, 2) and list(mb[:, 0]) == list(gs['b'].members)
    assert list(me[0]) == list(gs['e'].members)


def test_nestedness_is_inferred_from_the_sample():
    # two crossed factors become parent/child when, in the sample, one determines the other
    # 
    frame = pd.DataFrame({'A': [0, 0, 1, 1, 2, 2], 'B': [0, 0, 0, 0, 1, 1]})
    gs = groupset.init(frame)
    assert gs['B'] in gs['A'].parents and gs['A'] in gs['B'].children


def test_missing_values_rejected():
    with pytest.raises(ValueError):
        groupset.init(pd.DataFrame({'k': [1.0, np.nan, 2.0]}))
