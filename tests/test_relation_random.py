"""
Property test of the integer half against the reference itself: on random frames with nested,
crossed, twin and constant columns (integer, string and float labels, shuffled rows) the vectorised
group algebra must reproduce the reference's group graph and every gather index of
``Representation.map`` bit for bit (reference: /root/reference/sakkara/relation/*.py, run with the
stub modules of oracle/ref_harness.py).  Skipped where the reference sources are absent.
"""
import itertools

import numpy as np
import pandas as pd
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import ref_harness
from sakkara_b200.relation import MinimalTensorRepresentation as TR
from sakkara_b200.relation import init

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason='reference sources not present')


class Recorder:
    def __getitem__(self, index):
        return index


@st.composite
def frames(draw):
    n = draw(st.integers(4, 40))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    cols = {}
    base = rng.integers(0, draw(st.integers(2, 6)), n)                 # a free factor
    cols['a'] = base
    fine = base * 3 + rng.integers(0, 3, n)                            # nested inside a
    cols['b'] = np.array([f'm{v}' for v in fine], dtype=object)
    cols['c'] = rng.integers(0, draw(st.integers(1, 4)), n)            # crossed with a
    if draw(st.booleans()):
        relabel = rng.permutation(fine.max() + 1).astype(float) / 2    # twin of b with float labels
        cols['d'] = relabel[fine]
    if draw(st.booleans()):
        cols['e'] = np.repeat('all', n).astype(object)                 # constant column
    frame = pd.DataFrame(cols).iloc[rng.permutation(n)].reset_index(drop=True)
    frame['obs'] = np.arange(n)
    return frame


def reference_groupset(frame):
    with ref_harness.reference_sakkara():
        from sakkara.relation import groupset as ref_groupset
        from sakkara.relation.representation import MinimalTensorRepresentation as RefTR
        return ref_groupset.init(ref_harness._object_strings(frame)), RefTR


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(frames())
def test_group_graph_and_gather_indices_match_the_reference(frame):
    ref, RefTR = reference_groupset(frame)
    mine = init(frame)
    names = list(frame.columns)
    for k in names:
        assert list(ref[k].members) == list(mine[k].members)
        assert {str(g) for g in ref[k].parents} == {str(g) for g in mine[k].parents}
        assert {str(g) for g in ref[k].children} == {str(g) for g in mine[k].children}
        assert {str(g) for g in ref[k].twins} == {str(g) for g in mine[k].twins}
        for col in ref[k].mapping.columns:
            assert np.array_equal(ref[k].mapping[col].values, mine[k].maps[col])
    rec = Recorder()
    singles = list(itertools.combinations(names, 1))
    pairs = list(itertools.combinations(names, 2))[:6]
    for src in singles + pairs[:3]:
        for tgt in singles + pairs:
            rs, rt = RefTR(*[ref[s] for s in src]), RefTR(*[ref[s] for s in tgt])
            ms, mt = TR(*[mine[s] for s in src]), TR(*[mine[s] for s in tgt])
            assert [str(g) for g in rs.groups] == [str(g) for g in ms.groups]
            try:
                expected, failed = rs.map(rec, rt), None
            except Exception as err:            # unmappable: both must refuse
                expected, failed = None, err
            if failed is not None:
                with pytest.raises(Exception):
                    ms.map(rec, mt)
                continue
            got = ms.map(rec, mt)
            if expected is rec:
                assert got is rec
                continue
            assert len(expected) == len(got)
            for a, b in zip(expected, got):
                assert b.dtype == np.int64 and a.shape == b.shape and np.array_equal(a, b)
