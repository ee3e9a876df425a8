"""Flatten a ModelPlan to a dict of arrays (for .npz golden fixtures) and compare two plans."""
import json

import numpy as np


def plan_to_arrays(plan) -> dict:
    out = {}
    meta = {'n_rows': plan.n_rows, 'blocks': [], 'terms': [], 'likelihood': {}}
    for i, b in enumerate(plan.blocks):
        params = {}
        for key, prm in b.params.items():
            if prm.is_const:
                out[f'b{i}_{key}_const'] = np.asarray(prm.const, dtype=np.float64)
                params[key] = 'const'
            elif prm.is_mixed:
                out[f'b{i}_{key}_const'] = np.asarray(prm.const, dtype=np.float64)
                out[f'b{i}_{key}_index'] = np.asarray(prm.index, dtype=np.int64)
                params[key] = 'member-wise'
            else:
                out[f'b{i}_{key}_index'] = np.asarray(prm.index, dtype=np.int64)
                params[key] = int(prm.block)
        meta['blocks'].append({'name': b.name, 'family': b.family, 'offset': b.offset, 'size': b.size,
                               'shape': list(b.shape), 'dims': list(b.dims) if b.dims is not None else None,
                               'transform': b.transform, 'params': params})
    for i, t in enumerate(plan.terms):
        out[f't{i}_index'] = np.asarray(t.index, dtype=np.int64)
        if t.x is not None:
            out[f't{i}_x'] = np.asarray(t.x, dtype=np.float64)
        meta['terms'].append({'block': t.block, 'has_x': t.x is not None})
    lik = plan.likelihood
    out['y'] = np.asarray(lik.y, dtype=np.float64)
    if lik.offset is not None:
        out['offset'] = np.asarray(lik.offset, dtype=np.float64)
    meta['likelihood'] = {'family': lik.family, 'name': lik.name, 'sigma_const': lik.sigma_const,
                          'sigma_block': lik.sigma_block, 'sigma_index': lik.sigma_index}
    if getattr(lik, 'sigma_rows_element', None) is not None:
        out['sigma_rows_element'] = np.asarray(lik.sigma_rows_element, dtype=np.int64)
        out['sigma_rows_const'] = np.asarray(lik.sigma_rows_const, dtype=np.float64)
    out['meta'] = np.array(json.dumps(meta))
    return out


def assert_same(arrays_a: dict, arrays_b: dict) -> None:
    """Bit-exact comparison of two flattened plans (integer arrays, float data and structure)."""
    meta_a, meta_b = json.loads(str(arrays_a['meta'])), json.loads(str(arrays_b['meta']))
    assert meta_a == meta_b, f'structure differs:\n{meta_a}\n{meta_b}'
    keys_a = sorted(k for k in arrays_a if k != 'meta')
    keys_b = sorted(k for k in arrays_b if k != 'meta')
    assert keys_a == keys_b
    for k in keys_a:
        a, b = np.asarray(arrays_a[k]), np.asarray(arrays_b[k])
        assert a.dtype == b.dtype and a.shape == b.shape, k
        assert np.array_equal(a, b), f'{k} differs'
