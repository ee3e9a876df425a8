"""
TEST INFRASTRUCTURE -- not part of the product path.

Runs the *reference's own* ``sakkara.model.build()`` (from ``/root/reference``, read-only) against
a fake ``pymc`` / ``pytensor`` whose distributions and data containers are the recording classes of
``sakkara_b200.model.sym`` (the "recording fake PyMC" harness of SURVEY.md §4 / probe C.3).  What
comes back is the graph the reference would have handed to PyMC -- free-variable creation order,
shapes, dims and every gather index computed by the reference's ``Representation.map`` -- which is
then lowered with the same ``lower_model`` and compared with what ``sakkara_b200.model.build``
produces.  PyMC and PyTensor are not installable in this image, so this is the
strongest executable pin of the integer half of the hot path.

Only ``tests/`` and ``tests/golden/make_golden.py`` import this module, and only where
``/root/reference`` exists (it does not on the GPU box).
"""
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('SAKKARA_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'sakkara'))


def _fake_modules():
    from sakkara_b200.model import dist, sym

    pm = types.ModuleType('pymc')
    pm.Model = sym.Model
    pm.ConstantData = sym.ConstantData
    pm.Deterministic = sym.Deterministic
    for name, family in dist.FAMILIES.items():
        setattr(pm, name, family)
    pm_data = types.ModuleType('pymc.data')
    pm_data.minibatch_index = lambda *a, **k: None      # only used by Group.get_minibatch (ADVI)
    pm.data = pm_data

    pt = types.ModuleType('pytensor')
    ptt = types.ModuleType('pytensor.tensor')
    ptt.TensorVariable = sym.Var
    ptt.Variable = sym.Var
    ptt.exp = sym.exp
    ptt.sigmoid = sym.sigmoid
    ptt.stack = sym.stack                               # composable/group.py:100
    pt.tensor = ptt
    return {'pymc': pm, 'pymc.data': pm_data, 'pytensor': pt, 'pytensor.tensor': ptt}


@contextlib.contextmanager
def reference_sakkara():
    """
    Context in which ``import sakkara`` resolves to the reference package bound to the fake
    PyMC.  Yields ``(sakkara.model module, fake pymc module, fake pytensor.tensor module)``.
    """
    if not reference_available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    import pandas as pd
    fakes = _fake_modules()
    saved = {k: sys.modules.get(k) for k in fakes}
    saved_sakkara = {k: v for k, v in sys.modules.items() if k == 'sakkara' or k.startswith('sakkara.')}
    for k in saved_sakkara:
        del sys.modules[k]
    sys.modules.update(fakes)
    sys.path.insert(0, REFERENCE_ROOT)
    # pandas 3 returns Arrow-backed strings from .unique(); the reference (pandas 1.5) expects
    # object arrays (SURVEY.md probe C.1)
    infer = pd.get_option('future.infer_string')
    pd.set_option('future.infer_string', False)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', SyntaxWarning)
            import sakkara.model as ref_model
        yield ref_model, fakes['pymc'], fakes['pytensor.tensor']
    finally:
        pd.set_option('future.infer_string', infer)
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == 'sakkara' or k.startswith('sakkara.')]:
            del sys.modules[k]
        sys.modules.update(saved_sakkara)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def reference_plan(make_likelihood, df):
    """
    Build ``make_likelihood(api, pm, pt)`` with the reference and lower the recorded graph.

    :param make_likelihood: callable ``(api, pm, pt) -> Likelihood`` written against the public DSL
        (``api.DistributionComponent`` ...), so the same function can be run on both
        implementations.
    """
    from sakkara_b200.plan.lower import lower_model
    with reference_sakkara() as (api, pm, pt):
        recorded = api.build(_object_strings(df), make_likelihood(api, pm, pt))
    return lower_model(recorded)


def _object_strings(df):
    """pandas-3 string columns -> object dtype, which is what the reference (pandas 1.5) sees."""
    import pandas as pd
    out = df.copy()
    for column in out.columns:
        if pd.api.types.is_string_dtype(out[column].dtype) and out[column].dtype != object:
            out[column] = out[column].astype(object)
    return out
