"""
TEST INFRASTRUCTURE -- the real reference as the oracle, when it can run ("if a real
PyMC ever appears, the harness must prefer ``model.logp_dlogp_function()`` as the oracle and report
which oracle was used").

PyMC 5 / PyTensor are not installable in the build image (no wheel, no network), so nothing in this
module has been executed there: every entry point is guarded by :func:`available`, the tests that use
it are skipped without PyMC, and ``tests/conftest.py`` prints in the pytest header which oracle the
floating-point half is pinned by ("parity unpinned" for the NumPy restatement, "PyMC x.y" otherwise).

With PyMC importable the flow is the reference's own: ``sakkara.model.build(df, likelihood)`` (the
package under ``/root/reference`` or an installed ``sakkara``) returns a ``pm.Model``;
``model.logp_dlogp_function()`` is the compiled value-and-gradient function that NUTS calls per
leapfrog step (pymc 5.3.1 ``pymc/model.py``: ``ValueGradFunction``), evaluated on the raveled
unconstrained point in the order of ``model.value_vars`` -- the same order as ``ModelPlan.blocks``
(``tests/test_build_parity.py`` pins that order against the reference without PyMC).
"""
import contextlib
import os
import sys
from typing import Callable, Tuple

import numpy as np

REFERENCE_ROOT = os.environ.get('SAKKARA_REFERENCE', '/root/reference')


def available() -> bool:
    """True when PyMC >= 5 and a ``sakkara`` package (installed, or the reference checkout) can be imported."""
    try:
        import pymc  # noqa: F401
        import pytensor  # noqa: F401
    except Exception:
        return False
    try:
        with _reference_on_path():
            import sakkara.model  # noqa: F401
        return True
    except Exception:
        return False


def describe() -> str:
    if not available():
        return 'NumPy/C restatement of the PyTensor graph (floating-point parity unpinned: PyMC is not importable)'
    import pymc
    import pytensor
    return f'real reference: sakkara on PyMC {pymc.__version__} / PyTensor {pytensor.__version__}'


@contextlib.contextmanager
def _reference_on_path():
    added = False
    if os.path.isdir(os.path.join(REFERENCE_ROOT, 'sakkara')) and REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
        added = True
    try:
        yield
    finally:
        if added:
            sys.path.remove(REFERENCE_ROOT)


def reference_logp_dlogp(make_likelihood: Callable, df, floatX: str = 'float64'):
    """
    Build ``make_likelihood(api, pm, pt)`` with the reference on real PyMC and return
    ``(f, value_var_names)`` where ``f(theta) -> (logp, dlogp)`` takes the raveled unconstrained point.
    """
    if not available():
        raise RuntimeError('PyMC / sakkara are not importable: ' + describe())
    import pymc as pm
    import pytensor
    import pytensor.tensor as pt
    with _reference_on_path():
        import sakkara.model as api
        with pytensor.config.change_flags(floatX=floatX):
            model = api.build(df, make_likelihood(api, pm, pt))
            compiled = model.logp_dlogp_function()
    compiled.set_extra_values({})
    names = [v.name for v in model.value_vars]
    sizes = [int(np.prod(model.initial_point()[name].shape, dtype=np.int64)) for name in names]

    def f(theta) -> Tuple[float, np.ndarray]:
        theta = np.asarray(theta, dtype=floatX)
        assert theta.size == sum(sizes), (theta.size, sizes)
        try:                                    # pymc >= 5.18: raveled array in, (logp, dlogp) out
            logp, dlogp = compiled(theta)
        except Exception:                       # pymc 5.3: the point as RaveledVars / dict
            from pymc.blocking import DictToArrayBijection
            point, at = {}, 0
            for name, size in zip(names, sizes):
                point[name] = theta[at:at + size].reshape(model.initial_point()[name].shape)
                at += size
            logp, dlogp = compiled(DictToArrayBijection.map(point))
        return float(logp), np.asarray(dlogp, dtype=np.float64)

    return f, names
