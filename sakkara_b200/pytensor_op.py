"""
PyTensor ``Op`` + PyMC glue: exposes the fused evaluator to ``pm.sample()``.

When PyMC is importable, ``sakkara_b200.model.build(df, component)`` returns the ``pm.Model`` made here,
so the reference's usage -- ``with build(df, likelihood): idata = pm.sample()``
(``/root/reference/README.md:33-34``, ``sakkara/model/utils.py:31-33``) -- works verbatim.

STATUS: PyMC / PyTensor are not installable in the build image (no wheel, no network).
The module is written against the documented PyMC 5 / PyTensor 2 interfaces (``Op.make_node/perform/
grad``, ``pm.Potential``, ``pm.Flat`` / ``pm.HalfFlat``) and is executed in the test-suite against
``tests/pymc_stub.py``, an evaluating stand-in for exactly those interfaces
(``tests/test_pytensor_op.py``: raveled order, ``<name>_log__`` value variables, the Jacobian
bookkeeping, one evaluator call per point for value and gradient, pickling).

Design
------
* ``LogpDlogpOp``: one input ``theta[P]`` (raveled, unconstrained scale, creation order), two outputs
  ``(logp: scalar, dlogp: vector[P])``.  ``grad`` returns ``g_logp * dlogp`` where ``dlogp`` is the
  second output of the *same* application (PyTensor's merge rewrite deduplicates the two ``Apply``
  nodes), so value and gradient cost one kernel launch sequence.
* ``to_pymc_model``: a ``pm.Model`` with one free variable per parameter block -- same names, dims
  and transforms as the reference model, hence the same value variables (``<name>_log__``) in the same
  raveled order -- declared ``pm.Flat`` / ``pm.HalfFlat`` so that PyMC itself contributes only the
  log-Jacobian of the log transform; the joint log-posterior enters through one ``pm.Potential``
  (the Jacobian terms our kernel already contains are subtracted again).  Data columns are registered
  with ``pm.ConstantData`` under the reference's names (``sakkara/model/fixed/data.py:41``).
* The device plan is created lazily per process and dropped on pickling: PyMC runs one worker
  process per chain and a CUDA context does not survive ``fork``.
"""
import numpy as np


def _require_pymc():
    try:
        import pymc as pm
        import pytensor
        import pytensor.tensor as pt
        from pytensor.graph.basic import Apply
        from pytensor.graph.op import Op
    except ImportError as err:
        raise ImportError('to_pymc() needs PyMC >= 5 and PyTensor; use model.logp_dlogp_function() '
                          'with a sampler that takes a value-and-gradient callable instead') from err
    return pm, pytensor, pt, Apply, Op


def pymc_available() -> bool:
    try:
        _require_pymc()
    except ImportError:
        return False
    return True


def make_op(value_grad):
    """Wrap a :class:`ValueGradFunction` as a PyTensor Op ``theta -> (logp, dlogp)``."""
    pm, pytensor, pt, Apply, Op = _require_pymc()
    dtype = str(value_grad.dtype)

    class LogpDlogpOp(Op):
        __props__ = ()

        def make_node(self, theta):
            theta = pt.as_tensor_variable(theta)
            if str(theta.dtype) != dtype:
                theta = theta.astype(dtype)
            return Apply(self, [theta], [pt.scalar(dtype=dtype), pt.vector(dtype=dtype)])

        def perform(self, node, inputs, outputs):
            logp, grad = value_grad(np.asarray(inputs[0], dtype=dtype))
            outputs[0][0] = np.asarray(logp, dtype=dtype)
            outputs[1][0] = np.asarray(grad, dtype=dtype)

        def connection_pattern(self, node):
            return [[True, False]]          # only logp is differentiable w.r.t. theta

        def grad(self, inputs, output_grads):
            _, dlogp = self(*inputs)        # merged with the forward application by PyTensor
            return [output_grads[0] * dlogp]

    return LogpDlogpOp()


def to_pymc_model(b200_model, dtype: str = 'float64', device=None, value_grad=None):
    """
    ``pm.Model`` equivalent to the reference's model whose logp is evaluated by libskk_b200.so.

    :param value_grad: the evaluator behind the Op (default: ``b200_model.logp_dlogp_function()``, whose
        device plan is only created at the first evaluation, in the process that evaluates).
    """
    pm, pytensor, pt, Apply, Op = _require_pymc()
    plan = b200_model.plan
    if value_grad is None:
        value_grad = b200_model.logp_dlogp_function(dtype=dtype, device=device)
    op = make_op(value_grad)
    with pm.Model(coords=dict(b200_model.coords)) as model:
        for data in b200_model.recorded.data_vars:          # reference: fixed/data.py:40-41
            if hasattr(pm, 'ConstantData'):
                pm.ConstantData(data.name, data.values)
            elif hasattr(pm, 'Data'):
                pm.Data(data.name, data.values)
        pieces, jacobian = [], 0.0
        for block in plan.blocks:
            dims = tuple(block.dims) if block.dims else None
            kwargs = {'dims': dims} if dims else {'shape': block.shape}
            if block.transform == 'log':
                rv = pm.HalfFlat(block.name, **kwargs)
                unconstrained = pt.log(rv)
                jacobian = jacobian + unconstrained.sum()
            else:
                rv = pm.Flat(block.name, **kwargs)
                unconstrained = rv
            pieces.append(unconstrained.ravel())
        theta = pt.concatenate(pieces).astype(dtype)
        logp, _ = op(theta)
        # PyMC adds the log-Jacobian of HalfFlat's log transform itself; ours is inside `logp`
        pm.Potential('sakkara_b200_logp', logp - jacobian)
    model.b200 = b200_model
    model.b200_value_grad = value_grad
    return model
