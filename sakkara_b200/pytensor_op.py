"""
PyTensor ``Op`` + PyMC glue: exposes the fused evaluator to ``pm.sample()``.

STATUS: PyMC / PyTensor are not installable in the build image (no wheel, no network),
so this module is import-guarded and has NOT been executed.  It is written against the documented
PyMC 5 / PyTensor 2 interfaces (``Op.make_node/perform/grad``, ``pm.Potential``, ``pm.Flat`` /
``pm.HalfFlat``).  Everything below this layer (DSL, lowering, plans, C ABI, kernels) is exercised by
the test-suite through :class:`sakkara_b200.valuegrad.ValueGradFunction`, which has the calling
convention of PyMC's own ``ValueGradFunction``.

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
  (the Jacobian terms our kernel already contains are subtracted again).
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
    except ImportError as err:          # pragma: no cover - not installable in the build image
        raise ImportError('to_pymc() needs PyMC >= 5 and PyTensor; use model.logp_dlogp_function() '
                          'with a sampler that takes a value-and-gradient callable instead') from err
    return pm, pytensor, pt, Apply, Op


def make_op(value_grad):
    """Wrap a :class:`ValueGradFunction` as a PyTensor Op ``theta -> (logp, dlogp)``."""
    pm, pytensor, pt, Apply, Op = _require_pymc()
    dtype = str(value_grad.dtype)

    class LogpDlogpOp(Op):
        __props__ = ()

        def make_node(self, theta):
            theta = pt.as_tensor_variable(theta).astype(dtype)
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


def to_pymc_model(b200_model, dtype: str = 'float64', device=None):
    """``pm.Model`` equivalent to the reference's model whose logp is evaluated by libskk_b200.so."""
    pm, pytensor, pt, Apply, Op = _require_pymc()
    plan = b200_model.plan
    value_grad = b200_model.logp_dlogp_function(dtype=dtype, device=device)
    op = make_op(value_grad)
    with pm.Model(coords={k: v for k, v in b200_model.coords.items() if k != 'obs'}) as model:
        pieces, jacobian = [], 0.0
        for block in plan.blocks:
            dims = block.dims if block.dims and 'obs' not in block.dims else None
            if block.transform == 'log':
                rv = pm.HalfFlat(block.name, shape=block.shape, dims=dims)
                unconstrained = pt.log(rv)
                jacobian = jacobian + unconstrained.sum()
            else:
                rv = pm.Flat(block.name, shape=block.shape, dims=dims)
                unconstrained = rv
            pieces.append(unconstrained.ravel())
        theta = pt.concatenate(pieces).astype(dtype)
        logp, _ = op(theta)
        # PyMC adds the log-Jacobian of HalfFlat's log transform itself; ours is inside `logp`
        pm.Potential('sakkara_b200_logp', logp - jacobian)
    model.b200_value_grad = value_grad
    return model
