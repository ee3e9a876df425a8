"""
TEST INFRASTRUCTURE -- a stand-in for the device behind ``ValueGradFunction`` so that the Python that composes
several plans (PartitionedValueGrad, SplitSigmaValueGrad, MinibatchValueGrad, the host-resident samplers, the ADVI
loop) runs end to end without a GPU.  ``install(monkeypatch)`` replaces ``ValueGradFunction.skk`` by an object with
``SkkPlan``'s host interface (``eval`` / ``close`` / ``stats``) that evaluates

* the likelihood from the function's own *KernelPlan* (``oracle/kernel_plan_eval.py``: sorted rows, segments, slot
  tables -- what the kernels are given), and
* the priors, when the plan includes them, from the NumPy oracle on the model plan without rows.

Nothing under ``sakkara_b200/`` knows about this file; the product path has no CPU fallback.
"""
import numpy as np

from oracle.kernel_plan_eval import likelihood_from_kernel_plan
from oracle.logp_numpy import logp_dlogp


class DeviceDouble:
    def __init__(self, function):
        self.function = function
        self.kernel_plan = function.kernel_plan
        self.priors = function.model_plan.take_rows([]) if function._include_priors else None
        self.dtype = function.dtype
        self.n_params = function.model_plan.n_params
        self.evals = 0

    def _one(self, theta):
        theta = np.asarray(theta, dtype=self.dtype).astype(np.float64)          # the point the device would see
        logp, grad = 0.0, np.zeros(self.n_params)
        if self.kernel_plan.n_rows:
            logp, grad = likelihood_from_kernel_plan(self.kernel_plan, theta)
            # (a plan that includes its priors carries their normalisers in its constant; the oracle below adds them)
        if self.priors is not None:
            prior_logp, prior_grad = logp_dlogp(self.priors, theta)
            logp, grad = logp + prior_logp, grad + prior_grad
        return logp, grad

    def eval(self, theta, time_kernels=False):
        self.evals += 1
        theta = np.asarray(theta)
        if theta.ndim == 1:
            logp, grad = self._one(theta)
            return self.dtype.type(logp), grad.astype(self.dtype)
        out = [self._one(t) for t in theta]
        return (np.array([o[0] for o in out], dtype=self.dtype), np.stack([o[1] for o in out]).astype(self.dtype))

    def stats(self):
        return {'evals': self.evals, 'launches': 2 * self.evals}

    def close(self):
        pass


def install(monkeypatch):
    from sakkara_b200.valuegrad import ValueGradFunction

    def skk(self):
        if self._skk is None:
            self._skk = DeviceDouble(self)
        return self._skk

    monkeypatch.setattr(ValueGradFunction, 'skk', property(skk))
    monkeypatch.setattr(ValueGradFunction, 'close', lambda self: setattr(self, '_skk', None))
