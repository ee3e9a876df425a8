"""
``ValueGradFunction``: the callable PyMC's samplers invoke once per leapfrog step
(``pymc.Model.logp_dlogp_function()`` of the reference stack), evaluated by
libskk_b200.so.

    f = model.logp_dlogp_function(dtype='float64')
    logp, dlogp = f(theta)            # theta: raveled unconstrained parameters, creation order

A batch ``theta[B, P]`` evaluates B chains / leapfrog points in one call so that the row data is
read from HBM once per batch tile.  The device plan is created lazily in the calling process
(CUDA contexts do not survive ``fork``; PyMC runs one worker process per chain) and is dropped
when the object is pickled.
"""
import os
from typing import Optional

import numpy as np

from sakkara_b200.plan.kernel_plan import KernelPlan, make_kernel_plan
from sakkara_b200.plan.model_plan import ModelPlan


class ValueGradFunction:
    def __init__(self, plan: ModelPlan, dtype: str = 'float64', max_batch: int = 1, device: Optional[int] = None,
                 rows=None, n_rows_total: Optional[int] = None, narrow_y: bool = True):
        self.model_plan = plan
        self.dtype = np.dtype(dtype)
        self.max_batch = int(max_batch)
        self.device = device
        self._rows = rows
        self._n_rows_total = n_rows_total
        self._narrow_y = narrow_y
        self._kernel_plan: Optional[KernelPlan] = None
        self._skk = None
        self._pid = None

    # -- PyMC ValueGradFunction surface -------------------------------------------------------
    @property
    def size(self) -> int:
        return self.model_plan.n_params

    def set_extra_values(self, extra_vars) -> None:
        """All variables of the supported models are continuous; nothing to set."""

    def get_extra_values(self):
        return {}

    def __call__(self, theta, grad_out=None, extra_vars=None):
        theta = np.asarray(theta, dtype=self.dtype)
        logp, grad = self.skk.eval(theta)
        if grad_out is not None:
            grad_out[...] = grad
            return logp, grad_out
        return logp, grad

    # -- plan management ------------------------------------------------------------------------
    @property
    def kernel_plan(self) -> KernelPlan:
        if self._kernel_plan is None:
            self._kernel_plan = make_kernel_plan(self.model_plan, dtype=self.dtype, rows=self._rows,
                                                 n_rows_total=self._n_rows_total, narrow_y=self._narrow_y)
        return self._kernel_plan

    @property
    def skk(self):
        if self._skk is None or self._pid != os.getpid():
            from sakkara_b200.runtime import SkkPlan
            device = self.device
            if device is None:
                device = int(os.environ.get('LOCAL_RANK', '0')) if 'LOCAL_RANK' in os.environ else 0
            self._skk = SkkPlan(self.kernel_plan, max_batch=self.max_batch, device=device)
            self._pid = os.getpid()
        return self._skk

    def stats(self) -> dict:
        return self.skk.stats()

    def close(self) -> None:
        if self._skk is not None and self._pid == os.getpid():
            self._skk.close()
        self._skk = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_skk'] = None
        state['_pid'] = None
        return state
