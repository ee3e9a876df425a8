"""
``build(df, component)``: entry point of the DSL (reference ``sakkara/model/utils.py:9-33``).

Same steps as the reference -- copy the frame, add the ``global`` and ``obs`` columns, collect the
groups the component tree refers to, infer the group graph, build the tree inside a model
context -- but the context records into :mod:`sakkara_b200.model.sym`, and the recorded model is
lowered to a flat :class:`~sakkara_b200.plan.model_plan.ModelPlan` that the CUDA evaluator
consumes.  The returned :class:`B200Model` can be used as a context manager like a ``pm.Model``.
"""
from typing import Optional

import numpy as np
import pandas as pd

from sakkara_b200.model import sym
from sakkara_b200.model.components import ModelComponent
from sakkara_b200.relation.groupset import GroupSet, init


class B200Model:
    """
    A built model: the recorded variables, the group graph and the lowered plan.

    ``logp_dlogp_function()`` returns the callable PyMC's samplers use per leapfrog step
    (``ValueGradFunction`` semantics: raveled unconstrained theta in, ``(logp, dlogp)`` out),
    evaluated by the sm_100a kernels through the C ABI.
    """

    def __init__(self, recorded: sym.Model, groupset: GroupSet, plan: Optional['ModelPlan'], unsupported=None):
        self.recorded = recorded
        self.groupset = groupset
        self._plan = plan
        self._unsupported = unsupported

    @property
    def plan(self) -> 'ModelPlan':
        """
        The lowered model.  A component graph the fused kernels cannot express (no observed likelihood, a predictor
        that is not affine in theta, an unknown family ...) still builds -- like in the reference, where ``build()``
        takes any component (``tests/components/test_group_component.py``) -- and raises here, when something asks
        for its evaluation.
        """
        if self._plan is None:
            raise self._unsupported
        return self._plan

    @property
    def is_lowered(self) -> bool:
        return self._plan is not None

    # pm.Model-like surface --------------------------------------------------------------------
    @property
    def coords(self):
        return self.recorded.coords

    @property
    def free_RVs(self):
        return self.recorded.free_RVs

    @property
    def observed_RVs(self):
        return self.recorded.observed_RVs

    @property
    def named_vars(self):
        return self.recorded.named_vars

    def __getitem__(self, name: str):
        return self.recorded.named_vars[name]

    def __enter__(self) -> 'B200Model':
        self.recorded.__enter__()
        return self

    def __exit__(self, *exc) -> None:
        self.recorded.__exit__(*exc)

    @property
    def value_var_names(self):
        return [b.value_name for b in self.plan.blocks]

    def initial_point(self, dtype='float64') -> np.ndarray:
        """theta = 0 on the unconstrained scale (PyMC's default start for these families)."""
        return np.zeros(self.plan.n_params, dtype=dtype)

    def logp_dlogp_function(self, dtype: str = 'float64', max_batch: int = 1, device: Optional[int] = None,
                            **options):
        from sakkara_b200.valuegrad import MinibatchValueGrad, SplitSigmaValueGrad, ValueGradFunction
        if self.plan.likelihood.sigma_rows_element is not None:
            if self.plan.minibatch is not None:
                raise NotImplementedError('a mini-batched likelihood with a sigma that differs between rows')
            # sigma per group as a parameter: the row kernel gathers log(sigma) per row; rows with constant sigmas
            # next to them (or per_element=True) make it a sum of plans
            if (self.plan.likelihood.sigma_rows_element >= 0).all() and not options.get('per_element'):
                options.pop('per_element', None)
                f = ValueGradFunction(self.plan, dtype=dtype, max_batch=max_batch, device=device, **options)
                if not options and len(self.plan.terms) >= 3:
                    # (log(sigma) per row is one more term on its level: see the end of this function)
                    from sakkara_b200.plan.lower import UnsupportedModelError
                    from sakkara_b200.valuegrad import PartitionedValueGrad
                    try:
                        f.kernel_plan
                    except UnsupportedModelError as err:
                        if 'compiled kernels cover' not in str(err) and 'crossed grouping factors' not in str(err):
                            raise
                        return PartitionedValueGrad(self.plan, dtype=dtype, max_batch=max_batch, device=device)
                return f
            return SplitSigmaValueGrad(self.plan, dtype=dtype, max_batch=max_batch, device=device, **options)
        options.pop('per_element', None)      # (only a sigma that differs between rows has elements to split by)
        if self.plan.minibatch is not None:
            # MinibatchLikelihood: a fresh random subset of the rows per call, scaled to the whole data set
            return MinibatchValueGrad(self.plan, dtype=dtype, device=device, **options)
        from sakkara_b200.valuegrad import PartitionedValueGrad, partition_rows, secondary_chunk_groups
        chunk = options.pop('chunk_groups', None) or secondary_chunk_groups(dtype, max_batch)
        # (only a model with two gathers over more than `chunk` elements can have a secondary level -- the smaller of two
        #  crossed factors -- that large, and only one with three non-constant gathers can have three crossed factors:
        #  everything else skips the structure pass)
        gathers = [t for t in self.plan.terms if t.block < 0 or self.plan.blocks[t.block].size > 1]
        wide = sum(1 for t in gathers if t.block < 0 or self.plan.blocks[t.block].size > chunk)
        if not options and (wide >= 2 or len(gathers) >= 3) and len(partition_rows(self.plan, chunk)) > 1:
            # more than one plan: a secondary level too large for the shared-memory tables, or a third crossed factor
            return PartitionedValueGrad(self.plan, dtype=dtype, max_batch=max_batch, device=device, chunk_groups=chunk)
        f = ValueGradFunction(self.plan, dtype=dtype, max_batch=max_batch, device=device, **options)
        if not options and len(self.plan.terms) >= 3:
            # (2 global + 2 primary-level + 1 secondary-level terms is what one plan takes: with three terms or more ask the
            #  planner now; what exceeds it is evaluated with terms folded into each other, priors from a plan of their own)
            from sakkara_b200.plan.lower import UnsupportedModelError
            try:
                f.kernel_plan
            except UnsupportedModelError as err:
                if 'compiled kernels cover' not in str(err):
                    raise
                return PartitionedValueGrad(self.plan, dtype=dtype, max_batch=max_batch, device=device, chunk_groups=chunk)
        return f

    def sample(self, draws: int = 1000, tune: int = 1000, chains: int = 64, target_accept: float = 0.8,
               random_seed: int = 0, max_treedepth: int = 8, dtype: str = 'float64', device: Optional[int] = None):
        """
        ``pm.sample``'s arguments on the many-chain NUTS of :mod:`sakkara_b200.sampling`: all chains advance in
        lock-step on the device, one batched ``skk_eval`` per leapfrog step.  Returns an ``HMCResult`` (draws on the
        unconstrained scale, ``constrained(plan)``, sampler statistics, ``summary(plan)`` with ESS and R-hat).
        """
        from sakkara_b200.sampling import sample_nuts
        return sample_nuts(self, chains=chains, draws=draws, tune=tune, max_depth=max_treedepth,
                           target_accept=target_accept, dtype=dtype, seed=random_seed, device=device)

    def fit(self, n: int = 10_000, method: str = 'advi', learning_rate: float = 1e-3, random_seed: int = 0,
            dtype: str = 'float64', device: Optional[int] = None, **kwargs):
        """
        ``pm.fit``'s arguments on the mean-field ADVI of :mod:`sakkara_b200.variational` (a model with a
        ``MinibatchLikelihood`` evaluates a fresh random subset of the rows per step).  Returns the approximation
        (``mean``, ``sd``, ``elbo``, ``sample()``).
        """
        if method != 'advi':
            raise NotImplementedError(f"method {method!r}: only mean-field 'advi' is built here")
        from sakkara_b200.variational import fit_advi
        return fit_advi(self, n=n, learning_rate=learning_rate, seed=random_seed, dtype=dtype, device=device, **kwargs)

    def to_pymc(self, dtype: str = 'float64', device: Optional[int] = None):
        """A ``pm.Model`` with the same free variables whose joint logp is one fused-kernel Op
        (needs PyMC; see sakkara_b200/pytensor_op.py)."""
        from sakkara_b200.pytensor_op import to_pymc_model
        return to_pymc_model(self, dtype=dtype, device=device)


def build(df: pd.DataFrame, component: ModelComponent, backend: str = 'auto', dtype: str = 'float64',
          device: Optional[int] = None, strict: bool = True):
    """
    Build a model from a single component (typically a ``Likelihood``); all underlying components
    and their groups are traced from it.

    :param df: frame holding the columns that define the groups used by the components.
    :param component: component of the lowest hierarchy level.
    :param backend: ``'auto'`` (default) returns what the reference returns whenever that is possible: a
        ``pm.Model`` -- here one whose joint logp is the fused-kernel Op -- so that
        ``with build(df, likelihood): pm.sample()`` (reference ``README.md:33-34``) works unchanged; without an
        importable PyMC it returns the :class:`B200Model` (same context-manager surface, samplers take
        ``model.logp_dlogp_function()``).  ``'pymc'`` insists on the former, ``'b200'`` on the latter.
        The ``pm.Model`` carries the :class:`B200Model` as ``model.b200``.
    :param dtype: floating type of the evaluator behind the ``pm.Model`` (``pytensor.config.floatX`` on the
        reference side).
    :param device: CUDA device ordinal (default: ``LOCAL_RANK`` or 0).
    :param strict: a likelihood the fused kernels cannot evaluate raises ``UnsupportedModelError`` here (default);
        ``False`` returns the recorded model all the same -- variables, shapes, dims and the component tree can be
        inspected -- and the error is raised when its plan or ``logp_dlogp_function()`` is asked for.  There is no CPU
        fallback either way.
    """
    if backend not in ('auto', 'pymc', 'b200'):
        raise ValueError("backend must be 'auto', 'pymc' or 'b200'")
    from sakkara_b200.plan.lower import lower_model    # late import: plan.lower needs model.sym

    frame = df.copy()
    frame['global'] = 'global'
    frame['obs'] = np.arange(len(df))

    names = set(component.retrieve_groups()) | {'global', 'obs'}
    missing = names - set(frame.columns)
    if missing:
        raise KeyError(f'Group columns {sorted(missing)} are not in the data frame')
    groupset = init(frame.loc[:, [c for c in frame.columns if c in names]])

    with sym.Model(coords=groupset.coords()) as recorded:
        component.build(groupset)
    from sakkara_b200.plan.lower import UnsupportedModelError
    try:
        model = B200Model(recorded, groupset, lower_model(recorded))
    except UnsupportedModelError as err:
        if len(recorded.observed_RVs) == 1 and strict:
            raise           # a likelihood the kernels cannot evaluate: say so at once
        model = B200Model(recorded, groupset, None, err)    # a bare component: the recorded graph only
    if backend == 'b200' or not model.is_lowered:
        return model
    from sakkara_b200.pytensor_op import pymc_available, to_pymc_model
    if backend == 'pymc' or pymc_available():
        return to_pymc_model(model, dtype=dtype, device=device)
    return model
