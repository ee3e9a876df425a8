"""
Distribution families understood by the B200 lowering (formulas in §A).

Each family is a callable with the calling convention of the PyMC distribution of the same
name -- ``Family(name, **params, shape=..., dims=..., observed=...)`` -- because that is how
``DistributionComponent.build_variable`` invokes its generator
(reference ``sakkara/model/composable/hierarchical/distribution.py:49-51``).  A component may be
given either one of these or the real ``pm.Normal`` etc.; :func:`resolve_generator` maps the latter
to the former by class name, so user code written for the reference keeps working unchanged.
"""
from typing import Any, Callable, Dict, Optional, Tuple

from sakkara_b200.model import sym


class Family:
    """
    :param name: family name, equal to the PyMC class name.
    :param params: accepted parameter names with their PyMC defaults (``None`` = required).
    :param positive: True when PyMC samples the variable on the log scale (LogTransform), i.e. the
        value variable is ``<name>_log__`` and the logp gains the log-Jacobian.
    :param discrete: True for families that can only be observed here (discrete ones, and StudentT).
    """

    def __init__(self, name: str, params: Dict[str, Any], positive: bool = False, discrete: bool = False):
        self.name = name
        self.__name__ = name
        self.params = params
        self.positive = positive
        self.discrete = discrete

    def __call__(self, name: str, *args, shape: Tuple[int, ...] = (), dims: Optional[Tuple[str, ...]] = None,
                 observed: Any = None, total_size: Any = None, **kwargs) -> sym.RandomVar:
        if args:
            # PyMC allows positional parameters in declaration order
            for key, value in zip(self.params, args):
                if key in kwargs:
                    raise TypeError(f'{self.name}: parameter {key} given twice')
                kwargs[key] = value
        params = {}
        for key, default in self.params.items():
            if key in kwargs:
                params[key] = kwargs.pop(key)
            elif default is not None:
                params[key] = default
        if kwargs:
            raise NotImplementedError(
                f'{self.name}: parameters {sorted(kwargs)} are not supported by the B200 lowering '
                f'(supported: {sorted(self.params)})')
        missing = [k for k, d in self.params.items() if d is None and k not in params]
        if self.name == 'Bernoulli':
            if ('p' in params) == ('logit_p' in params):
                raise ValueError('Bernoulli: give exactly one of p and logit_p')
        elif missing:
            raise TypeError(f'{self.name}: missing parameters {missing}')
        if isinstance(shape, int):
            shape = (shape,)
        if total_size is not None and observed is None:
            raise TypeError(f'{self.name}: total_size is a property of observed variables')
        var = sym.RandomVar(self.name, name, params, tuple(shape), dims, observed, total_size)
        return sym.Model.current().register(var)

    def __repr__(self):
        return f'sakkara_b200.dist.{self.name}'


Normal = Family('Normal', {'mu': 0.0, 'sigma': 1.0})
HalfNormal = Family('HalfNormal', {'sigma': 1.0}, positive=True)
Exponential = Family('Exponential', {'lam': None}, positive=True)
# families with constant parameters only (theta-dependent part on the device, normaliser in the plan's constant)
HalfCauchy = Family('HalfCauchy', {'beta': None}, positive=True)
Gamma = Family('Gamma', {'alpha': None, 'beta': None}, positive=True)
InverseGamma = Family('InverseGamma', {'alpha': None, 'beta': 1.0}, positive=True)
Cauchy = Family('Cauchy', {'alpha': None, 'beta': None})
Laplace = Family('Laplace', {'mu': None, 'b': None})
LogNormal = Family('LogNormal', {'mu': 0.0, 'sigma': 1.0}, positive=True)
HalfStudentT = Family('HalfStudentT', {'nu': None, 'sigma': 1.0}, positive=True)
# exactly one of p / logit_p must be given (checked in Family.__call__)
Bernoulli = Family('Bernoulli', {'p': None, 'logit_p': None}, discrete=True)
Poisson = Family('Poisson', {'mu': None}, discrete=True)
NegativeBinomial = Family('NegativeBinomial', {'mu': None, 'alpha': None}, discrete=True)
# (observed only: as a prior it would need three parameters)
StudentT = Family('StudentT', {'nu': None, 'mu': 0.0, 'sigma': 1.0}, discrete=True)

FAMILIES: Dict[str, Family] = {f.name: f for f in (Normal, HalfNormal, Exponential, HalfCauchy, Gamma, InverseGamma,
                                                   Cauchy, Laplace, LogNormal, HalfStudentT, Bernoulli, Poisson,
                                                   NegativeBinomial, StudentT)}


def resolve_generator(generator: Callable) -> Family:
    """
    Family for a generator given to a ``DistributionComponent``: one of the objects above, or a
    PyMC distribution class of the same name (``pm.Normal`` ...).
    """
    if isinstance(generator, Family):
        return generator
    name = getattr(generator, '__name__', None) or type(generator).__name__
    if name in FAMILIES:
        return FAMILIES[name]
    raise NotImplementedError(
        f'Distribution {name!r} cannot be lowered to the fused B200 kernel '
        f'(supported: {sorted(FAMILIES)}); build with the stock PyMC graph instead')
