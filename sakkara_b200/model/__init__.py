"""Public DSL, same names as ``sakkara.model`` (reference ``sakkara/model/__init__.py:1-10``)."""
from sakkara_b200.model.components import (GroupComponent, Likelihood, MinibatchLikelihood,
                                           DistributionComponent, Reshaper, DeterministicComponent,
                                           UnrepeatableComponent, DataComponent, data_components,
                                           FunctionComponent, f_, ModelComponent)
from sakkara_b200.model.build import build, B200Model
from sakkara_b200.model import dist, sym
