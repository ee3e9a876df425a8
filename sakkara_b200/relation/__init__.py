"""Group algebra: infers parent/child/twin relations between data-frame columns and turns
"parameter defined on group g" into integer gather indices."""
from sakkara_b200.relation.group import Group
from sakkara_b200.relation.groupset import GroupSet, init, get_parent_df, get_twin_df
from sakkara_b200.relation.representation import (Representation, TensorRepresentation,
                                                  MinimalTensorRepresentation, UnrepeatableRepresentation)
