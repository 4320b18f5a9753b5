from .discrete_group import (DiscreteGroupImageCanonicalization, GroupEquivariantImageCanonicalization,  # noqa: F401
                             OptimizedGroupEquivariantImageCanonicalization)
