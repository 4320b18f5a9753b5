"""Import path of equiadapt.images.canonicalization.discrete_group."""
from ...canonicalizers_images import (DiscreteGroupImageCanonicalization,  # noqa: F401
                                      GroupEquivariantImageCanonicalization,
                                      OptimizedGroupEquivariantImageCanonicalization)
