"""Import path of equiadapt.images.canonicalization_networks.custom_group_equivariant_layers."""
from ...networks_images import (RotationEquivariantConv, RotationEquivariantConvLift,  # noqa: F401
                                RotoReflectionEquivariantConv, RotoReflectionEquivariantConvLift)
