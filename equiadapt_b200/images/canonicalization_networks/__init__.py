from .custom_equivariant_networks import CustomEquivariantNetwork  # noqa: F401
from .custom_group_equivariant_layers import (RotationEquivariantConv, RotationEquivariantConvLift,  # noqa: F401
                                              RotoReflectionEquivariantConv, RotoReflectionEquivariantConvLift)
from .escnn_networks import ESCNNEquivariantNetwork  # noqa: F401
