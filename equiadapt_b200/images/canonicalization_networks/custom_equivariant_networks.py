"""Import path of equiadapt.images.canonicalization_networks.custom_equivariant_networks."""
from ...networks_images import CustomEquivariantNetwork  # noqa: F401
