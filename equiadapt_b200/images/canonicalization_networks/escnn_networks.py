"""Import path of equiadapt.images.canonicalization_networks.escnn_networks (ESCNNEquivariantNetwork only;
the steerable and wide-ResNet variants are out of scope)."""
from ...networks_images import ESCNNEquivariantNetwork  # noqa: F401
