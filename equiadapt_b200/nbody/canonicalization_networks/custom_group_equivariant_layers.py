"""Import path of equiadapt.nbody.canonicalization_networks.custom_group_equivariant_layers."""
from ...networks_frames import VNLeakyReLU, VNSoftplus  # noqa: F401
