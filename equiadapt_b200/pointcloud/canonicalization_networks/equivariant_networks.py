"""Import path of equiadapt.pointcloud.canonicalization_networks.equivariant_networks."""
from ...networks_frames import VNSmall  # noqa: F401
