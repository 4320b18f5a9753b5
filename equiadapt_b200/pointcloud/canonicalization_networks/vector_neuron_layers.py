"""Import path of equiadapt.pointcloud.canonicalization_networks.vector_neuron_layers (the layers VNSmall uses)."""
from ...networks_frames import VNBatchNorm, VNLinearLeakyReLU  # noqa: F401
