"""Import path of equiadapt.nbody.canonicalization_networks.custom_equivariant_networks."""
from ...networks_frames import SequentialMultiple, VNDeepSetLayer, VNDeepSets  # noqa: F401
