from .custom_equivariant_networks import VNDeepSets  # noqa: F401
