from .equivariant_networks import VNSmall  # noqa: F401
