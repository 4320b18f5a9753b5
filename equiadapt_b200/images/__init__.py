"""Mirror of equiadapt.images."""
from . import canonicalization, canonicalization_networks, utils  # noqa: F401
