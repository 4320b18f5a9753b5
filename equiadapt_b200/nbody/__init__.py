"""Mirror of equiadapt.nbody."""
from . import canonicalization  # noqa: F401
from . import canonicalization_networks  # noqa: F401
