"""Mirror of equiadapt.nbody."""
from . import canonicalization  # noqa: F401
