"""Mirror of equiadapt.pointcloud."""
from . import canonicalization  # noqa: F401
