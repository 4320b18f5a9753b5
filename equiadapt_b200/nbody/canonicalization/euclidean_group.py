"""Import path of equiadapt.nbody.canonicalization.euclidean_group."""
from ...canonicalizers_frames import EuclideanGroupNBody  # noqa: F401
