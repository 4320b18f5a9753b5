from .euclidean_group import EuclideanGroupNBody  # noqa: F401
