"""Import path of equiadapt.common.utils (gram_schmidt only; LieParameterization is out of scope)."""
from ..canonicalizers_frames import gram_schmidt  # noqa: F401
