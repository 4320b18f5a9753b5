"""Mirror of equiadapt.common (re-exports; implementation in canonicalizers_base / canonicalizers_frames)."""
from ..canonicalizers_base import (BaseCanonicalization, ContinuousGroupCanonicalization,
                                   DiscreteGroupCanonicalization, IdentityCanonicalization)
from ..canonicalizers_frames import gram_schmidt

__all__ = ["BaseCanonicalization", "ContinuousGroupCanonicalization", "DiscreteGroupCanonicalization",
           "IdentityCanonicalization", "gram_schmidt"]
