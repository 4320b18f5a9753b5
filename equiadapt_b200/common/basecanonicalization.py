"""Import path of equiadapt.common.basecanonicalization."""
from ..canonicalizers_base import *  # noqa: F401,F403
from ..canonicalizers_base import (BaseCanonicalization, ContinuousGroupCanonicalization,  # noqa: F401
                                   DiscreteGroupCanonicalization, IdentityCanonicalization)
