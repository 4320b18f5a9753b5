"""Import path of equiadapt.images.canonicalization.continuous_group."""
from ...canonicalizers_images_continuous import (ContinuousGroupImageCanonicalization,  # noqa: F401
                                                 OptimizedSteerableImageCanonicalization,
                                                 SteerableImageCanonicalization)
