"""Import path of equiadapt.pointcloud.canonicalization.continuous_group."""
from ...canonicalizers_frames import (ContinuousGroupPointcloudCanonicalization,  # noqa: F401
                                      EquivariantPointcloudCanonicalization)
