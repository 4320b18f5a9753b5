"""equiadapt_b200: the canonicalization hot path of arnab39/equiadapt as hand-written sm_100a CUDA.

Drop-in module tree (same import paths below the package name as `equiadapt`):
    equiadapt_b200.common.basecanonicalization, equiadapt_b200.common.utils,
    equiadapt_b200.images.canonicalization.discrete_group, equiadapt_b200.images.utils,
    equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks,
    equiadapt_b200.pointcloud.canonicalization.continuous_group,
    equiadapt_b200.nbody.canonicalization.euclidean_group
The kernels live in csrc/ behind the C ABI of include/equiadapt_b200.h (native.py binds it with ctypes).
"""
from . import common, images, nbody, pointcloud  # noqa: F401
from .canonicalizers_base import (BaseCanonicalization, ContinuousGroupCanonicalization,  # noqa: F401
                                  DiscreteGroupCanonicalization, IdentityCanonicalization)

__version__ = "0.1.0"
