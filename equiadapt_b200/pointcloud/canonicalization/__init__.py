from .continuous_group import (ContinuousGroupPointcloudCanonicalization,  # noqa: F401
                               EquivariantPointcloudCanonicalization)
