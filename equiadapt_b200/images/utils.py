"""Import path of equiadapt.images.utils (feature-map group action)."""
from ..canonicalizers_images import get_action_on_image_features, group_element_to_index  # noqa: F401
