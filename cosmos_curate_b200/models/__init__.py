"""ModelInterface implementations of the path (same class names as cosmos_curate/models/{clip,aesthetics,clip_aesthetics,transnetv2}.py)."""

from .clip import CLIPImageEmbeddings  # noqa: F401
from .aesthetics import AestheticScorer  # noqa: F401
from .clip_aesthetics import CLIPAestheticScorer  # noqa: F401
from .transnetv2 import TransNetV2  # noqa: F401
