"""B200-native stand-ins for `aps.transform` (same public names: aps/transform/__init__.py:1-2)."""
from .asr import FeatureTransform as AsrTransform
from .enh import FeatureTransform as EnhTransform

__all__ = ["AsrTransform", "EnhTransform"]
