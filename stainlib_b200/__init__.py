"""stainlib_b200 -- B200-native (sm_100a) drop-in for the stain hot path of sebastianffx/stainlib.

Same module layout, class names and fit/transform/pop behaviour as ``stainlib`` (re-exports mirror
stainlib/__init__.py:19-30); ``stainlib_b200.compat.install_as_stainlib()`` additionally registers the package under
the name ``stainlib`` so existing ``from stainlib... import ...`` lines keep working.
"""
from .extraction.macenko_stain_extractor import MacenkoStainExtractor, MacenkoExtractor
from .extraction.vahadane_stain_extractor import VahadaneStainExtractor, VahadaneExtractor
from .normalization.normalizer import (ExtractiveStainNormalizer, ReinhardStainNormalizer, MacenkoNormalizer,
                                       VahadaneNormalizer, ReinhardNormalizer)
from .utils.stain_utils import LuminosityStandardizer
from .augmentation.augmenter import (HedLighterColorAugmenter, HedLightColorAugmenter, HedStrongColorAugmenter,
                                     GrayscaleAugmentor, StainAugmentor)

__version__ = "0.1.0"
