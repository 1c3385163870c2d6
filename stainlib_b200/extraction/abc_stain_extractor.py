"""stainlib/extraction/abc_stain_extractor.py:3-12 (kept for import compatibility)."""
from stainlib_b200.utils.stain_utils import ABCStainExtractor  # noqa: F401
