"""Registers stainlib_b200 under the import name ``stainlib`` so unmodified user code
(``from stainlib.normalization.normalizer import ExtractiveStainNormalizer`` ...) runs on the CUDA path."""
import importlib
import sys

_SUBMODULES = [
    "utils", "utils.excepts", "utils.stain_utils", "extraction", "extraction.abc_stain_extractor",
    "extraction.macenko_stain_extractor", "extraction.vahadane_stain_extractor", "normalization",
    "normalization.normalizer", "augmentation", "augmentation.augmenter",
]


def install_as_stainlib():
    pkg = importlib.import_module("stainlib_b200")
    sys.modules["stainlib"] = pkg
    for name in _SUBMODULES:
        sys.modules["stainlib." + name] = importlib.import_module("stainlib_b200." + name)
    return pkg
