"""The drop-in claim: after ``install_as_stainlib()`` the reference's import paths, class names,
constructor / method signatures, defaults and argument errors are there.  CPU only -- nothing here touches a GPU."""
import inspect
import sys

import numpy as np
import pytest


@pytest.fixture()
def stainlib():
    from stainlib_b200.compat import install_as_stainlib
    before = {k: v for k, v in sys.modules.items() if k == "stainlib" or k.startswith("stainlib.")}
    pkg = install_as_stainlib()
    yield pkg
    for k in [k for k in sys.modules if k == "stainlib" or k.startswith("stainlib.")]:
        del sys.modules[k]
    sys.modules.update(before)


def _defaults(fn):
    return {k: v.default for k, v in inspect.signature(fn).parameters.items() if v.default is not inspect.Parameter.empty}


def test_import_paths_and_top_level_exports(stainlib):
    # stainlib/__init__.py:19-30
    for name in ("MacenkoStainExtractor", "VahadaneStainExtractor", "HedLighterColorAugmenter", "HedLightColorAugmenter",
                 "HedStrongColorAugmenter", "GrayscaleAugmentor", "ExtractiveStainNormalizer", "ReinhardStainNormalizer",
                 "LuminosityStandardizer"):
        assert hasattr(stainlib, name), name
    from stainlib.extraction.macenko_stain_extractor import MacenkoStainExtractor
    from stainlib.extraction.vahadane_stain_extractor import VahadaneStainExtractor
    from stainlib.normalization.normalizer import ExtractiveStainNormalizer, ReinhardStainNormalizer
    from stainlib.augmentation.augmenter import StainAugmentor, HedLightColorAugmenter, GrayscaleAugmentor       # noqa: F401
    from stainlib.utils.stain_utils import (LuminosityThresholdTissueLocator, LuminosityStandardizer, get_concentrations,   # noqa: F401
                                            convert_RGB_to_OD, convert_OD_to_RGB, normalize_matrix_rows, is_uint8_image,
                                            standardize_brightness, lab_split, merge_back, get_mean_std, get_sign, is_image,
                                            ABCStainExtractor, ABCTissueLocator)      # everything normalizer.py / augmenter.py import
    from stainlib.utils.excepts import TissueMaskException, InvalidRangeError                                     # noqa: F401
    assert stainlib.MacenkoStainExtractor is MacenkoStainExtractor and stainlib.ExtractiveStainNormalizer is ExtractiveStainNormalizer
    # north_star spellings
    for name in ("MacenkoExtractor", "VahadaneExtractor", "MacenkoNormalizer", "VahadaneNormalizer", "ReinhardNormalizer"):
        assert hasattr(stainlib, name), name
    assert StainAugmentor.transform is StainAugmentor.pop
    assert VahadaneStainExtractor is stainlib.VahadaneExtractor and ReinhardStainNormalizer is stainlib.ReinhardNormalizer


def test_signatures_and_defaults(stainlib):
    from stainlib.augmentation.augmenter import StainAugmentor
    from stainlib.utils.stain_utils import LuminosityStandardizer, LuminosityThresholdTissueLocator, get_concentrations
    d = _defaults(stainlib.MacenkoStainExtractor.get_stain_matrix)            # macenko_stain_extractor.py:7
    assert d["luminosity_threshold"] == 0.8 and d["angular_percentile"] == 99
    d = _defaults(stainlib.VahadaneStainExtractor.get_stain_matrix)           # vahadane_stain_extractor.py:19
    assert d["luminosity_threshold"] == 0.8 and d["regularizer"] == 0.1
    assert list(inspect.signature(stainlib.ExtractiveStainNormalizer.__init__).parameters)[:2] == ["self", "method"]
    assert _defaults(stainlib.ReinhardStainNormalizer.__init__) == {"target_means": 0, "target_stds": 0}        # normalizer.py:56
    d = _defaults(stainlib.ReinhardStainNormalizer.transform)                                                    # normalizer.py:70
    assert d == {"mask_background": False, "luminosity_threshold": 0.8}
    d = _defaults(StainAugmentor.__init__)                                                                       # augmenter.py:405
    assert d == {"sigma1": 0.2, "sigma2": 0.2, "augment_background": False}
    assert _defaults(get_concentrations)["regularizer"] == 0.01                                                  # stain_utils.py:69
    assert _defaults(LuminosityStandardizer.standardize)["percentile"] == 95                                     # stain_utils.py:53
    assert _defaults(LuminosityThresholdTissueLocator.get_tissue_mask)["luminosity_threshold"] == 0.8            # stain_utils.py:32
    for cls in (stainlib.HedLighterColorAugmenter, stainlib.HedLightColorAugmenter, stainlib.HedStrongColorAugmenter):
        a = cls()                                                                                                # augmenter.py:362-372
        assert callable(a.randomize) and callable(a.transform)
    for cls in (stainlib.ExtractiveStainNormalizer, stainlib.ReinhardStainNormalizer):
        assert callable(getattr(cls, "fit")) and callable(getattr(cls, "transform"))


def test_argument_errors_match_the_reference(stainlib):
    from stainlib.utils.excepts import InvalidRangeError
    from stainlib.augmentation.augmenter import HedColorAugmenter
    with pytest.raises(Exception, match="Method not recognized."):                 # normalizer.py:25
        stainlib.ExtractiveStainNormalizer("reinhard")
    assert stainlib.ExtractiveStainNormalizer("MaCeNkO").extractor is stainlib.MacenkoStainExtractor   # case-insensitive (normalizer.py:18-20)
    for bad in (np.zeros((8, 8, 3), np.float32), np.zeros((8, 8), np.uint8), np.zeros((8, 8, 4), np.uint8)):
        with pytest.raises(AssertionError, match="Image should be RGB uint8."):   # macenko_stain_extractor.py:16
            stainlib.MacenkoStainExtractor.get_stain_matrix(bad)
        with pytest.raises(AssertionError, match="Image should be RGB uint8."):
            stainlib.ExtractiveStainNormalizer("macenko").transform(bad)
    with pytest.raises(InvalidRangeError):                                          # augmenter.py:160-185
        HedColorAugmenter((-2.0, 0.1), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0.05, 0.95))
    with pytest.raises(InvalidRangeError):
        HedColorAugmenter((0.1, -0.1), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0.05, 0.95))
