"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/stainb200.h declares.
No compute call is made (there is no GPU here); the product path must refuse to run without one."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "stainb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib_built):
    lib = ctypes.CDLL(lib_built)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in stainb200.h but not exported"


def test_binding_lists_every_symbol(lib_built):
    from stainlib_b200 import _native as nv
    assert sorted(nv.EXPORTS) == declared_symbols()
    lib = nv.load_library()
    assert lib.sb_version() >= 100
    p = nv.default_params()
    assert (p.luminosity_threshold, p.angular_percentile, p.lasso_lambda, p.conc_percentile, p.dl_lambda) == \
        (0.8, 99.0, 0.01, 99.0, 0.1)
    assert lib.sb_error_string(-1).decode() == "invalid argument"


def test_sass_is_sm100a(lib_built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stainlib_b200 import _native as nv
    from stainlib_b200 import ExtractiveStainNormalizer
    n = ExtractiveStainNormalizer("macenko")
    with pytest.raises(nv.NativeError):
        n.fit(np.zeros((16, 16, 3), np.uint8))


def test_product_does_not_import_oracle():
    """The product package must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "stainlib_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_reference_api_surface():
    import stainlib_b200 as s
    for name in ["MacenkoStainExtractor", "VahadaneStainExtractor", "HedLighterColorAugmenter", "HedLightColorAugmenter",
                 "HedStrongColorAugmenter", "GrayscaleAugmentor", "ExtractiveStainNormalizer", "ReinhardStainNormalizer",
                 "LuminosityStandardizer"]:
        assert hasattr(s, name)
    from stainlib_b200.utils.excepts import InvalidRangeError
    from stainlib_b200.augmentation.augmenter import HedColorAugmenter, StainAugmentor
    with pytest.raises(InvalidRangeError):
        HedColorAugmenter((-2, 0), None, None, None, None, None, None)
    with pytest.raises(Exception, match="Method not recognized."):
        s.ExtractiveStainNormalizer("reinhard")
    with pytest.raises(Exception, match="Method not recognized."):
        StainAugmentor("foo")
    h = s.HedLightColorAugmenter()
    assert h._sigmas == [-0.1] * 3 and h._biases == [-0.1] * 3 and tuple(h._cutoff_range) == (0.05, 0.95)
    np.random.seed(7)
    h.randomize()
    np.random.seed(7)
    expect = [np.random.uniform(-0.1, 0.1) for _ in range(6)]
    assert h._sigmas == expect[:3] and h._biases == expect[3:]
