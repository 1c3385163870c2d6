"""Shared helpers for the test-suite (kept out of conftest so they can be imported by name)."""
import numpy as np

GOLDEN_CASES = ["c1_256", "s_64", "ragged_96x80", "s_128", "odd_67x53"]


def lsb_stats(a, b, wrap=False):
    """max |a-b| and the fraction of exactly equal bytes.

    ``wrap=True`` counts uint8 wrap-around as distance 1 (255 vs 0).  That is right ONLY for the unclipped
    ``ExtractiveStainNormalizer.transform`` (normalizer.py:49-50: astype(uint8) wraps modulo 256, so a value of 255.9999 vs
    256.0001 is a 1-LSB disagreement that shows up as 255 vs 0).  Every clipped operator (HED, grayscale, StainAugmentor,
    Reinhard) must be compared with the plain absolute difference, or a saturation bug (a 255-LSB error) would pass."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    if wrap:
        d = np.minimum(d, 256 - d)
    return int(d.max()), float((d == 0).mean())
