"""Shared helpers for the test-suite (kept out of conftest so they can be imported by name)."""
import numpy as np

GOLDEN_CASES = ["c1_256", "s_64", "ragged_96x80", "s_128", "odd_67x53"]


def lsb_stats(a, b):
    """max |a-b| and fraction of exactly equal bytes, with uint8 wrap-around counted as distance 1 (255 vs 0)."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    d = np.minimum(d, 256 - d)
    return int(d.max()), float((d == 0).mean())
