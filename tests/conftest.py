import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the REAL reference (see oracle/gen_golden.py for provenance of each key group)."""
    with np.load(GOLDEN_PATH, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}




@pytest.fixture(scope="session")
def lib_built():
    from stainlib_b200 import build
    return build.build()
