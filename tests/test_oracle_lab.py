"""CPU: the integer restatement of OpenCV's 8-bit RGB<->LAB (oracle/cv_lab.py, the specification the CUDA kernels
follow) is pinned exhaustively against cv2 on all 2**24 colours, both directions, plus the tissue-mask collapse."""
import cv2
import numpy as np
import pytest

from oracle import cv_lab


@pytest.fixture(scope="module")
def cube():
    r = np.arange(256, dtype=np.uint8)
    return np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(4096, 4096, 3)


def test_rgb2lab_exhaustive(cube):
    assert np.array_equal(cv2.cvtColor(cube, cv2.COLOR_RGB2LAB), cv_lab.rgb2lab_u8(cube))


def test_lab2rgb_exhaustive(cube):
    assert np.array_equal(cv2.cvtColor(cube, cv2.COLOR_LAB2RGB), cv_lab.lab2rgb_u8(cube))


def test_mask_collapses_to_y_index(cube):
    L = cv2.cvtColor(cube, cv2.COLOR_RGB2LAB)[..., 0]
    Yi = cv_lab.luminosity_y_index(cube)
    assert cv_lab.mask_y_bound(0.8) == 1146
    for thr in (0.0, 0.1, 0.5, 0.75, 0.8, 0.9, 1.0, 1.5):
        assert np.array_equal((L / 255.0) < thr, Yi <= cv_lab.mask_y_bound(thr)), thr


def test_generated_tables_are_current():
    """stainlib_b200/csrc/sb_tables.inc must be what oracle/gen_tables.py produces from these formulas."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "stainlib_b200", "csrc", "sb_tables.inc")).read()

    def arr(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", txt, re.S).group(1)
        return [float(x) for x in body.replace("\n", " ").split(",") if x.strip()]
    assert arr("SB_GAMMA_TAB") == cv_lab.srgb_gamma_tab().tolist()
    assert arr("SB_CBRT_TAB") == cv_lab.lab_cbrt_tab().tolist()
    assert arr("SB_LAB2YF_TAB") == cv_lab.lab_to_yf_tab().reshape(-1).tolist()
    assert arr("SB_INVGAMMA_TAB") == cv_lab.srgb_inv_gamma_tab().tolist()
    od = np.maximum(-np.log(np.maximum(np.arange(256), 1) / 255.0), 1e-6)
    assert arr("SB_OD_TAB") == od.tolist()
