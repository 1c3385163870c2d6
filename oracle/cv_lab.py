"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Integer restatement of the 8-bit colour conversions the reference calls through OpenCV:

* ``cv.cvtColor(I, cv.COLOR_RGB2LAB)`` -- reference call sites ``stainlib/utils/stain_utils.py:41`` (tissue mask),
  ``:62`` (LuminosityStandardizer) and ``:152`` (lab_split, used by Reinhard);
* ``cv.cvtColor(I, cv.COLOR_LAB2RGB)`` -- reference call sites ``stainlib/utils/stain_utils.py:66`` and ``:172``.

OpenCV is a third-party dependency of the reference (pinned ``opencv-python 4.4.0.46`` in
``stainlib/utils/environment.yml:143``; 4.13.0 is installed here).  Its 8-bit sRGB<->CIELAB path is fixed-point table
arithmetic; this module restates that published algorithm (tables generated from their defining formulas) so the CUDA
kernels have an exact integer specification to follow.  The restatement is pinned exhaustively against ``cv2`` on all
2**24 colours in ``tests/test_oracle_lab.py`` (both directions, all three channels).

The generated tables are also what ``oracle/gen_tables.py`` writes into ``stainlib_b200/csrc/sb_tables.inc``.
"""
import numpy as np

GAMMA_SHIFT = 3
LAB_SHIFT = 12
LAB_SHIFT2 = 15
CBRT_TAB_SIZE = 256 * 3 // 2 * (1 << GAMMA_SHIFT)  # 3072
BASE_SHIFT = 14
BASE = 1 << BASE_SHIFT
INV_GAMMA_TAB_SIZE = 4096

# D65 white point and the sRGB->XYZ matrix, as used by OpenCV for the 8-bit Lab path
_D65 = (0.950456, 1.0, 1.088754)
_RGB2XYZ = ((0.412453, 0.357580, 0.180423),
            (0.212671, 0.715160, 0.072169),
            (0.019334, 0.119193, 0.950227))
_XYZ2RGB = ((3.240479, -1.53715, -0.498535),
            (-0.969256, 1.875991, 0.041556),
            (0.055648, -0.204043, 1.057311))


def _round_half_even(x):
    return np.rint(x)


def srgb_gamma_tab():
    """g[v] = round(255 * 8 * gamma(v/255)), v = 0..255 (u16).  Forward sRGB linearisation, scale 2040."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    lin = np.where(v <= np.float32(0.04045), v * np.float32(1.0 / 12.92),
                   np.power(((v.astype(np.float64) + 0.055) / 1.055), 2.4).astype(np.float32))
    return _round_half_even(np.float32(255.0 * (1 << GAMMA_SHIFT)) * lin.astype(np.float32)).astype(np.uint16)


def lab_cbrt_tab():
    """cb[i] = round(2**15 * f(i / 2040)), i < 3072, f the CIELAB companding function evaluated in float32."""
    x = np.arange(CBRT_TAB_SIZE, dtype=np.float32) / np.float32(255.0 * (1 << GAMMA_SHIFT))
    f = np.where(x < np.float32(0.008856), x * np.float32(7.787) + np.float32(0.13793103448275862),
                 np.cbrt(x.astype(np.float32)))
    return _round_half_even(np.float32(1 << LAB_SHIFT2) * f.astype(np.float32)).astype(np.uint16)


def rgb2lab_coeffs():
    """3x3 integer matrix: round(4096 * RGB2XYZ[i][j] / whitepoint[i])."""
    c = np.zeros((3, 3), dtype=np.int64)
    for i in range(3):
        for j in range(3):
            c[i, j] = int(np.rint((1 << LAB_SHIFT) * _RGB2XYZ[i][j] / _D65[i]))
    return c


L_SCALE = (116 * 255 + 50) // 100                      # 296
L_SHIFT = -((16 * 255 * (1 << LAB_SHIFT2) + 50) // 100)  # -1336934


def rgb2lab_u8(I):
    """8-bit RGB -> 8-bit LAB, integer path.  I: uint8 [...,3] -> uint8 [...,3]."""
    g = srgb_gamma_tab().astype(np.int64)
    cb = lab_cbrt_tab().astype(np.int64)
    C = rgb2lab_coeffs()
    I = np.asarray(I)
    R, G, B = g[I[..., 0]], g[I[..., 1]], g[I[..., 2]]
    rnd = 1 << (LAB_SHIFT - 1)
    fX = cb[(R * C[0, 0] + G * C[0, 1] + B * C[0, 2] + rnd) >> LAB_SHIFT]
    fY = cb[(R * C[1, 0] + G * C[1, 1] + B * C[1, 2] + rnd) >> LAB_SHIFT]
    fZ = cb[(R * C[2, 0] + G * C[2, 1] + B * C[2, 2] + rnd) >> LAB_SHIFT]
    rnd2 = 1 << (LAB_SHIFT2 - 1)
    L = (L_SCALE * fY + L_SHIFT + rnd2) >> LAB_SHIFT2
    a = (500 * (fX - fY) + 128 * (1 << LAB_SHIFT2) + rnd2) >> LAB_SHIFT2
    b = (200 * (fY - fZ) + 128 * (1 << LAB_SHIFT2) + rnd2) >> LAB_SHIFT2
    out = np.stack([L, a, b], axis=-1)
    return np.clip(out, 0, 255).astype(np.uint8)


def luminosity_y_index(I):
    """Yi = (871 g[R] + 2929 g[G] + 296 g[B] + 2048) >> 12 -- the index into the cbrt table that decides L."""
    g = srgb_gamma_tab().astype(np.int64)
    C = rgb2lab_coeffs()
    I = np.asarray(I)
    return (g[I[..., 0]] * C[1, 0] + g[I[..., 1]] * C[1, 1] + g[I[..., 2]] * C[1, 2] + (1 << (LAB_SHIFT - 1))) >> LAB_SHIFT


def l_of_y_index():
    """L (0..255) as a function of the Y index 0..3071 (monotone non-decreasing)."""
    cb = lab_cbrt_tab().astype(np.int64)
    L = (L_SCALE * cb + L_SHIFT + (1 << (LAB_SHIFT2 - 1))) >> LAB_SHIFT2
    return np.clip(L, 0, 255)


def mask_y_bound(luminosity_threshold):
    """Largest Y index whose L satisfies ``L/255.0 < luminosity_threshold`` (float64, as stain_utils.py:42-43).
    Returns -1 if no L qualifies.  For 0.8 this is 1146."""
    L = l_of_y_index()
    ok = (L / 255.0) < luminosity_threshold
    if not ok.any():
        return -1
    # monotone, so the qualifying indices are a prefix
    return int(np.nonzero(ok)[0].max())


# ------------------------------------------------------------------------------------------------ inverse direction
def lab_to_yf_tab():
    """LabToYF[L] = (round(BASE*Y), round(BASE*fY)) for L = 0..255, float32 arithmetic."""
    out = np.zeros((256, 2), dtype=np.int32)
    for i in range(256):
        li = np.float32(i) * np.float32(100.0) / np.float32(255.0)
        if li <= np.float32(8.0):  # lThresh = 0.008856*903.3
            y = li / np.float32(903.3)
            fy = np.float32(7.787) * y + np.float32(16.0) / np.float32(116.0)
        else:
            fy = (li + np.float32(16.0)) / np.float32(116.0)
            y = fy * fy * fy
        out[i, 0] = int(np.rint(np.float32(y) * np.float32(BASE)))
        out[i, 1] = int(np.rint(np.float32(fy) * np.float32(BASE)))
    return out


def _cdiv(a, b):
    """C truncating integer division on numpy int64 arrays (b > 0)."""
    a = np.asarray(a, dtype=np.int64)
    return np.where(a >= 0, a // b, -((-a) // b))


def ab_to_xz(t):
    """xz(t): t <= 3390 ? trunc(108 t / 841) - 290 (linear branch) : trunc(trunc(t*t/BASE) * t / BASE)."""
    t = np.asarray(t, dtype=np.int64)
    # linear branch of the inverse companding function: (t - BASE*16/116) * 108/841, in truncating integer division
    lin = _cdiv(t * 108, 841) - (BASE * 16 * 108 // 116 // 841)
    cub = _cdiv(_cdiv(t * t, BASE) * t, BASE)
    return np.where(t <= 3390, lin, cub)


def xyz2rgb_coeffs():
    """3x3 integer matrix: round(4096 * XYZ2RGB[i][j] * whitepoint[j])."""
    c = np.zeros((3, 3), dtype=np.int64)
    for i in range(3):
        for j in range(3):
            c[i, j] = int(np.rint((1 << LAB_SHIFT) * np.float32(_XYZ2RGB[i][j]) * np.float32(_D65[j])))
    return c


def srgb_inv_gamma_tab():
    """u8 table: round(255 * inverse_gamma(i/4096)), i < 4096."""
    x = np.arange(INV_GAMMA_TAB_SIZE, dtype=np.float64) / INV_GAMMA_TAB_SIZE
    x32 = x.astype(np.float32)
    y = np.where(x32 <= np.float32(0.0031308), x32 * np.float32(12.92),
                 (1.055 * np.power(x, 1.0 / 2.4) - 0.055).astype(np.float32))
    return np.clip(_round_half_even(np.float32(255.0) * y.astype(np.float32)), 0, 255).astype(np.uint8)


def lab2rgb_u8(LAB):
    """8-bit LAB -> 8-bit RGB, integer path.  LAB: uint8 [...,3] -> uint8 [...,3]."""
    LAB = np.asarray(LAB)
    yf = lab_to_yf_tab().astype(np.int64)
    inv = srgb_inv_gamma_tab()
    C = xyz2rgb_coeffs()
    L = LAB[..., 0].astype(np.int64)
    a = LAB[..., 1].astype(np.int64)
    b = LAB[..., 2].astype(np.int64)
    y = yf[L, 0]
    ify = yf[L, 1]
    adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * BASE // 500
    bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * BASE // 200 + 1
    x = ab_to_xz(ify + adiv)
    z = ab_to_xz(ify - bdiv)
    rnd = 1 << (BASE_SHIFT - 1)
    out = []
    for k in range(3):
        v = (C[k, 0] * x + C[k, 1] * y + C[k, 2] * z + rnd) >> BASE_SHIFT
        v = np.clip(v, 0, INV_GAMMA_TAB_SIZE - 1)
        out.append(inv[v])
    return np.stack(out, axis=-1).astype(np.uint8)
