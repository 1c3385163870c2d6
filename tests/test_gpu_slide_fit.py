"""GPU parity of the slide-level (multi-tile) Macenko fit: the five slide_pass kernels + the host selection logic
against the oracle's fit of the same tiles concatenated into one image (what the reference would compute)."""
import numpy as np
import pytest
import torch

from oracle import stain_oracle as so
from sb_testutil import lsb_stats
from stainlib_b200.synth import synth_tile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(lib_built):
    import stainlib_b200
    return stainlib_b200


def _oracle_fit(tiles):
    o = so.ExtractiveStainNormalizer("macenko")
    o.fit(np.concatenate(list(tiles), axis=0))
    return o


@pytest.mark.parametrize("shape,T", [((128, 128), 5), ((96, 112), 4), ((61, 53), 3), ((256, 256), 9)])
def test_slide_fit_vs_concatenated_oracle(sb, shape, T):
    tiles = np.stack([synth_tile(70 + i, *shape) for i in range(T)])
    tiles[-1] = 255                                           # a background-only tile contributes pixels but no tissue
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(torch.from_numpy(tiles).cuda())                     # a batch of more than one tile = one slide
    o = _oracle_fit(tiles)
    np.testing.assert_allclose(n.stain_matrix_target, o.stain_matrix_target, rtol=0, atol=1e-5)
    np.testing.assert_allclose(n.maxC_target, o.maxC_target, rtol=1e-5)
    src = synth_tile(3, 128)
    mx, frac = lsb_stats(n.transform(src), o.transform(src), wrap=True)
    assert mx <= 1 and frac >= 0.999, (mx, frac)


def test_slide_fit_of_one_tile_equals_tile_fit(sb):
    tgt = synth_tile(1, 256, kind="target")
    a, b = sb.ExtractiveStainNormalizer("macenko"), sb.ExtractiveStainNormalizer("macenko")
    a.fit(tgt)
    b.fit(torch.from_numpy(tgt[None]).cuda(), slide=True)
    np.testing.assert_allclose(b.stain_matrix_target, a.stain_matrix_target, rtol=0, atol=1e-6)
    np.testing.assert_allclose(b.maxC_target, a.maxC_target, rtol=1e-6)


def test_slide_fit_numpy_batch_and_errors(sb):
    from stainlib_b200.utils.excepts import TissueMaskException
    n = sb.ExtractiveStainNormalizer("macenko")
    with pytest.raises(TissueMaskException):
        n.fit(np.full((3, 64, 64, 3), 255, np.uint8))
    tiles = np.stack([synth_tile(90 + i, 64) for i in range(3)])
    n.fit(tiles)                                              # numpy [T,H,W,3] works as well
    o = _oracle_fit(tiles)
    np.testing.assert_allclose(n.stain_matrix_target, o.stain_matrix_target, rtol=0, atol=1e-5)


@pytest.mark.parametrize("shape,T", [((256, 256), 6), ((160, 176), 4), ((61, 53), 5)])
def test_vahadane_slide_fit_vs_oracle(sb, shape, T):
    from test_slide_fit_cpu import oracle_vahadane_slide
    tiles = np.stack([synth_tile(120 + i, *shape) for i in range(T)])
    tiles[-1] = 255
    n = sb.ExtractiveStainNormalizer("vahadane")
    n.fit(torch.from_numpy(tiles).cuda())
    M_ref, C_ref = oracle_vahadane_slide(list(tiles))
    np.testing.assert_allclose(n.stain_matrix_target, M_ref, rtol=0, atol=1e-4)
    np.testing.assert_allclose(n.maxC_target, C_ref, rtol=1e-3)
    src = synth_tile(3, 128)
    assert n.transform(src).shape == src.shape
