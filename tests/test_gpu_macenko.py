"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, called through the reference-shaped Python API
and the C-ABI underneath, against the CPU oracle and the committed reference goldens.

Tolerances: tissue mask bit-exact; stain matrices <= 1e-5 abs; maxC <= 1e-5 rel; images within
1 uint8 LSB per channel (wrap-around counted as distance 1) with >= 99.9 % of bytes exactly equal.
"""
import numpy as np
import pytest
import torch

from oracle import stain_oracle as so
from sb_testutil import GOLDEN_CASES, lsb_stats
from stainlib_b200.synth import synth_tile, synth_batch, edge_case_tiles

pytestmark = pytest.mark.gpu

M_ATOL = 1e-5
MAXC_RTOL = 1e-5


@pytest.fixture(scope="module")
def sb(lib_built):
    import stainlib_b200
    return stainlib_b200


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mask_bit_exact_vs_golden(sb, golden, name):
    from stainlib_b200.utils.stain_utils import LuminosityThresholdTissueLocator as L
    src = golden[f"in/{name}/src"]
    assert np.array_equal(L.get_tissue_mask(src), golden[f"mask/{name}"])
    assert np.array_equal(L.get_tissue_mask(src, luminosity_threshold=0.75), golden[f"mask075/{name}"])


def test_mask_exhaustive_rgb_cube(sb):
    """All 2**24 colours, bit-exact against cv2 (the reference's own call, stain_utils.py:41-43)."""
    import cv2
    from stainlib_b200.utils.stain_utils import LuminosityThresholdTissueLocator as L
    r = np.arange(256, dtype=np.uint8)
    cube = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(4096, 4096, 3)
    for thr in (0.8, 0.5, 0.93):
        ref = (cv2.cvtColor(cube, cv2.COLOR_RGB2LAB)[..., 0] / 255.0) < thr
        got = L.get_tissue_mask(cube, luminosity_threshold=thr)
        assert np.array_equal(ref, got), thr


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_macenko_matrix_vs_golden(sb, golden, name):
    for which in ("src", "tgt"):
        M = sb.MacenkoStainExtractor.get_stain_matrix(golden[f"in/{name}/{which}"])
        assert M.shape == (2, 3) and M.dtype == np.float64
        np.testing.assert_allclose(M, golden[f"macenko_M/{name}/{which}"], rtol=0, atol=M_ATOL)
    M95 = sb.MacenkoStainExtractor.get_stain_matrix(golden[f"in/{name}/src"], angular_percentile=95)
    np.testing.assert_allclose(M95, golden[f"macenko_M95/{name}/src"], rtol=0, atol=M_ATOL)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_macenko_fit_transform_vs_golden(sb, golden, name):
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(golden[f"in/{name}/tgt"])
    np.testing.assert_allclose(n.stain_matrix_target, golden[f"macenko_norm/{name}/M_target"], rtol=0, atol=M_ATOL)
    np.testing.assert_allclose(n.maxC_target, golden[f"macenko_norm/{name}/maxC_target"], rtol=MAXC_RTOL)
    assert n.maxC_target.shape == (1, 2)
    out = n.transform(golden[f"in/{name}/src"])
    ref = golden[f"macenko_norm/{name}/out"]
    assert out.shape == ref.shape and out.dtype == np.uint8
    mx, frac = lsb_stats(out, ref, wrap=True)
    assert mx <= 1 and frac >= 0.999, (mx, frac)


def test_concentrations_vs_golden(sb, golden):
    from stainlib_b200.utils.stain_utils import get_concentrations
    for name in ("s_64", "ragged_96x80"):
        C = get_concentrations(golden[f"in/{name}/src"], golden[f"macenko_M/{name}/src"])
        np.testing.assert_allclose(C, golden[f"macenko_norm/{name}/conc_src"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("shape", [(256, 256), (512, 512), (300, 200), (64, 1000)])
def test_transform_vs_oracle_synthetic(sb, shape):
    H, W = shape
    tgt = synth_tile(101, H, W, kind="target")
    o = so.ExtractiveStainNormalizer("macenko")
    o.fit(tgt)
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(tgt)
    np.testing.assert_allclose(n.stain_matrix_target, o.stain_matrix_target, rtol=0, atol=M_ATOL)
    np.testing.assert_allclose(n.maxC_target, o.maxC_target, rtol=MAXC_RTOL)
    for seed in (5, 6):
        src = synth_tile(seed, H, W)
        mx, frac = lsb_stats(n.transform(src), o.transform(src), wrap=True)
        assert mx <= 1 and frac >= 0.999, (shape, seed, mx, frac)


def test_batch_equals_single_and_cluster_sizes(sb):
    """A tile's result must not depend on its position in a batch, on the cluster size that processed it, nor on the
    device/host entry point (sharded == unsharded determinism)."""
    tgt = synth_tile(1, 256, kind="target")
    batch = torch.from_numpy(synth_batch(200, 9, 256))
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(tgt)
    ref = n.transform(batch.cuda()).cpu()
    for i in range(batch.shape[0]):
        assert torch.equal(ref[i], torch.from_numpy(n.transform(batch[i].numpy())))
    host = n.transform(batch.pin_memory(), chunk_tiles=4)
    assert torch.equal(host, ref)
    for S in (1, 2, 4, 8):
        ns = sb.ExtractiveStainNormalizer("macenko", cluster_size=S)
        ns.fit(tgt)
        out = ns.transform(batch.cuda()).cpu()
        # per-tile sums are fixed-point integers: the cluster size must not change a single byte
        assert torch.equal(out, ref), (S, lsb_stats(out.numpy(), ref.numpy(), wrap=True))
        Ms = sb.MacenkoStainExtractor.get_stain_matrix(batch.cuda(), cluster_size=S)
        assert torch.equal(Ms, sb.MacenkoStainExtractor.get_stain_matrix(batch.cuda(), cluster_size=1)), S
        # splitting the batch (what sharding across GPUs does) changes nothing
        a = ns.transform(batch[:4].cuda()).cpu()
        b = ns.transform(batch[4:].cuda()).cpu()
        assert torch.equal(torch.cat([a, b]), out)


def test_edge_cases(sb, golden):
    from stainlib_b200.utils.excepts import TissueMaskException
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(golden["in/s_64/tgt"])
    with pytest.raises(TissueMaskException):
        sb.MacenkoStainExtractor.get_stain_matrix(golden["in/edge_all_white"])
    with pytest.raises(TissueMaskException):
        n.transform(golden["in/edge_all_white"])
    with pytest.raises(np.linalg.LinAlgError):
        sb.MacenkoStainExtractor.get_stain_matrix(golden["in/edge_one_tissue_pixel"])
    with pytest.raises(AssertionError):
        sb.MacenkoStainExtractor.get_stain_matrix(golden["in/s_64/src"].astype(np.float32))
    for name in ("dark", "saturated_bands"):
        I = golden[f"in/edge_{name}"]
        np.testing.assert_allclose(sb.MacenkoStainExtractor.get_stain_matrix(I), golden[f"macenko_M/edge_{name}"], atol=M_ATOL)
        mx, frac = lsb_stats(n.transform(I), golden[f"macenko_norm/edge_{name}/out"], wrap=True)
        assert mx <= 1 and frac >= 0.998, (name, mx, frac)
    # batched call: flagged tiles are reported through last_status and passed through, nothing raises
    batch = torch.from_numpy(np.stack([golden["in/edge_all_white"], golden["in/s_64/src"], golden["in/edge_one_tissue_pixel"]])).cuda()
    out = n.transform(batch)
    st = n.last_status.cpu().numpy()
    assert st[0] & 1 and st[1] == 0 and st[2] & 2
    assert torch.equal(out[0], batch[0]) and torch.equal(out[2], batch[2])


def test_wraparound_no_clip(sb):
    """normalizer.py:49-50 does not clip: with a negative entry in the target matrix 255*exp(.) exceeds 255 and wraps."""
    src = synth_tile(3, 128)
    o = so.ExtractiveStainNormalizer("macenko")
    n = sb.ExtractiveStainNormalizer("macenko")
    Mt = np.array([[0.60, 0.75, -0.28], [0.05, 0.95, 0.30]])
    Mt /= np.linalg.norm(Mt, axis=1)[:, None]
    for x in (o, n):
        x.stain_matrix_target = Mt
        x.maxC_target = np.array([[1.9, 1.2]])
    ref = o.transform(src)
    got = n.transform(src)
    assert (255 * np.exp(-so.get_concentrations(src, so.macenko_stain_matrix(src)) @ Mt)).max() > 256  # the case is real
    mx, frac = lsb_stats(got, ref, wrap=True)
    assert mx <= 1 and frac >= 0.999, (mx, frac)


def test_recombine_kernel_vs_oracle(sb):
    """K4 alone through the C ABI (sb_recombine) with per-tile source matrices and scales."""
    import ctypes
    from stainlib_b200 import _native as nv
    tiles = synth_batch(40, 5, 192, 160)
    Ms = np.stack([so.macenko_stain_matrix(t) for t in tiles])
    rng = np.random.default_rng(0)
    scale = rng.uniform(0.5, 1.8, size=(5, 2))
    Mt = so.macenko_stain_matrix(synth_tile(1, 128, kind="target"))
    b = nv.Batch(torch.from_numpy(tiles).cuda())
    out = b.new_like()
    dM, dS, dT = (torch.as_tensor(x, dtype=torch.float64).cuda() for x in (Ms, scale, Mt))
    nv.check(nv.load_library().sb_recombine(b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, nv.ptr(dM), nv.ptr(dS), nv.ptr(dT),
                                            0.01, nv.stream_ptr(b.idx)))
    got = out.cpu().numpy()
    for i in range(5):
        mx, frac = lsb_stats(got[i], so.recombine(tiles[i], Ms[i], scale[i], Mt), wrap=True)
        assert mx <= 1 and frac >= 0.999, (i, mx, frac)


def test_full_size_properties(sb):
    """BASELINE config-2 sized tiles (512x512), properties that need no oracle: a tile normalised to its OWN fitted
    statistics reproduces itself up to the LASSO residual, and the result is idempotent under batching."""
    tile = synth_tile(77, 512)
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(tile)
    batch = torch.from_numpy(np.stack([tile] * 6)).cuda()
    out = n.transform(batch)
    assert all(torch.equal(out[0], out[i]) for i in range(1, 6))
    # scale factors are exactly 1 => out = 255*exp(-C M): equal to the oracle's reconstruction
    ref = so.recombine(tile, so.macenko_stain_matrix(tile), [1.0, 1.0], so.macenko_stain_matrix(tile))
    mx, frac = lsb_stats(out[0].cpu().numpy(), ref, wrap=True)
    assert mx <= 1 and frac >= 0.999


def test_large_tile_mask_recompute_path(sb):
    """A slice of more than 262,144 pixels per CTA cannot cache its mask bits in shared memory and recomputes them
    per pass; same results required (1600x1400 px, cluster of 8 -> 280,000 px per CTA)."""
    src = synth_tile(91, 1600, 1400)
    tgt = synth_tile(92, 512, kind="target")
    o = so.ExtractiveStainNormalizer("macenko")
    o.fit(tgt)
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(tgt)
    np.testing.assert_allclose(sb.MacenkoStainExtractor.get_stain_matrix(src), so.macenko_stain_matrix(src), rtol=0, atol=M_ATOL)
    mx, frac = lsb_stats(n.transform(src), o.transform(src), wrap=True)
    assert mx <= 1 and frac >= 0.999, (mx, frac)


def _edge_batch_256():
    """256x256 tiles (large enough for the streaming passes) that exercise every exit of the streaming path: normal tiles,
    an all-white tile (EMPTY_MASK), one tissue pixel (FEW_TISSUE), a tile with too little tissue for the sampled brackets
    (-> fused fallback), a dark tile, saturated bands, a near-single-stain tile."""
    from stainlib_b200.synth import edge_case_tiles
    e = edge_case_tiles(256, 256)
    sparse = np.full((256, 256, 3), 255, np.uint8)
    sparse[:64, :128] = synth_tile(77, 64, 128)                      # ~6000 tissue pixels: below the 16,384 of the sampled path
    return np.stack([synth_tile(70, 256), e["all_white"], e["one_tissue_pixel"], sparse, e["dark"], e["saturated_bands"],
                     e["near_single_stain"], synth_tile(71, 256)])


@pytest.mark.parametrize("method", ["macenko", "vahadane"])
def test_streaming_path_equals_fused_path_on_edge_tiles(sb, method):
    """The streaming passes (automatic choice for aligned tiles of >= 32,768 pixels) and the fused per-tile kernel
    (cluster_size=1) must give the same bytes, stain matrices and status words, tile by tile, including flagged tiles and
    tiles the streaming path hands back to the fused kernel."""
    from stainlib_b200 import _native as nv
    tiles = torch.from_numpy(_edge_batch_256()).cuda()
    tgt = synth_tile(1, 256, kind="target")
    a = sb.ExtractiveStainNormalizer(method)
    b = sb.ExtractiveStainNormalizer(method, cluster_size=1)
    a.fit(tgt)
    b.fit(tgt)
    assert np.array_equal(a.stain_matrix_target, b.stain_matrix_target) and np.array_equal(a.maxC_target, b.maxC_target)
    nv.stream_fallbacks(reset=True)
    out_a, out_b = a.transform(tiles), b.transform(tiles)
    assert torch.equal(a.last_status, b.last_status)
    st = a.last_status.cpu().numpy()
    # (one tissue pixel: np.cov is NaN for Macenko -> FEW_TISSUE; the dictionary learner accepts it and the 99th percentile
    #  of its concentrations is zero -> ZERO_MAXC, output zeros like the reference's division by zero)
    assert st[0] == 0 and st[1] & 1 and st[2] != 0 and st[7] == 0
    assert (st[2] & 2) if method == "macenko" else (st[2] & 4)
    assert torch.equal(out_a, out_b)
    assert torch.equal(out_a[1], tiles[1])                                              # flagged tiles pass through
    if method == "macenko":
        assert torch.equal(out_a[2], tiles[2])
    assert nv.stream_fallbacks()[0] >= 1                                              # the sparse tile took the fused kernel
    ext = sb.MacenkoStainExtractor if method == "macenko" else sb.VahadaneStainExtractor
    Ma, Mb = ext.get_stain_matrix(tiles), ext.get_stain_matrix(tiles, cluster_size=1)
    assert torch.equal(torch.nan_to_num(Ma, nan=-7.0), torch.nan_to_num(Mb, nan=-7.0))
    assert bool(torch.isnan(Ma[1]).all()) and not bool(torch.isnan(Ma[0]).any())


def test_streaming_path_parameters(sb):
    """Non-default parameters through the streaming passes against the oracle: angular percentile, luminosity threshold."""
    src, tgt = synth_tile(81, 256), synth_tile(1, 256, kind="target")
    M = sb.MacenkoStainExtractor.get_stain_matrix(src, luminosity_threshold=0.7, angular_percentile=95)
    np.testing.assert_allclose(M, so.macenko_stain_matrix(src, luminosity_threshold=0.7, angular_percentile=95), atol=M_ATOL)
    batch = torch.from_numpy(np.stack([src, synth_tile(82, 256)])).cuda()
    Mb = sb.MacenkoStainExtractor.get_stain_matrix(batch, luminosity_threshold=0.7, angular_percentile=95)
    np.testing.assert_allclose(Mb[0].cpu().numpy(), M, atol=0)
