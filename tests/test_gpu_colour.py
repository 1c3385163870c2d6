"""GPU parity for the LAB / HED / augmentation rows of SURVEY section 8 (a7, a8, a9, a10, a11, f1, f4).

Reinhard and the luminosity standardiser are integer/LUT pipelines on the GPU exactly as in OpenCV, so the bar is
bit-exact; HED, grayscale and StainAugmentor are fp32 vs the fp64 reference: <= 1 LSB, >= 99.9 % of bytes equal."""
import numpy as np
import pytest
import torch

from oracle import stain_oracle as so
from sb_testutil import GOLDEN_CASES, lsb_stats
from stainlib_b200.synth import synth_tile, synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(lib_built):
    import stainlib_b200
    return stainlib_b200


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reinhard_vs_golden(sb, golden, name):
    r = sb.ReinhardStainNormalizer()
    r.fit(golden[f"in/{name}/tgt"])
    np.testing.assert_allclose(np.array(r.target_means).reshape(3), golden[f"reinhard/{name}/means"], rtol=1e-12)
    np.testing.assert_allclose(np.array(r.target_stds).reshape(3), golden[f"reinhard/{name}/stds"], rtol=1e-12)
    assert r.target_means[0].shape == (1, 1)
    src = golden[f"in/{name}/src"]
    assert np.array_equal(r.transform(src), golden[f"reinhard/{name}/out"])
    assert np.array_equal(r.transform(src, mask_background=True), golden[f"reinhard/{name}/out_masked"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_luminosity_standardizer_vs_golden(sb, golden, name):
    assert np.array_equal(sb.LuminosityStandardizer.standardize(golden[f"in/{name}/src"]), golden[f"lumstd/{name}"])


def test_lab_paths_exhaustive_cube(sb):
    """Every 8-bit colour through the CUDA integer RGB->LAB->RGB path (via the standardiser and Reinhard), bit-exact
    against the cv2-based oracle."""
    r = np.arange(256, dtype=np.uint8)
    cube = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(4096, 4096, 3)
    assert np.array_equal(sb.LuminosityStandardizer.standardize(cube, percentile=80), so.luminosity_standardize(cube, 80))
    tgt = synth_tile(1, 128, kind="target")
    a, b = sb.ReinhardStainNormalizer(), so.ReinhardStainNormalizer()
    a.fit(tgt)
    b.fit(tgt)
    assert np.array_equal(a.transform(cube), b.transform(cube))
    assert np.array_equal(a.transform(cube, mask_background=True, luminosity_threshold=0.7),
                          b.transform(cube, mask_background=True, luminosity_threshold=0.7))


def test_reinhard_batch_and_edges(sb, golden):
    from stainlib_b200.utils.excepts import TissueMaskException
    tgt = golden["in/s_64/tgt"]
    r, o = sb.ReinhardStainNormalizer(), so.ReinhardStainNormalizer()
    r.fit(tgt)
    o.fit(tgt)
    batch = synth_batch(900, 7, 96, 112)
    out = r.transform(torch.from_numpy(batch).cuda()).cpu().numpy()
    for i in range(7):
        assert np.array_equal(out[i], o.transform(batch[i]))
    for name in ("one_tissue_pixel", "saturated_bands", "near_single_stain", "dark"):
        I = golden[f"in/edge_{name}"]
        assert np.array_equal(r.transform(I), golden[f"reinhard/edge_{name}/out"])
    with pytest.raises(TissueMaskException):
        r.transform(golden["in/edge_all_white"], mask_background=True)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_hed_vs_golden(sb, golden, name):
    src = golden[f"in/{name}/src"]
    for cls, tag in ((sb.HedLighterColorAugmenter, "lighter"), (sb.HedLightColorAugmenter, "light"),
                     (sb.HedStrongColorAugmenter, "strong")):
        h = cls()
        mx, frac = lsb_stats(h.transform(src), golden[f"hed/{name}/{tag}/default"])
        assert mx <= 1 and frac >= 0.999, (tag, "default", mx, frac)
        np.random.seed(7)
        h.randomize()
        assert np.allclose(h._sigmas, golden[f"hed/{name}/{tag}/sigmas"]) and np.allclose(h._biases, golden[f"hed/{name}/{tag}/biases"])
        mx, frac = lsb_stats(h.transform(src), golden[f"hed/{name}/{tag}/out"])
        assert mx <= 1 and frac >= 0.999, (tag, mx, frac)


def test_hed_cutoff_and_batch(sb, golden):
    h = sb.HedLightColorAugmenter()
    white = golden["in/edge_all_white"]
    assert h.transform(white) is white                      # augmenter.py:329-331 returns the same object
    tiles = synth_batch(50, 4, 80, 96)
    tiles[3] = 255
    rng = np.random.default_rng(3)
    sig, bia = rng.uniform(-0.1, 0.1, (4, 3)), rng.uniform(-0.1, 0.1, (4, 3))
    out = h.transform(torch.from_numpy(tiles).cuda(), sigmas=sig, biases=bia).cpu().numpy()
    for i in range(4):
        ref = so.hed_augment(tiles[i], sig[i], bia[i])
        mx, frac = lsb_stats(out[i], ref)
        assert mx <= 1 and frac >= 0.999, (i, mx, frac)
    assert h.last_status.cpu().tolist() == [0, 0, 0, 1]
    # large tiles split over many CTAs: the last CTA of a tile decides the gate and restores the input of a skipped tile
    big = synth_batch(60, 3, 512, 512)
    big[1] = 255
    big[2] = (big[2].astype(np.float32) * 0.04).astype(np.uint8)          # mean below the 0.05 cutoff
    sig, bia = rng.uniform(-0.1, 0.1, (3, 3)), rng.uniform(-0.1, 0.1, (3, 3))
    out = h.transform(torch.from_numpy(big).cuda(), sigmas=sig, biases=bia).cpu().numpy()
    assert h.last_status.cpu().tolist() == [0, 1, 1]
    assert np.array_equal(out[1], big[1]) and np.array_equal(out[2], big[2])
    mx, frac = lsb_stats(out[0], so.hed_augment(big[0], sig[0], bia[0]))
    assert mx <= 1 and frac >= 0.999, (mx, frac)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_grayscale_vs_golden(sb, golden, name):
    g = sb.GrayscaleAugmentor()
    g.fit(golden[f"in/{name}/src"])
    np.random.seed(5)
    mx, frac = lsb_stats(g.pop(), golden[f"gray/{name}/pop"])
    assert mx <= 1 and frac >= 0.999, (mx, frac)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_stain_augmentor_vs_golden(sb, golden, name):
    from stainlib_b200.augmentation.augmenter import StainAugmentor
    src = golden[f"in/{name}/src"]
    a = StainAugmentor("macenko")
    a.fit(src)
    np.testing.assert_allclose(a.stain_matrix, golden[f"macenko_M/{name}/src"], atol=1e-5)
    np.random.seed(1234)
    for k in ("pop0", "pop1"):
        mx, frac = lsb_stats(a.pop(), golden[f"stain_aug/{name}/{k}"])
        assert mx <= 1 and frac >= 0.999, (k, mx, frac)
    a = StainAugmentor("macenko", sigma1=0.4, sigma2=0.3, augment_background=True)
    a.fit(src)
    np.random.seed(99)
    mx, frac = lsb_stats(a.pop(), golden[f"stain_aug/{name}/pop_bg"])
    assert mx <= 1 and frac >= 0.999, (mx, frac)


def test_stain_augmentor_batch(sb):
    from stainlib_b200.augmentation.augmenter import StainAugmentor
    tiles = synth_batch(70, 3, 128)
    a = StainAugmentor("macenko")
    a.fit(torch.from_numpy(tiles).cuda())
    al, be = np.array([[1.1, 0.9], [0.85, 1.15], [1.0, 1.0]]), np.array([[0.05, -0.1], [-0.15, 0.1], [0.0, 0.0]])
    out = a.pop(alphas=al, betas=be).cpu().numpy()
    for i in range(3):
        o = so.StainAugmentor("macenko")
        o.fit(tiles[i])
        mx, frac = lsb_stats(out[i], o.pop(alphas=al[i], betas=be[i]))
        assert mx <= 1 and frac >= 0.999, (i, mx, frac)


@pytest.mark.parametrize("shape", [(128, 128), (67, 53)])
def test_hed_skimage018_variant(sb, shape):
    """scikit-image >= 0.18 definition of rgb2hed / hed2rgb (the one an unpinned install of the reference uses today):
    aligned tiles take the TMA ring operator, the odd shape takes the register-staged kernel.  <= 1 LSB, >= 99.9 % equal
    (plain difference: the operator clips)."""
    rng = np.random.default_rng(5)
    tiles = synth_batch(70, 3, *shape)
    sig, bia = rng.uniform(-0.1, 0.1, (3, 3)), rng.uniform(-0.1, 0.1, (3, 3))
    h = sb.HedLightColorAugmenter()
    out = h.transform(torch.from_numpy(tiles).cuda(), sigmas=sig, biases=bia, skimage_version="0.18").cpu().numpy()
    for i in range(3):
        ref = so.hed_augment(tiles[i], sig[i], bia[i], skimage_version="0.18")
        mx, frac = lsb_stats(out[i], ref)
        assert mx <= 1 and frac >= 0.999, (i, mx, frac)
    one = h.transform(tiles[0], sigmas=sig[0], biases=bia[0], skimage_version="0.18")
    assert np.array_equal(one, out[0])
    # the two definitions are different functions: the variant switch must actually change the result
    assert not np.array_equal(out[0], h.transform(tiles[0], sigmas=sig[0], biases=bia[0]))


@pytest.mark.parametrize("version", ["0.17", "0.18"])
def test_hed_float_patches(sb, version):
    """augmenter.py:288-291, 323-327: a float patch in [0,1] is transformed as it is and a float patch comes back.
    Tolerance 5e-6 absolute (fp32 log2f / exp2f against the float64 restatement)."""
    rng = np.random.default_rng(9)
    tile = synth_batch(80, 1, 96, 80)[0]
    for dt in (np.float32, np.float64):
        patch = (tile.astype(np.float64) / 255.0).astype(dt)
        h = sb.HedLightColorAugmenter()
        np.random.seed(3)
        h.randomize()
        got = h.transform(patch, skimage_version=version)
        ref = so.hed_augment(patch, h._sigmas, h._biases, skimage_version=version)
        assert got.dtype == dt and got.shape == patch.shape
        assert np.abs(got.astype(np.float64) - ref).max() < 5e-6
    white = np.ones((32, 32, 3), np.float32)
    assert sb.HedLightColorAugmenter().transform(white) is white          # outside the cutoff: same object
    tb = torch.from_numpy(np.stack([patch.astype(np.float32), white[:1].repeat(96, 0).repeat(80, 1)[:96, :80]])).cuda()
    sig, bia = rng.uniform(-0.1, 0.1, (2, 3)), rng.uniform(-0.1, 0.1, (2, 3))
    h = sb.HedLightColorAugmenter()
    out = h.transform(tb, sigmas=sig, biases=bia, skimage_version=version)
    assert out.is_cuda and out.dtype == torch.float32 and h.last_status.cpu().tolist() == [0, 1]
    assert torch.equal(out[1], tb[1])
    ref = so.hed_augment(tb[0].cpu().numpy(), sig[0], bia[0], skimage_version=version)
    assert np.abs(out[0].cpu().numpy().astype(np.float64) - ref).max() < 5e-6


def test_convert_rgb_od_roundtrip(sb, golden):
    """convert_RGB_to_OD / convert_OD_to_RGB (stain_utils.py:101-124) through their CUDA entry points, against the
    reference's own output (golden od/*) and the oracle."""
    from stainlib_b200.utils.stain_utils import convert_OD_to_RGB, convert_RGB_to_OD
    src = golden["in/s_64/src"]
    od = convert_RGB_to_OD(src)
    assert od.dtype == np.float64 and od.shape == src.shape
    assert np.array_equal(od, golden["od/s_64"])                          # the reference's own output; bit-exact: a 256-entry float64 table
    odd = golden["in/odd_67x53/src"]                                      # a size that is not a multiple of 16 values
    assert np.array_equal(convert_RGB_to_OD(odd), so.convert_RGB_to_OD(odd))
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    assert np.array_equal(convert_RGB_to_OD(ramp), so.convert_RGB_to_OD(ramp))
    batch = torch.from_numpy(synth_batch(90, 2, 40, 56)).cuda()
    odb = convert_RGB_to_OD(batch)
    assert odb.is_cuda and odb.dtype == torch.float64 and odb.shape == batch.shape
    assert np.array_equal(odb.cpu().numpy(), so.convert_RGB_to_OD(batch.cpu().numpy()))
    od32 = convert_RGB_to_OD(batch, dtype=torch.float32)
    assert od32.dtype == torch.float32 and torch.equal(od32, odb.float())
    # OD -> RGB: exp in float64 on both sides; a value may differ by one where 255*exp(-OD) is within an ulp of an integer
    rng = np.random.default_rng(0)
    OD = rng.uniform(0.0, 5.6, (33, 47, 3))
    got, ref = convert_OD_to_RGB(OD), so.convert_OD_to_RGB(OD)
    assert got.dtype == np.uint8 and np.array_equal(got, ref)
    mx, frac = lsb_stats(convert_OD_to_RGB(od), so.convert_OD_to_RGB(golden["od/s_64"]))
    assert mx <= 1 and frac >= 0.98, (mx, frac)
    with pytest.raises(AssertionError):
        convert_OD_to_RGB(np.array([[[0.1, -0.2, 0.3]]]))


def test_caller_owned_workspace(sb):
    """sb_workspace_bytes / sb_set_workspace: with a lent workspace the results are the same bytes."""
    from stainlib_b200 import _native as nv
    tgt = synth_tile(1, 128, kind="target")
    batch = torch.from_numpy(synth_batch(120, 6, 128)).cuda()
    n = sb.ExtractiveStainNormalizer("vahadane")
    n.fit(tgt)
    ref = n.transform(batch)
    need = nv.workspace_bytes(6, 128, 128)
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    nv.set_workspace(ws)
    try:
        assert torch.equal(n.transform(batch), ref)
        h = sb.HedLightColorAugmenter()
        a = h.transform(batch)
        nv.set_workspace(None)
        assert torch.equal(h.transform(batch), a)
    finally:
        nv.set_workspace(None)


def test_two_devices_one_process(sb):
    """One handle per device in one process: kernels with > 48 KB of dynamic shared memory must opt in on EVERY device,
    entry points must run on their handle's device and leave the caller's current device alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tgt = synth_tile(1, 128, kind="target")
    batch = torch.from_numpy(synth_batch(130, 3, 128))
    n = sb.ExtractiveStainNormalizer("macenko")
    n.fit(tgt)
    a = n.transform(batch.to("cuda:0"))
    assert torch.cuda.current_device() == 0
    b = n.transform(batch.to("cuda:1"))
    assert torch.cuda.current_device() == 0 and b.device.index == 1
    assert torch.equal(a.cpu(), b.cpu())
    r = sb.ReinhardStainNormalizer()
    r.fit(tgt)
    assert torch.equal(r.transform(batch.to("cuda:1")).cpu(), r.transform(batch.to("cuda:0")).cpu())
    h = sb.HedLightColorAugmenter()
    assert torch.equal(h.transform(batch.to("cuda:1")).cpu(), h.transform(batch.to("cuda:0")).cpu())


@pytest.mark.parametrize("name", ["s_64", "odd_67x53"])
def test_lab_utilities_vs_reference(sb, golden, name):
    """standardize_brightness / lab_split / get_mean_std / merge_back (stain_utils.py:146-194), each through its own CUDA
    entry point, against the reference's own outputs: bit-exact (integer LAB, float32 / float64 arithmetic as numpy's)."""
    from stainlib_b200.utils.stain_utils import get_mean_std, lab_split, merge_back, standardize_brightness
    src = golden[f"in/{name}/src"]
    assert np.array_equal(standardize_brightness(src), golden[f"labutil/{name}/bright"])
    I1, I2, I3 = lab_split(src)
    for got, key in ((I1, "I1"), (I2, "I2"), (I3, "I3")):
        assert got.dtype == np.float32 and np.array_equal(got, golden[f"labutil/{name}/{key}"])
    m, sd = get_mean_std(src)
    np.testing.assert_allclose(np.array(m).reshape(3), golden[f"labutil/{name}/means"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.array(sd).reshape(3), golden[f"labutil/{name}/stds"], rtol=1e-12, atol=1e-12)
    assert m[0].shape == (1, 1) and sd[2].shape == (1, 1)
    J = [(I1 * 0.9 + 3.0), (I2 * 1.1 - 2.0), (I3 * 0.8 + 1.5)]
    for dt, key in ((np.float32, "merge_f32"), (np.float64, "merge_f64")):
        planes = [x.astype(dt).copy() for x in J]
        keep = planes[0].copy()
        assert np.array_equal(merge_back(*planes), golden[f"labutil/{name}/{key}"])
        assert np.array_equal(planes[0], keep * dt(2.55) if dt is np.float32 else keep * 2.55)      # scaled in place like the reference
    # batched tensors
    batch = torch.from_numpy(np.stack([src, src[::-1].copy()])).cuda()
    P = lab_split(batch)
    assert P[0].shape == batch.shape[:3] and np.array_equal(P[0][0].cpu().numpy(), golden[f"labutil/{name}/I1"])
    mb, sb_ = get_mean_std(batch)
    np.testing.assert_allclose(mb[0].cpu().numpy(), golden[f"labutil/{name}/means"], rtol=1e-12)
    assert torch.equal(merge_back(*P)[1].cpu(), torch.from_numpy(so.merge_back(*[p[1].cpu().numpy().copy() for p in P])))


def _unaligned_copy(t):
    """The same tile batch at a data pointer that is NOT a multiple of 16: routes an operator to its register-staged kernel."""
    flat = torch.empty(t.numel() + 16, dtype=torch.uint8, device=t.device)
    v = flat[1:1 + t.numel()].view(t.shape)
    v.copy_(t)
    assert v.data_ptr() % 16 != 0 and v.is_contiguous()
    return v


@pytest.mark.parametrize("shape", [(5, 128, 160), (3, 512, 512), (2, 272, 1008), (4, 128, 128), (300, 128, 144), (1, 1040, 1024)])
def test_reinhard_ring_passes_equal_tile_kernel(sb, shape, monkeypatch):
    """The streaming passes of sb_reinhard.cu (aligned tiles) and lab_tile_kernel give the same bytes and the same statistics:
    fit, transform (both mask modes), luminosity standardiser.  The tile kernel is reached through an unaligned copy of the
    batch and through the diagnostic switch SB_REINHARD_TILE_KERNEL."""
    B, H, W = shape
    x = torch.from_numpy(synth_batch(4100, B, H, W)).cuda()
    x[0, : H // 2] = 255                               # a half-white tile
    if B > 2:
        x[2] = 255                                     # an all-background tile: EMPTY_MASK in mask mode
    xu = _unaligned_copy(x)
    tgt = synth_tile(1, 128, kind="target")
    r = sb.ReinhardStainNormalizer()
    r.fit(tgt)
    outs = {}
    for mode in ("ring", "tile"):
        monkeypatch.setenv("SB_REINHARD_TILE_KERNEL", "1" if mode == "tile" else "0")
        f = sb.ReinhardStainNormalizer()
        f.fit(x[B - 1].cpu().numpy())
        outs[mode] = (r.transform(x), r.transform(x, mask_background=True), r.last_status.clone(),
                      sb.LuminosityStandardizer.standardize(x, percentile=93), np.array(f.target_means).ravel(), np.array(f.target_stds).ravel())
    monkeypatch.setenv("SB_REINHARD_TILE_KERNEL", "0")
    for a, b in zip(outs["ring"], outs["tile"]):
        assert (torch.equal(a, b) if isinstance(a, torch.Tensor) else np.array_equal(a, b))
    assert torch.equal(r.transform(xu), outs["ring"][0])
    assert torch.equal(r.transform(xu, mask_background=True), outs["ring"][1])
    assert torch.equal(sb.LuminosityStandardizer.standardize(xu, percentile=93), outs["ring"][3])
    if B > 2:
        assert int(outs["ring"][2][2]) == 1 and int(outs["ring"][2][1]) == 0


def test_concentrations_ring_pass_equals_register_kernel(sb):
    from stainlib_b200.utils.stain_utils import get_concentrations
    x = torch.from_numpy(synth_batch(4200, 3, 256, 320)).cuda()
    M = sb.MacenkoStainExtractor.get_stain_matrix(x)
    a = get_concentrations(x, M)
    b = get_concentrations(_unaligned_copy(x), M)
    assert torch.equal(a, b)
    ref = so.get_concentrations(x[1].cpu().numpy(), M[1].cpu().numpy())
    assert np.abs(a[1].cpu().numpy() - ref).max() < 2e-5


def test_reinhard_tile_bytes_do_not_depend_on_the_batch(sb):
    """SURVEY section 8-e for the streaming Reinhard passes: how the chunks of a tile fall onto CTAs depends on the batch around it;
    the statistics are integer sums and histograms, so a tile's bytes and statistics do not."""
    x = torch.from_numpy(synth_batch(4300, 37, 256, 304)).cuda()
    r = sb.ReinhardStainNormalizer()
    r.fit(synth_tile(1, 128, kind="target"))
    full = r.transform(x, mask_background=True)
    lum = sb.LuminosityStandardizer.standardize(x)
    for lo, hi in ((0, 1), (5, 6), (36, 37), (10, 23)):
        assert torch.equal(r.transform(x[lo:hi].contiguous(), mask_background=True), full[lo:hi])
        assert torch.equal(sb.LuminosityStandardizer.standardize(x[lo:hi].contiguous()), lum[lo:hi])
    from stainlib_b200.utils.stain_utils import get_mean_std
    m_all, s_all = get_mean_std(x)
    m_one, s_one = get_mean_std(x[7:8].contiguous())
    assert torch.equal(m_all[7:8], m_one) and torch.equal(s_all[7:8], s_one)
