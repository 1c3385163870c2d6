"""Mirror of stainlib/augmentation/augmenter.py: HED colour augmentation (HedColorAugmenter and its Lighter / Light /
Strong presets, lines 86-372), GrayscaleAugmentor (374-401) and StainAugmentor (403-449) with the reference's
constructor / randomize / transform / fit / pop signatures.

All random draws stay on the host and use numpy's global RNG in the reference's order (sigma_H, sigma_E, sigma_D,
bias_H, bias_E, bias_D for HED; alpha_0, beta_0, alpha_1, beta_1 for StainAugmentor), so a seeded run draws the same
parameters as the reference; the per-pixel work runs in CUDA kernels.
"""
import numpy as np
import torch

from stainlib_b200 import _native as nv
from stainlib_b200.extraction.macenko_stain_extractor import MacenkoStainExtractor
from stainlib_b200.extraction.vahadane_stain_extractor import VahadaneStainExtractor
from stainlib_b200.utils.excepts import InvalidRangeError
from stainlib_b200.utils.stain_utils import LuminosityThresholdTissueLocator, get_concentrations, is_uint8_image

HED_LOG_BASE = 10.0   # scikit-image 0.16-0.17 (pinned 0.17.2 in the reference's environment.yml:107); use np.e for <= 0.15
# Which skimage.color.rgb2hed / hed2rgb the augmenters reproduce (augmenter.py:295,319): "0.17" = the version pinned by
# the reference's environment.yml (-log10(rgb + 2) @ inv(M)); "0.18" = scikit-image >= 0.18, which is what the
# reference's unpinned setup.py installs today (max(0, ln(max(rgb, 1e-6)) / ln(1e-6) @ inv(M)), no "+2").
HED_SKIMAGE = "0.17"


def _hed_variant(v):
    v = str(HED_SKIMAGE if v is None else v)
    if v in ("0.17", "0.16", "17"):
        return 17
    if v in ("0.18", "18") or v.startswith("0.19") or v.startswith("0.2"):
        return 18
    raise ValueError(f"unknown scikit-image variant {v!r}: use '0.17' or '0.18'")


class AugmenterBase(object):
    """Base class for patch augmentation (augmenter.py:19-69)."""

    def __init__(self, keyword):
        super().__init__()
        self._keyword = keyword

    @property
    def keyword(self):
        return self._keyword

    def shapes(self, target_shapes):
        return target_shapes

    def transform(self, patch):
        pass

    def randomize(self):
        pass


class ColorAugmenterBase(AugmenterBase):
    """Base class for color patch augmentation (augmenter.py:71-84)."""

    def __init__(self, keyword):
        super().__init__(keyword=keyword)


def _check_range(title, rng, lo, hi):
    if rng is not None and (len(rng) != 2 or rng[1] < rng[0] or rng[0] < lo or hi < rng[1]):
        raise InvalidRangeError(title, rng)


class HedColorAugmenter(ColorAugmenterBase):
    """Colour correction in HED space: each of the H, E, D channels becomes value * (1 + sigma) + bias
    (augmenter.py:86-344)."""

    _CHANNELS = ("Haematoxylin", "Eosin", "Dab")

    def __init__(self, haematoxylin_sigma_range, haematoxylin_bias_range, eosin_sigma_range, eosin_bias_range,
                 dab_sigma_range, dab_bias_range, cutoff_range):
        super().__init__(keyword="hed_color")
        self._sigma_ranges = None
        self._bias_ranges = None
        self._cutoff_range = None
        self._sigmas = None
        self._biases = None
        self._setsigmaranges(haematoxylin_sigma_range, eosin_sigma_range, dab_sigma_range)
        self._setbiasranges(haematoxylin_bias_range, eosin_bias_range, dab_bias_range)
        self._setcutoffrange(cutoff_range)
        self.last_status = None

    def _setsigmaranges(self, haematoxylin_sigma_range, eosin_sigma_range, dab_sigma_range):
        ranges = [haematoxylin_sigma_range, eosin_sigma_range, dab_sigma_range]
        for name, r in zip(self._CHANNELS, ranges):
            _check_range(f"{name} Sigma", r, -1.0, 1.0)
        self._sigma_ranges = ranges
        # until randomize() is called the lower bound is used (augmenter.py:194-198)
        self._sigmas = [r[0] if r is not None else 0.0 for r in ranges]

    def _setbiasranges(self, haematoxylin_bias_range, eosin_bias_range, dab_bias_range):
        ranges = [haematoxylin_bias_range, eosin_bias_range, dab_bias_range]
        for name, r in zip(self._CHANNELS, ranges):
            _check_range(f"{name} Bias", r, -1.0, 1.0)
        self._bias_ranges = ranges
        self._biases = [r[0] if r is not None else 0.0 for r in ranges]

    def _setcutoffrange(self, cutoff_range):
        _check_range("Cutoff", cutoff_range, 0.0, 1.0)
        self._cutoff_range = cutoff_range if cutoff_range is not None else [0.0, 1.0]

    def randomize(self):
        """Randomize sigma and bias per channel (augmenter.py:333-344); a ``None`` sigma range yields 1.0 as in the
        reference."""
        self._sigmas = [np.random.uniform(low=r[0], high=r[1], size=None) if r is not None else 1.0
                        for r in self._sigma_ranges]
        self._biases = [np.random.uniform(low=r[0], high=r[1], size=None) if r is not None else 0.0
                        for r in self._bias_ranges]

    def transform(self, patch, sigmas=None, biases=None, skimage_version=None):
        """Apply the colour deformation (augmenter.py:276-331).  ``sigmas`` / ``biases`` ([B,3]) override the
        instance's current draw for batched calls with per-tile parameters.  Float patches (values in [0,1]) are
        transformed as they are and come back as floats of the same dtype, like the reference (augmenter.py:288-291,
        323-327).  ``skimage_version``: "0.17" / "0.18" (default: the module's HED_SKIMAGE)."""
        import ctypes
        variant = _hed_variant(skimage_version)
        is_float = (isinstance(patch, np.ndarray) and patch.dtype.kind == "f") or \
                   (isinstance(patch, torch.Tensor) and patch.dtype.is_floating_point)
        lib = nv.load_library()
        if is_float:
            return self._transform_float(patch, sigmas, biases, variant)
        assert is_uint8_image(patch), "Image should be RGB uint8."
        b = nv.Batch(patch)
        par = self._params_on_device(b.B, sigmas, biases, b.dev.device)
        out = b.new_like()
        status = b.dev_tensor((b.B,), torch.int32)
        nv.check(lib.sb_hed_augment(
            b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, nv.ptr(par), ctypes.c_void_p(par.data_ptr() + b.B * 24),
            float(self._cutoff_range[0]), float(self._cutoff_range[1]), float(HED_LOG_BASE), variant, nv.ptr(status),
            nv.stream_ptr(b.idx)))
        self.last_status = status
        if b.single and b.kind == "numpy" and int(status[0]) == 1:
            return patch          # outside the cutoff the reference returns the same object (augmenter.py:329-331)
        return b.give_back(out)

    def _params_on_device(self, B, sigmas, biases, device):
        sg = np.broadcast_to(np.asarray(self._sigmas if sigmas is None else sigmas, dtype=np.float64), (B, 3))
        bs = np.broadcast_to(np.asarray(self._biases if biases is None else biases, dtype=np.float64), (B, 3))
        return torch.as_tensor(np.ascontiguousarray(np.concatenate([sg, bs], axis=0))).to(device)

    def _transform_float(self, patch, sigmas, biases, variant):
        import ctypes
        is_np = isinstance(patch, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(patch)) if is_np else patch
        assert t.dim() in (3, 4) and t.shape[-1] == 3, "Image should be RGB."
        single = t.dim() == 3
        h, idx = nv.get_handle(t.device if t.is_cuda else None)
        dev = t.to(device=f"cuda:{idx}", dtype=torch.float32).contiguous()
        if single:
            dev = dev[None]
        B, H, W = int(dev.shape[0]), int(dev.shape[1]), int(dev.shape[2])
        par = self._params_on_device(B, sigmas, biases, dev.device)
        out = torch.empty_like(dev)
        status = torch.empty(B, dtype=torch.int32, device=dev.device)
        nv.check(nv.load_library().sb_hed_augment_f32(
            h, nv.ptr(dev), nv.ptr(out), B, H, W, nv.ptr(par), ctypes.c_void_p(par.data_ptr() + B * 24),
            float(self._cutoff_range[0]), float(self._cutoff_range[1]), float(HED_LOG_BASE), variant, nv.ptr(status),
            nv.stream_ptr(idx)))
        self.last_status = status
        if single and is_np and int(status[0]) == 1:
            return patch
        out = out[0] if single else out
        out = out.to(t.dtype)
        if is_np:
            return out.cpu().numpy()
        return out if t.is_cuda else out.cpu()


class HedColorAugmenter1(HedColorAugmenter):
    """All six ranges = (-thresh, thresh), cutoff (0.05, 0.95) (augmenter.py:346-360)."""

    def __init__(self, thresh):
        r = (-thresh, thresh)
        super().__init__(r, r, r, r, r, r, (0.05, 0.95))


class HedLighterColorAugmenter(HedColorAugmenter1):
    def __init__(self):
        super().__init__(0.03)


class HedLightColorAugmenter(HedColorAugmenter1):
    def __init__(self):
        super().__init__(0.1)


class HedStrongColorAugmenter(HedColorAugmenter1):
    def __init__(self):
        super().__init__(1.0)


class GrayscaleAugmentor(object):
    """augmenter.py:374-401.  Like the reference, ``pop`` ignores sigma1/sigma2 (hard-coded 0.2) and the mask."""

    def __init__(self, sigma1=0.2, sigma2=0.2, augment_background=False):
        self.sigma1 = sigma1
        self.sigma2 = sigma2
        self.augment_background = augment_background

    def fit(self, I):
        assert is_uint8_image(I), "Image should be RGB uint8."
        self.image_shape = I.shape
        self.tissue_mask = LuminosityThresholdTissueLocator.get_tissue_mask(I)
        self.tissue_mask = self.tissue_mask.ravel() if isinstance(self.tissue_mask, np.ndarray) else self.tissue_mask
        self.image = I

    def pop(self):
        b = nv.Batch(self.image)
        # per tile alpha ~ U(0.8, 1.2) then beta ~ U(-0.2, 0.2): one vectorised call consumes numpy's global stream in
        # exactly the order of the reference's scalar calls (augmenter.py:395-396), tile after tile
        draws = np.random.uniform(low=[1 - 0.2, -0.2], high=[1 + 0.2, 0.2], size=(b.B, 2))
        par = torch.as_tensor(np.ascontiguousarray(draws.T)).to(b.dev.device)   # [2,B]: alphas then betas
        out = b.new_like()
        import ctypes
        nv.check(nv.load_library().sb_grayscale_augment(b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, nv.ptr(par),
                                                        ctypes.c_void_p(par.data_ptr() + b.B * 8), nv.stream_ptr(b.idx)))
        return b.give_back(out)


class StainAugmentor(object):
    """augmenter.py:403-449.  ``fit`` estimates the tile's own stain matrix on the GPU; ``pop`` draws
    (alpha_0, beta_0, alpha_1, beta_1) per tile from numpy's global RNG and runs one fused kernel: closed-form
    concentrations, alpha*C+beta on tissue pixels (all pixels with ``augment_background``), recombination with the
    tile's own matrix, clip to [0,255]."""

    def __init__(self, method, sigma1=0.2, sigma2=0.2, augment_background=False):
        if method.lower() == 'macenko':
            self.extractor = MacenkoStainExtractor
        elif method.lower() == 'vahadane':
            self.extractor = VahadaneStainExtractor
        else:
            raise Exception('Method not recognized.')
        self.sigma1 = sigma1
        self.sigma2 = sigma2
        self.augment_background = augment_background

    def fit(self, I):
        assert is_uint8_image(I), "Image should be RGB uint8."
        self.image_shape = I.shape
        self._batch = nv.Batch(I)
        self.stain_matrix = self.extractor.get_stain_matrix(I)
        self.n_stains = 2
        self._image = I

    @property
    def source_concentrations(self):
        return get_concentrations(self._image, self.stain_matrix)

    @property
    def tissue_mask(self):
        m = LuminosityThresholdTissueLocator.get_tissue_mask(self._image)
        return m.ravel() if isinstance(m, np.ndarray) else m

    def pop(self, alphas=None, betas=None):
        b = self._batch
        if alphas is None:
            # (alpha_0, beta_0, alpha_1, beta_1) per tile, drawn in the reference's order (augmenter.py:435-437) by one
            # vectorised call on numpy's global stream
            d = np.random.uniform(low=[1 - self.sigma1, -self.sigma2], high=[1 + self.sigma1, self.sigma2], size=(b.B, 2, 2))
            alphas, betas = d[:, :, 0], d[:, :, 1]
        al = np.broadcast_to(np.asarray(alphas, dtype=np.float64), (b.B, 2))
        be = np.broadcast_to(np.asarray(betas, dtype=np.float64), (b.B, 2))
        par = torch.as_tensor(np.ascontiguousarray(np.concatenate([al, be], axis=0))).to(b.dev.device)
        M = self.stain_matrix
        if isinstance(M, torch.Tensor) and M.is_cuda:
            Mt = M.to(torch.float64).reshape(-1, 2, 3).contiguous()          # batched fit: already on the device
        else:
            Mt = torch.as_tensor(np.asarray(M), dtype=torch.float64).reshape(-1, 2, 3).contiguous().to(b.dev.device)
        out = b.new_like()
        import ctypes
        nv.check(nv.load_library().sb_stain_augment(
            b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, nv.ptr(Mt), nv.ptr(par),
            ctypes.c_void_p(par.data_ptr() + b.B * 16), int(bool(self.augment_background)), 0.8, 0.01, nv.stream_ptr(b.idx)))
        return b.give_back(out)

    transform = pop   # north_star spelling
