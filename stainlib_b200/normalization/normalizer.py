"""Mirror of stainlib/normalization/normalizer.py: ExtractiveStainNormalizer (lines 16-50) and
ReinhardStainNormalizer (lines 54-94) with the reference's constructor / fit / transform signatures and attribute
names, running on sm_100a kernels through libstainb200.so.

Extensions over the reference (which takes one numpy image at a time):
  * ``transform`` accepts ``torch.uint8 [B,H,W,3]`` batches.  CUDA batches are normalised in place on the device by one
    persistent fused kernel; CPU (ideally pinned) batches are streamed H2D -> kernel -> D2H in overlapped chunks by
    the C library (``sb_normalize_host``).
  * ``fit`` under ``torch.distributed``: the fitted target statistics (2x3 stain matrix + 1x2 maxC, or Reinhard's
    3 means + 3 stds) are shared with ONE all-reduce(SUM): ``src`` rank contributes them, the others contribute zeros.
  * per-tile ``last_status`` instead of raising mid-batch; single-image calls raise like the reference.
"""
import ctypes

import numpy as np
import torch

from stainlib_b200 import _native as nv
from stainlib_b200.extraction.macenko_stain_extractor import MacenkoStainExtractor
from stainlib_b200.extraction.vahadane_stain_extractor import VahadaneStainExtractor
from stainlib_b200.utils.stain_utils import get_concentrations, is_uint8_image, raise_for_status


from stainlib_b200.distributed import share_fit_statistics as _share_fit_statistics


class ExtractiveStainNormalizer(object):
    def __init__(self, method, **extractor_kwargs):
        if method.lower() == 'macenko':
            self.extractor = MacenkoStainExtractor
            self._method = nv.SB_METHOD_MACENKO
        elif method.lower() == 'vahadane':
            self.extractor = VahadaneStainExtractor
            self._method = nv.SB_METHOD_VAHADANE
        else:
            raise Exception('Method not recognized.')
        self._kw = extractor_kwargs          # e.g. dl_iters=..., cluster_size=...
        self.last_status = None
        self._target = None

    def _params(self):
        kw = dict(self._kw)
        if self._method == nv.SB_METHOD_VAHADANE:
            kw.setdefault("dl_iters", VahadaneStainExtractor.n_iter)
            kw.setdefault("dl_sample_iters", VahadaneStainExtractor.n_sample_iter)
            kw.setdefault("dl_anderson", VahadaneStainExtractor.anderson)
        return nv.default_params(self._method, **kw)

    def _fit_local(self, target):
        """Stain matrix + maxC of one target tile on this rank's GPU -> float64 vector [M(6), maxC(2)]."""
        assert is_uint8_image(target), "Image should be RGB uint8."
        b = nv.Batch(target)
        assert b.B == 1, "fit() takes one target tile"
        M = b.dev_tensor((1, 2, 3), torch.float64)
        maxC = b.dev_tensor((1, 2), torch.float64)
        status = b.dev_tensor((1,), torch.int32)
        p = self._params()
        nv.check(nv.load_library().sb_fit(b.handle, nv.ptr(b.dev), 1, b.H, b.W, ctypes.byref(p), nv.ptr(M), nv.ptr(maxC),
                                          nv.ptr(status), nv.stream_ptr(b.idx)))
        st = status.cpu()
        self.last_status = st
        raise_for_status(st, True)
        self._target = target
        return torch.cat([M.reshape(6), maxC.reshape(2)]).cpu()

    def fit(self, target, src=0, group=None, slide=None):
        """Fit to a target image (normalizer.py:27-36): stain matrix of the target and the 99th percentile of its
        concentrations, one fused kernel.  Under torch.distributed only rank ``src`` needs a real target.

        Slide-level fit (``slide=True``, or a batch of more than one tile): ``target`` is this rank's share
        ``uint8 [T,H,W,3]`` of the tiles of ONE target slide (``None`` / ``T = 0`` allowed on a rank); the result is what
        the reference would compute on the concatenation of all ranks' tiles -- see normalization/slide_fit.py.  Every
        rank of ``group`` must make the call with ``slide=True``; ``src`` is ignored."""
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
        if slide is None:
            slide = (isinstance(target, (torch.Tensor, np.ndarray)) and target.ndim == 4 and target.shape[0] != 1)
        if slide:
            from stainlib_b200.normalization.slide_fit import macenko_slide_fit, vahadane_slide_fit
            tiles = None
            if isinstance(target, np.ndarray):
                target = torch.from_numpy(np.ascontiguousarray(target))
            if target is not None and target.shape[0] > 0:
                assert target.dim() == 4 and is_uint8_image(target), "Image should be RGB uint8."
                tiles = nv.Batch(target).dev
            p = self._params()
            if self._method == nv.SB_METHOD_MACENKO:
                self.stain_matrix_target, self.maxC_target = macenko_slide_fit(
                    tiles, p.luminosity_threshold, p.angular_percentile, p.lasso_lambda, p.conc_percentile, group=group)
            else:
                self.stain_matrix_target, self.maxC_target = vahadane_slide_fit(
                    tiles, p.luminosity_threshold, p.dl_lambda, p.dl_iters, p.dl_sample_iters, p.dl_anderson, p.lasso_lambda,
                    p.conc_percentile, group=group)
            self._target = None
            return
        vec = torch.zeros(8, dtype=torch.float64)
        if not distributed or dist.get_rank(group) == src:
            vec = self._fit_local(target)
        vec = _share_fit_statistics(vec, src, group)
        self.stain_matrix_target = vec[:6].reshape(2, 3).numpy().copy()
        self.maxC_target = vec[6:].reshape(1, 2).numpy().copy()

    def _target_on_device(self, device):
        """float64 [8] = stain_matrix_target (6) + maxC_target (2) on ``device``; uploaded once per fit (the attributes
        may also be assigned by hand, so the cache is keyed on their current values)."""
        vec = np.concatenate([np.asarray(self.stain_matrix_target, dtype=np.float64).reshape(6),
                              np.asarray(self.maxC_target, dtype=np.float64).reshape(2)])
        hit = getattr(self, "_tgt_cache", None)
        if hit is None or hit[0] != device or not np.array_equal(hit[1], vec):
            hit = (device, vec, torch.as_tensor(vec).to(device))
            self._tgt_cache = hit
        return hit[2]

    @property
    def target_concentrations(self):
        """normalizer.py:35 keeps the N x 2 target concentrations; nothing reads them, so they are computed on demand."""
        if self._target is None:
            return None
        return get_concentrations(self._target, self.stain_matrix_target)

    def transform(self, I, chunk_tiles=0, out=None):
        """Transform an image (normalizer.py:39-50).  Output is not clipped: like the reference's astype(np.uint8) it
        wraps modulo 256.  ``out`` (host batches only): a preallocated, ideally pinned, uint8 tensor of I's shape to
        receive the result, so that a streaming caller does not allocate pinned memory per call."""
        assert is_uint8_image(I), "Image should be RGB uint8."
        lib = nv.load_library()
        p = self._params()
        if isinstance(I, torch.Tensor) and not I.is_cuda and I.dim() == 4:
            # host batch: chunked, overlapped H2D / kernel / D2H inside the C library
            h, idx = nv.get_handle()
            src = I.contiguous()
            if out is None:
                out = torch.empty_like(src, pin_memory=src.is_pinned())
            assert out.shape == src.shape and out.dtype == torch.uint8 and not out.is_cuda and out.is_contiguous()
            status = torch.empty(src.shape[0], dtype=torch.int32)
            Mt = np.ascontiguousarray(self.stain_matrix_target, dtype=np.float64)
            Ct = np.ascontiguousarray(self.maxC_target, dtype=np.float64)
            nv.check(lib.sb_normalize_host(h, nv.ptr(src), nv.ptr(out), int(src.shape[0]), int(src.shape[1]),
                                           int(src.shape[2]), ctypes.byref(p), Mt.ctypes.data_as(ctypes.c_void_p),
                                           Ct.ctypes.data_as(ctypes.c_void_p), nv.ptr(status), int(chunk_tiles)))
            self.last_status = status
            return out
        b = nv.Batch(I)
        out = b.new_like()
        tgt = self._target_on_device(b.dev.device)
        status = b.dev_tensor((b.B,), torch.int32)
        nv.check(lib.sb_normalize(b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, ctypes.byref(p), nv.ptr(tgt),
                                  ctypes.c_void_p(tgt.data_ptr() + 48), None, None, nv.ptr(status), nv.stream_ptr(b.idx)))
        if b.single or b.kind != "cuda":
            st = status.cpu()
            self.last_status = st
            raise_for_status(st, b.single)
        else:
            self.last_status = status        # stays on the device: no sync on the hot path
        return b.give_back(out)


class ReinhardStainNormalizer(object):
    """Normalize a patch stain to the target image using the method of E. Reinhard et al., 'Color transfer between
    images' (normalizer.py:54-94): brightness standardisation, 8-bit LAB split, per-channel mean/std matching."""

    def __init__(self, target_means=0, target_stds=0):
        self.target_means = target_means
        self.target_stds = target_stds
        self.last_status = None

    def fit(self, target, src=0, group=None):
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
        vec = torch.zeros(6, dtype=torch.float64)
        if not distributed or dist.get_rank(group) == src:
            assert is_uint8_image(target), "Image should be RGB uint8."
            b = nv.Batch(target)
            assert b.B == 1, "fit() takes one target tile"
            means = b.dev_tensor((1, 3), torch.float64)
            stds = b.dev_tensor((1, 3), torch.float64)
            nv.check(nv.load_library().sb_reinhard_stats(b.handle, nv.ptr(b.dev), 1, b.H, b.W, nv.ptr(means), nv.ptr(stds),
                                                         nv.stream_ptr(b.idx)))
            vec = torch.cat([means.reshape(3), stds.reshape(3)]).cpu()
        vec = _share_fit_statistics(vec, src, group).numpy()
        # the reference keeps tuples of (1,1) float64 arrays (cv.meanStdDev output)
        self.target_means = tuple(np.array([[vec[k]]]) for k in range(3))
        self.target_stds = tuple(np.array([[vec[3 + k]]]) for k in range(3))

    def transform(self, I, mask_background=False, luminosity_threshold=0.8):
        assert is_uint8_image(I), "Image should be RGB uint8."
        b = nv.Batch(I)
        out = b.new_like()
        t = torch.as_tensor(np.concatenate([np.asarray(self.target_means, dtype=np.float64).reshape(3),
                                            np.asarray(self.target_stds, dtype=np.float64).reshape(3)])).to(b.dev.device)
        status = b.dev_tensor((b.B,), torch.int32)
        nv.check(nv.load_library().sb_reinhard_transform(
            b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, nv.ptr(t), ctypes.c_void_p(t.data_ptr() + 24),
            int(bool(mask_background)), float(luminosity_threshold), nv.ptr(status), nv.stream_ptr(b.idx)))
        if mask_background and (b.single or b.kind != "cuda"):
            st = status.cpu()
            self.last_status = st
            raise_for_status(st, b.single)
        else:
            self.last_status = status
        return b.give_back(out)


class MacenkoNormalizer(ExtractiveStainNormalizer):   # north_star spellings
    def __init__(self, **kw):
        super().__init__('macenko', **kw)


class VahadaneNormalizer(ExtractiveStainNormalizer):
    def __init__(self, **kw):
        super().__init__('vahadane', **kw)


ReinhardNormalizer = ReinhardStainNormalizer
