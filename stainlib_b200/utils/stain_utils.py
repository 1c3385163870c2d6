"""Host-side mirror of stainlib/utils/stain_utils.py: same names and argument meaning, CUDA kernels underneath.

Every function accepts what the reference accepts (one ``np.uint8 [H,W,3]`` image) and returns what the reference
returns; as an extension it also accepts ``torch.uint8 [B,H,W,3]`` batches (CPU/pinned or CUDA) and then returns a
tensor of the same kind with a leading batch dimension.
"""
import ctypes
from abc import ABC, abstractmethod

import numpy as np
import torch

from stainlib_b200 import _native as nv
from stainlib_b200.utils.excepts import TissueMaskException


class ABCStainExtractor(ABC):
    """stain_utils.py:8-17."""

    @staticmethod
    @abstractmethod
    def get_stain_matrix(I):
        """Estimate the stain matrix given an image."""


class ABCTissueLocator(ABC):
    """stain_utils.py:19-27."""

    @staticmethod
    @abstractmethod
    def get_tissue_mask(I):
        """Get a boolean tissue mask."""


def is_image(I):
    """stain_utils.py:126-134 (numpy) -- extended to torch tensors."""
    if isinstance(I, torch.Tensor):
        return I.dim() in (3, 4)
    if not isinstance(I, np.ndarray):
        return False
    return I.ndim == 3


def is_uint8_image(I):
    """stain_utils.py:136-144.  Unlike the reference, a channel count other than 3 is rejected."""
    if not is_image(I):
        return False
    if isinstance(I, torch.Tensor):
        return I.dtype == torch.uint8 and I.shape[-1] == 3
    return I.dtype == np.uint8 and I.shape[2] == 3


def raise_for_status(status, single):
    """Reference error behaviour for single-tile calls (stain_utils.py:46-47; np.linalg.eigh on a NaN covariance)."""
    if not single:
        return
    s = int(status[0])
    if s & nv.SB_STATUS_EMPTY_MASK:
        raise TissueMaskException("Empty tissue mask computed")
    if s & (nv.SB_STATUS_FEW_TISSUE | nv.SB_STATUS_DEGENERATE):
        raise np.linalg.LinAlgError("Eigenvalues did not converge")


class LuminosityThresholdTissueLocator(ABCTissueLocator):
    """stain_utils.py:29-48."""

    @staticmethod
    def get_tissue_mask(I, luminosity_threshold=0.8):
        assert is_uint8_image(I), "Image should be RGB uint8."
        b = nv.Batch(I)
        mask = b.dev_tensor((b.B, b.H, b.W), torch.bool)       # the kernel writes 0 / 1 bytes: a bool tensor, no conversion pass
        status = b.dev_tensor((b.B,), torch.int32)
        nv.check(nv.load_library().sb_tissue_mask(b.handle, nv.ptr(b.dev), b.B, b.H, b.W, float(luminosity_threshold),
                                                  nv.ptr(mask), nv.ptr(status), nv.stream_ptr(b.idx)))
        st = status.cpu()
        LuminosityThresholdTissueLocator.last_status = st
        raise_for_status(st, b.single)
        return b.give_back(mask)


class LuminosityStandardizer(object):
    """stain_utils.py:50-67."""

    @staticmethod
    def standardize(I, percentile=95):
        assert is_uint8_image(I), "Image should be RGB uint8."
        b = nv.Batch(I)
        out = b.new_like()
        nv.check(nv.load_library().sb_luminosity_standardize(b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W,
                                                             float(percentile), nv.stream_ptr(b.idx)))
        return b.give_back(out)


def _matrices_to_device(M, b):
    """[2,3] or [B,2,3] (numpy / tensor) -> float64 CUDA tensor [B,2,3]."""
    Mt = torch.as_tensor(np.asarray(M.detach().cpu()) if isinstance(M, torch.Tensor) else np.asarray(M), dtype=torch.float64)
    if Mt.dim() == 2:
        Mt = Mt[None].expand(b.B, 2, 3)
    return Mt.contiguous().to(b.dev.device)


def get_concentrations(I, stain_matrix, regularizer=0.01):
    """stain_utils.py:69-78.  Returns N x 2 (float64 for numpy input, as the reference; float32 tensors for batches:
    [B,N,2])."""
    b = nv.Batch(I)
    M = _matrices_to_device(stain_matrix, b)
    C = b.dev_tensor((b.B, b.H * b.W, 2), torch.float32)
    nv.check(nv.load_library().sb_concentrations(b.handle, nv.ptr(b.dev), b.B, b.H, b.W, nv.ptr(M), float(regularizer),
                                                 nv.ptr(C), nv.stream_ptr(b.idx)))
    out = b.give_back(C)
    return out.astype(np.float64) if isinstance(out, np.ndarray) else out


def get_sign(x):
    """stain_utils.py:80-91."""
    if x > 0:
        return +1
    elif x < 0:
        return -1
    elif x == 0:
        return 0


def normalize_matrix_rows(A):
    """stain_utils.py:93-99."""
    if isinstance(A, torch.Tensor):
        return A / torch.linalg.norm(A, dim=-1, keepdim=True)
    return A / np.linalg.norm(A, axis=1)[:, None]


def _to_cuda(x, dtype=None):
    """numpy / CPU tensor / CUDA tensor -> (contiguous CUDA tensor, kind) for the flat element-wise entry points."""
    if isinstance(x, np.ndarray):
        kind, t = "numpy", torch.from_numpy(np.ascontiguousarray(x))
    elif isinstance(x, torch.Tensor):
        kind, t = ("cuda" if x.is_cuda else "cpu"), x
    else:
        raise AssertionError("expected a numpy array or a torch tensor")
    _, idx = nv.get_handle(t.device if t.is_cuda else None)
    t = t.to(f"cuda:{idx}")
    if dtype is not None:
        t = t.to(dtype)
    return t.contiguous(), kind, idx


def _from_cuda(t, kind):
    if kind == "numpy":
        return t.cpu().numpy()
    return t.cpu() if kind == "cpu" else t


def convert_RGB_to_OD(I, dtype=None):
    """stain_utils.py:101-112: OD = max(-ln(max(I,1)/255), 1e-6), a 256-entry float64 table lookup on the GPU
    (``sb_rgb_to_od``).  Returns float64 like the reference (``dtype=torch.float32`` halves the bytes written for
    tensor batches); the kernels of the hot path apply the same table inline and never materialise an OD image."""
    t, kind, idx = _to_cuda(I)
    assert t.dtype == torch.uint8, "Image should be RGB uint8."
    f32 = dtype in (torch.float32, np.float32)
    out = torch.empty(t.shape, dtype=torch.float32 if f32 else torch.float64, device=t.device)
    if t.numel():
        h, _ = nv.get_handle(idx)
        nv.check(nv.load_library().sb_rgb_to_od(h, nv.ptr(t), t.numel(), nv.ptr(out), int(f32), nv.stream_ptr(idx)))
    return _from_cuda(out, kind)


def convert_OD_to_RGB(OD):
    """stain_utils.py:114-124: uint8(255 * exp(-max(OD, 1e-6))) on the GPU (``sb_od_to_rgb``); asserts OD >= 0 like the
    reference."""
    t, kind, idx = _to_cuda(OD)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    neg = torch.zeros(1, dtype=torch.int32, device=t.device)
    if t.numel():
        h, _ = nv.get_handle(idx)
        nv.check(nv.load_library().sb_od_to_rgb(h, nv.ptr(t), int(t.dtype == torch.float32), t.numel(), nv.ptr(out), nv.ptr(neg),
                                                nv.stream_ptr(idx)))
    assert int(neg.item()) == 0, "Negative optical density."
    return _from_cuda(out, kind)


def standardize_brightness(I):
    """stain_utils.py:188-194: uint8(clip(I * 255 / percentile(I, 90), 0, 255)) per tile (exact 256-bin histogram
    percentile on the GPU)."""
    assert is_uint8_image(I), "Image should be RGB uint8."
    b = nv.Batch(I)
    out = b.new_like()
    nv.check(nv.load_library().sb_standardize_brightness(b.handle, nv.ptr(b.dev), nv.ptr(out), b.B, b.H, b.W, nv.stream_ptr(b.idx)))
    return b.give_back(out)


def lab_split(I):
    """stain_utils.py:146-158: 8-bit RGB -> LAB (OpenCV's integer path, bit-exact), split into three float32 planes
    I1 = L / 2.55, I2 = a - 128, I3 = b - 128.  One image -> three [H,W] arrays; a batch -> three [B,H,W] tensors."""
    assert is_uint8_image(I), "Image should be RGB uint8."
    b = nv.Batch(I)
    planes = [b.dev_tensor((b.B, b.H, b.W), torch.float32) for _ in range(3)]
    nv.check(nv.load_library().sb_lab_split(b.handle, nv.ptr(b.dev), b.B * b.H * b.W, nv.ptr(planes[0]), nv.ptr(planes[1]),
                                            nv.ptr(planes[2]), nv.stream_ptr(b.idx)))
    return tuple(b.give_back(p) for p in planes)


def merge_back(I1, I2, I3):
    """stain_utils.py:160-172: LAB planes -> RGB uint8.  Like the reference, numpy planes are scaled IN PLACE
    (I1 *= 2.55, I2 += 128, I3 += 128) as a side effect."""
    is_np = isinstance(I1, np.ndarray)
    t = [torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x for x in (I1, I2, I3)]
    _, idx = nv.get_handle(t[0].device if t[0].is_cuda else None)
    f64 = t[0].dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    d = [x.to(device=f"cuda:{idx}", dtype=dt).contiguous() for x in t]
    out = torch.empty(tuple(d[0].shape) + (3,), dtype=torch.uint8, device=d[0].device)
    h, _ = nv.get_handle(idx)
    nv.check(nv.load_library().sb_lab_merge(h, nv.ptr(d[0]), nv.ptr(d[1]), nv.ptr(d[2]), int(f64), d[0].numel(), nv.ptr(out),
                                            nv.stream_ptr(idx)))
    if is_np:
        I1 *= 2.55
        I2 += 128.0
        I3 += 128.0
        return out.cpu().numpy()
    return out if t[0].is_cuda else out.cpu()


def get_mean_std(I):
    """stain_utils.py:174-186: per-channel mean and population standard deviation of lab_split(I), as the reference's
    tuples of (1,1) float64 arrays (cv.meanStdDev's output); a batch returns two [B,3] tensors."""
    assert is_uint8_image(I), "Image should be RGB uint8."
    b = nv.Batch(I)
    means = b.dev_tensor((b.B, 3), torch.float64)
    stds = b.dev_tensor((b.B, 3), torch.float64)
    nv.check(nv.load_library().sb_lab_mean_std(b.handle, nv.ptr(b.dev), b.B, b.H, b.W, nv.ptr(means), nv.ptr(stds), nv.stream_ptr(b.idx)))
    if b.single:
        m, sd = means[0].cpu().numpy(), stds[0].cpu().numpy()
        return tuple(np.array([[m[k]]]) for k in range(3)), tuple(np.array([[sd[k]]]) for k in range(3))
    return (means, stds) if b.kind == "cuda" else (means.cpu(), stds.cpu())
