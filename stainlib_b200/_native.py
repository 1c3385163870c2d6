"""ctypes binding of libstainb200.so (C ABI in include/stainb200.h) plus the tensor plumbing around it.

PyTorch is used for device memory, streams and pinned host buffers only.  There is no CPU fallback: if the shared
library is missing, or a call is made without a CUDA device, this module raises.
"""
import ctypes
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstainb200.so")

SB_STATUS_EMPTY_MASK = 1
SB_STATUS_FEW_TISSUE = 2
SB_STATUS_ZERO_MAXC = 4
SB_STATUS_DEGENERATE = 8
SB_METHOD_MACENKO = 0
SB_METHOD_VAHADANE = 1

EXPORTS = [
    "sb_default_params", "sb_create", "sb_destroy", "sb_error_string", "sb_last_cuda_error", "sb_version",
    "sb_launch_count", "sb_tissue_mask", "sb_extract", "sb_fit", "sb_normalize", "sb_normalize_host",
    "sb_concentrations", "sb_recombine", "sb_stain_augment", "sb_reinhard_stats", "sb_reinhard_transform",
    "sb_luminosity_standardize", "sb_hed_augment", "sb_grayscale_augment",
    "sb_slide_grid", "sb_slide_moments", "sb_slide_angle_hist", "sb_slide_conc_hist", "sb_slide_dl_sums", "sb_decode_jpeg",
    "sb_workspace_bytes", "sb_set_workspace", "sb_stream_fallbacks", "sb_standardize_brightness", "sb_lab_mean_std",
    "sb_lab_split", "sb_lab_merge", "sb_set_pass_timing", "sb_get_pass_timing", "sb_rgb_to_od", "sb_od_to_rgb", "sb_hed_augment_f32",
]


class SbParams(ctypes.Structure):
    _fields_ = [
        ("method", ctypes.c_int),
        ("luminosity_threshold", ctypes.c_double),
        ("angular_percentile", ctypes.c_double),
        ("lasso_lambda", ctypes.c_double),
        ("conc_percentile", ctypes.c_double),
        ("dl_lambda", ctypes.c_double),
        ("dl_iters", ctypes.c_int),
        ("cluster_size", ctypes.c_int),
        ("dl_sample_iters", ctypes.c_int),
        ("dl_anderson", ctypes.c_int),
    ]


class NativeError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()
_handles = {}


def load_library():
    """Loads libstainb200.so (once).  Raises ImportError if it has not been built -- there is no fallback."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m stainlib_b200.build` "
                "(or __graft_entry__.build()).  stainlib_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        pp = ctypes.POINTER(SbParams)
        lib.sb_default_params.argtypes = [pp]
        lib.sb_default_params.restype = None
        lib.sb_create.argtypes = [ci, ctypes.POINTER(vp)]
        lib.sb_destroy.argtypes = [vp]
        lib.sb_error_string.argtypes = [ci]
        lib.sb_error_string.restype = ctypes.c_char_p
        lib.sb_last_cuda_error.restype = ctypes.c_char_p
        lib.sb_version.restype = ci
        lib.sb_launch_count.argtypes = [vp]
        lib.sb_launch_count.restype = ctypes.c_longlong
        lib.sb_tissue_mask.argtypes = [vp, vp, ci, ci, ci, cd, vp, vp, vp]
        lib.sb_extract.argtypes = [vp, vp, ci, ci, ci, pp, vp, vp, vp]
        lib.sb_fit.argtypes = [vp, vp, ci, ci, ci, pp, vp, vp, vp, vp]
        lib.sb_normalize.argtypes = [vp, vp, vp, ci, ci, ci, pp, vp, vp, vp, vp, vp, vp]
        lib.sb_normalize_host.argtypes = [vp, vp, vp, ci, ci, ci, pp, vp, vp, vp, ci]
        lib.sb_concentrations.argtypes = [vp, vp, ci, ci, ci, vp, cd, vp, vp]
        lib.sb_recombine.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp, cd, vp]
        lib.sb_stain_augment.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp, ci, cd, cd, vp]
        lib.sb_reinhard_stats.argtypes = [vp, vp, ci, ci, ci, vp, vp, vp]
        lib.sb_reinhard_transform.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, ci, cd, vp, vp]
        lib.sb_luminosity_standardize.argtypes = [vp, vp, vp, ci, ci, ci, cd, vp]
        lib.sb_hed_augment.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, cd, cd, cd, ci, vp, vp]
        lib.sb_hed_augment_f32.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, cd, cd, cd, ci, vp, vp]
        lib.sb_workspace_bytes.argtypes = [ci, ci, ci]
        lib.sb_workspace_bytes.restype = ctypes.c_size_t
        lib.sb_set_workspace.argtypes = [vp, vp, ctypes.c_size_t]
        lib.sb_stream_fallbacks.argtypes = [vp, ctypes.POINTER(ctypes.c_uint), ci]
        lib.sb_set_pass_timing.argtypes = [vp, ci]
        lib.sb_get_pass_timing.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_float), ctypes.c_char_p, ci]
        lib.sb_standardize_brightness.argtypes = [vp, vp, vp, ci, ci, ci, vp]
        lib.sb_lab_mean_std.argtypes = [vp, vp, ci, ci, ci, vp, vp, vp]
        lib.sb_lab_split.argtypes = [vp, vp, ctypes.c_size_t, vp, vp, vp, vp]
        lib.sb_lab_merge.argtypes = [vp, vp, vp, vp, ci, ctypes.c_size_t, vp, vp]
        lib.sb_rgb_to_od.argtypes = [vp, vp, ctypes.c_size_t, vp, ci, vp]
        lib.sb_od_to_rgb.argtypes = [vp, vp, ci, ctypes.c_size_t, vp, vp, vp]
        lib.sb_grayscale_augment.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
        lib.sb_decode_jpeg.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp]
        lib.sb_slide_grid.argtypes = [vp, ci, ci, ci]
        lib.sb_slide_moments.argtypes = [vp, vp, ci, ci, ci, cd, vp, vp]
        lib.sb_slide_angle_hist.argtypes = [vp, vp, ci, ci, ci, cd, vp, ci, vp, vp, vp]
        lib.sb_slide_conc_hist.argtypes = [vp, vp, ci, ci, ci, vp, cd, ci, vp, vp, vp]
        lib.sb_slide_dl_sums.argtypes = [vp, vp, ci, ci, ci, cd, vp, cd, ci, vp, vp]
        for name in EXPORTS:
            if name not in ("sb_default_params", "sb_error_string", "sb_last_cuda_error", "sb_launch_count", "sb_workspace_bytes"):
                getattr(lib, name).restype = ci
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        lib = load_library()
        msg = lib.sb_error_string(rc).decode()
        if rc == -2:
            msg += ": " + lib.sb_last_cuda_error().decode()
        raise NativeError(f"libstainb200: {msg} (code {rc})")


def default_params(method=SB_METHOD_MACENKO, **overrides):
    p = SbParams()
    load_library().sb_default_params(ctypes.byref(p))
    p.method = method
    for k, v in overrides.items():
        if v is not None:
            setattr(p, k, v)
    return p


def get_handle(device=None):
    """One sb_handle per CUDA device per process."""
    if not torch.cuda.is_available():
        raise NativeError("stainlib_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _handles.get(idx)
    if h is None:
        torch.cuda.init()
        with torch.cuda.device(idx):
            torch.zeros(1, device=f"cuda:{idx}")  # make sure the primary context exists
            hp = ctypes.c_void_p()
            check(load_library().sb_create(idx, ctypes.byref(hp)))
        h = hp
        _handles[idx] = h
    return h, idx


def set_workspace(tensor, device=None):
    """Lends ``tensor`` (a CUDA uint8 tensor, or None to take it back) to the device's handle as per-call scratch
    (include/stainb200.h: sb_set_workspace); ``workspace_bytes(B, H, W)`` says how much a batch needs.  The tensor must
    stay alive while it is lent: the handle cache keeps a reference."""
    h, idx = get_handle(device if tensor is None else tensor.device)
    if tensor is None:
        check(load_library().sb_set_workspace(h, None, 0))
        _workspaces.pop(idx, None)
        return
    assert tensor.is_cuda and tensor.dtype == torch.uint8 and tensor.is_contiguous()
    check(load_library().sb_set_workspace(h, ctypes.c_void_p(tensor.data_ptr()), tensor.numel()))
    _workspaces[idx] = tensor


def stream_fallbacks(device=None, reset=False):
    """[total, by reason 1..7] tiles that left the streaming statistics passes for the fused kernel (diagnostics)."""
    h, _ = get_handle(device)
    out = (ctypes.c_uint * 8)()
    check(load_library().sb_stream_fallbacks(h, out, int(bool(reset))))
    return [int(x) for x in out]


def set_pass_timing(enable, device=None):
    h, _ = get_handle(device)
    check(load_library().sb_set_pass_timing(h, int(bool(enable))))


def get_pass_timing(device=None):
    """[(name, ms), ...] of the statistics passes of the last extract / fit / transform (needs set_pass_timing(True))."""
    h, _ = get_handle(device)
    ms = (ctypes.c_float * 48)()
    names = ctypes.create_string_buffer(4096)
    n = load_library().sb_get_pass_timing(h, 48, ms, names, 4096)
    if n < 0:
        check(n)
    return list(zip(names.value.decode().split("\n")[:n], [float(ms[i]) for i in range(n)]))


def workspace_bytes(B, H, W):
    return int(load_library().sb_workspace_bytes(int(B), int(H), int(W)))


_workspaces = {}


def launch_count(device=None):
    h, _ = get_handle(device)
    return int(load_library().sb_launch_count(h))


def stream_ptr(idx):
    return ctypes.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# ------------------------------------------------------------------------------------------------ tensor plumbing
class Batch:
    """A uint8 [B,H,W,3] CUDA view of whatever the caller passed, remembering how to hand the result back."""

    def __init__(self, I, device=None):
        self.kind = None
        self.single = False
        if isinstance(I, np.ndarray):
            assert I.ndim == 3 and I.dtype == np.uint8, "Image should be RGB uint8."
            if I.shape[2] != 3:
                raise AssertionError("Image should be RGB uint8.")
            self.kind = "numpy"
            self.single = True
            t = torch.from_numpy(np.ascontiguousarray(I))[None]
        elif isinstance(I, torch.Tensor):
            assert I.dtype == torch.uint8 and I.dim() in (3, 4) and I.shape[-1] == 3, "Image should be RGB uint8."
            self.kind = "cuda" if I.is_cuda else "cpu"
            self.single = I.dim() == 3
            t = I[None] if self.single else I
        else:
            raise AssertionError("Image should be RGB uint8.")
        if t.is_cuda:
            self.idx = t.device.index
            self.dev = t.contiguous()
        else:
            _, self.idx = get_handle(device)
            self.host = t.contiguous()
            self.dev = self.host.to(f"cuda:{self.idx}", non_blocking=self.host.is_pinned())
        self.handle, _ = get_handle(self.idx)
        self.B, self.H, self.W = int(self.dev.shape[0]), int(self.dev.shape[1]), int(self.dev.shape[2])

    def give_back(self, t):
        """Device tensor -> the caller's kind (numpy / cpu tensor / cuda tensor), squeezing a single tile."""
        if self.single:
            t = t[0]
        if self.kind == "numpy":
            return t.cpu().numpy()
        if self.kind == "cpu":
            return t.cpu()
        return t

    def new_like(self):
        return torch.empty_like(self.dev)

    def dev_tensor(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.dev.device)
