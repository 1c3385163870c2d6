"""Mirror of stainlib/extraction/macenko_stain_extractor.py (MacenkoStainExtractor.get_stain_matrix, lines 7-44)."""
import torch

from stainlib_b200 import _native as nv
from stainlib_b200.utils.stain_utils import ABCStainExtractor, is_uint8_image, raise_for_status


def _extract(I, params):
    b = nv.Batch(I)
    M = b.dev_tensor((b.B, 2, 3), torch.float64)
    status = b.dev_tensor((b.B,), torch.int32)
    import ctypes
    nv.check(nv.load_library().sb_extract(b.handle, nv.ptr(b.dev), b.B, b.H, b.W, ctypes.byref(params), nv.ptr(M),
                                          nv.ptr(status), nv.stream_ptr(b.idx)))
    st = status.cpu()
    raise_for_status(st, b.single)
    return b.give_back(M), st


class MacenkoStainExtractor(ABCStainExtractor):
    last_status = None

    @staticmethod
    def get_stain_matrix(I, luminosity_threshold=0.8, angular_percentile=99, cluster_size=None):
        """Stain matrix estimation via the method of M. Macenko et al. -- one fused CUDA pass sequence per tile
        (mask + OD moments, fp64 eigenvectors, exact angular percentiles).  2x3 float64 (numpy image) or [B,2,3].
        ``cluster_size`` (extension; 1/2/4/8 CTAs per tile, default automatic) changes the schedule, never the result."""
        assert is_uint8_image(I), "Image should be RGB uint8."
        p = nv.default_params(nv.SB_METHOD_MACENKO, luminosity_threshold=float(luminosity_threshold),
                              angular_percentile=float(angular_percentile), cluster_size=cluster_size)
        M, st = _extract(I, p)
        MacenkoStainExtractor.last_status = st
        return M


MacenkoExtractor = MacenkoStainExtractor  # north_star spelling
