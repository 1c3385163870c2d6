"""Mirror of stainlib/extraction/vahadane_stain_extractor.py (VahadaneStainExtractor.get_stain_matrix, lines 19-43).

``spams.trainDL`` (1-second wall-clock budget, random initialisation) is replaced by a deterministic full-batch
sparse-NMF: ``n_iter`` alternations of closed-form sparse coding and one block-coordinate dictionary update, started
from the Ruifrok H/E vectors -- see DESIGN.md for the parity definition."""
from stainlib_b200 import _native as nv
from stainlib_b200.extraction.macenko_stain_extractor import _extract
from stainlib_b200.utils.stain_utils import ABCStainExtractor, is_uint8_image


class VahadaneStainExtractor(ABCStainExtractor):
    last_status = None
    n_iter = 50

    @staticmethod
    def get_stain_matrix(I, luminosity_threshold=0.8, regularizer=0.1, n_iter=None):
        assert is_uint8_image(I), "Image should be RGB uint8."
        p = nv.default_params(nv.SB_METHOD_VAHADANE, luminosity_threshold=float(luminosity_threshold),
                              dl_lambda=float(regularizer),
                              dl_iters=int(VahadaneStainExtractor.n_iter if n_iter is None else n_iter))
        M, st = _extract(I, p)
        VahadaneStainExtractor.last_status = st
        return M


VahadaneExtractor = VahadaneStainExtractor  # north_star spelling
