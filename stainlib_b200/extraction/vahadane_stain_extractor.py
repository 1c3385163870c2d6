"""Mirror of stainlib/extraction/vahadane_stain_extractor.py (VahadaneStainExtractor.get_stain_matrix, lines 19-43).

``spams.trainDL`` (1-second wall-clock budget, random initialisation) is replaced by a deterministic full-batch
sparse-NMF started from the Ruifrok H/E vectors: alternations of closed-form sparse coding and one block-coordinate
dictionary update -- ``n_sample_iter`` warm-start passes over a 1-in-16 sample of the tile, then ``n_iter`` passes over
every tissue pixel, the 6-component fixed-point map Anderson-accelerated with memory ``anderson``
(both phases stop early once the residual is small; ``n_sample_iter=0, anderson=0`` with a fixed ``n_iter`` is the plain iteration).  See DESIGN.md for the parity definition."""
from stainlib_b200 import _native as nv
from stainlib_b200.extraction.macenko_stain_extractor import _extract
from stainlib_b200.utils.stain_utils import ABCStainExtractor, is_uint8_image


class VahadaneStainExtractor(ABCStainExtractor):
    last_status = None
    n_iter = 10
    n_sample_iter = 12
    anderson = 4

    @staticmethod
    def get_stain_matrix(I, luminosity_threshold=0.8, regularizer=0.1, n_iter=None, n_sample_iter=None, anderson=None, cluster_size=None):
        assert is_uint8_image(I), "Image should be RGB uint8."
        cls = VahadaneStainExtractor
        p = nv.default_params(nv.SB_METHOD_VAHADANE, luminosity_threshold=float(luminosity_threshold),
                              dl_lambda=float(regularizer),
                              dl_iters=int(cls.n_iter if n_iter is None else n_iter),
                              dl_sample_iters=int(cls.n_sample_iter if n_sample_iter is None else n_sample_iter),
                              dl_anderson=int(cls.anderson if anderson is None else anderson), cluster_size=cluster_size)
        M, st = _extract(I, p)
        VahadaneStainExtractor.last_status = st
        return M


VahadaneExtractor = VahadaneStainExtractor  # north_star spelling
