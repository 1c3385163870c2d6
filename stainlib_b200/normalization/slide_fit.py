"""Slide-level (multi-tile) Macenko and Vahadane fit -- SURVEY section 8-f rank 2.

``ExtractiveStainNormalizer.fit`` (normalizer.py:27-36) fits ONE target image.  Whole-slide pipelines fit one stain
matrix per slide from many tiles; this module computes exactly what the reference would return for the tiles
concatenated into one image, without ever concatenating them:

    pass 0  masked OD moments of every tile            -> all-reduce(10 doubles) -> covariance, eigenvectors (host)
    pass 1  4096-bin histogram of the angle keys       -> all-reduce            -> the bins holding the 4 target ranks
    pass 2  2048-bin refinement inside those bins      -> all-reduce            -> exact order statistics -> stain matrix
    pass 3  2 x 4096-bin concentration histograms      -> all-reduce            -> bins of the 99th percentiles
    pass 4  refinement                                 -> all-reduce            -> exact maxC

Every statistic is a sum over tiles, so the tiles may be sharded over the ranks of a ``torch.distributed`` group in
any way (a rank may even hold none): five small all-reduces (80 B to 64 KB) over NCCL / NVLink replace the single
8-double all-reduce of the one-tile fit.  The per-pixel work is the slide_pass kernels of csrc/sb_pipeline.cu; the
host steps below are O(1) and mirror thread 0 of the fused tile kernel (macenko_stain_extractor.py:22-44).
"""
import ctypes
import math

import numpy as np
import torch

from stainlib_b200 import _native as nv
from stainlib_b200.utils.excepts import TissueMaskException

KEY_BITS, L1_BITS, L2_BITS = 23, 12, 11        # csrc/sb_device.cuh
CONC_KEY_K = 2.0


def angle_from_key(key):
    """csrc/sb_device.cuh: angle_from_key -- the angle whose monotone 23-bit "diamond" key is ``key``."""
    d = ((1.0 + key / 8388608.0) - 1.5) * 4.0
    if d > 1.0:
        return math.atan2(2.0 - d, -(d - 1.0))
    if d < -1.0:
        return math.atan2(-2.0 - d, -(-1.0 - d))
    return math.atan2(d, 1.0 - abs(d))


def conc_from_key(key):
    """csrc/sb_device.cuh: conc_from_key."""
    if key == 0:
        return 0.0
    t = 1.0 + key / 8388608.0
    return CONC_KEY_K * (t - 1.0) / (2.0 - t)


def percentile_index(n, pct):
    """numpy.percentile (linear): the two neighbouring ranks and the interpolation weight."""
    vi = (n - 1) * (pct / 100.0)
    lo = min(max(int(math.floor(vi)), 0), n - 1)
    return lo, min(lo + 1, n - 1), vi - math.floor(vi)


def lerp_np(a, b, t):
    d = b - a
    return b - d * (1.0 - t) if t >= 0.5 else a + d * t


def _all_reduce(t, group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:                                           # gloo (CPU tests of the host logic)
            c = t.cpu()
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
            t.copy_(c)
    return t


FIX_MOMENT, FIX_DL = float(1 << 32), float(1 << 30)     # csrc/sb_pipe_common.cuh


def _reduce_sums(t, group, scale):
    """All-reduce of a sums vector whose first nine entries are sums and the rest counts.  The device passes return it as
    int64 with the sums in fixed point: integer addition is exact and order-free, so sharded == unsharded to the last
    bit; float64 vectors (the CPU stand-ins of the tests) are reduced as they are."""
    a = _all_reduce(t, group).cpu().numpy()
    if a.dtype == np.int64:
        out = a.astype(np.float64)
        out[:9] /= scale
        return out
    return a


def _locate(hist, rank):
    """Bin of a cumulative histogram that holds 0-based ``rank`` and the rank inside that bin."""
    cum = np.cumsum(hist)
    b = int(np.searchsorted(cum, rank, side="right"))
    return b, int(rank - (cum[b - 1] if b > 0 else 0))


class SlidePasses(object):
    """The five device passes over this rank's tiles (uint8 [T,H,W,3] CUDA tensor, T may be 0)."""

    def __init__(self, tiles, luminosity_threshold, lasso_lambda, device=None):
        self.lib = nv.load_library()
        self.h, self.idx = nv.get_handle(device if tiles is None else tiles.device)
        self.dev = torch.device("cuda", self.idx)
        self.tiles = tiles if tiles is not None and tiles.shape[0] > 0 else None
        self.thr, self.lam = float(luminosity_threshold), float(lasso_lambda)

    def _shape(self):
        t = self.tiles
        return int(t.shape[0]), int(t.shape[1]), int(t.shape[2])

    def n_pixels(self):
        return 0 if self.tiles is None else int(self.tiles.shape[0] * self.tiles.shape[1] * self.tiles.shape[2])

    def moments(self):
        """int64 [10]: fixed-point (2^32) sums of od (3) and od x od (6) over the tissue pixels, and their count."""
        out = torch.zeros(10, dtype=torch.int64, device=self.dev)
        if self.tiles is not None:
            T, H, W = self._shape()
            grid = self.lib.sb_slide_grid(self.h, T, H, W)
            part = torch.zeros(grid, 10, dtype=torch.int64, device=self.dev)
            nv.check(self.lib.sb_slide_moments(self.h, nv.ptr(self.tiles), T, H, W, self.thr, nv.ptr(part), nv.stream_ptr(self.idx)))
            out = part.sum(dim=0)
        return out

    def dl_sums(self, D, lam, sample):
        """One dictionary pass under D (2x3, rows = atoms): int64 [10] = fixed-point (2^30) A00, A01, A11, B[:,0] (3),
        B[:,1] (3), and the pixel count."""
        out = torch.zeros(10, dtype=torch.int64, device=self.dev)
        if self.tiles is not None:
            T, H, W = self._shape()
            grid = self.lib.sb_slide_grid(self.h, T, H, W)
            part = torch.zeros(grid, 10, dtype=torch.int64, device=self.dev)
            Dc = (ctypes.c_double * 6)(*[float(x) for x in np.asarray(D).reshape(6)])
            nv.check(self.lib.sb_slide_dl_sums(self.h, nv.ptr(self.tiles), T, H, W, self.thr, Dc, float(lam), int(bool(sample)),
                                               nv.ptr(part), nv.stream_ptr(self.idx)))
            out = part.sum(dim=0)
        return out

    def _hist(self, fn, *args):
        hist = torch.zeros(8192, dtype=torch.int64, device=self.dev)
        if self.tiles is not None:
            T, H, W = self._shape()
            nv.check(fn(self.h, nv.ptr(self.tiles), T, H, W, *args, nv.ptr(hist), nv.stream_ptr(self.idx)))
        return hist

    def angle_hist(self, V, level, bins=None):
        Vc = (ctypes.c_double * 6)(*[float(x) for x in V])
        bc = (ctypes.c_uint * 4)(*[int(b) for b in (bins if bins is not None else (0, 0, 0, 0))])
        return self._hist(self.lib.sb_slide_angle_hist, self.thr, Vc, int(level), bc)

    def conc_hist(self, M, level, bins=None):
        Mc = (ctypes.c_double * 6)(*[float(x) for x in np.asarray(M).reshape(6)])
        bc = (ctypes.c_uint * 4)(*[int(b) for b in (bins if bins is not None else (0, 0, 0, 0))])
        return self._hist(self.lib.sb_slide_conc_hist, Mc, self.lam, int(level), bc)


def macenko_slide_fit(tiles, luminosity_threshold=0.8, angular_percentile=99.0, lasso_lambda=0.01, conc_percentile=99.0,
                      group=None, device=None, passes=None):
    """Stain matrix (2x3) and maxC (1x2) of the union of ``tiles`` over all ranks of ``group``.

    tiles: uint8 [T,H,W,3] CUDA tensor holding this rank's share of the slide (None or T = 0 for a rank without tiles).
    Every rank returns the same numbers.  Raises TissueMaskException when the whole slide has no tissue.
    ``passes``: object with the interface of SlidePasses (the CPU tests of this host logic inject one)."""
    sp = passes if passes is not None else SlidePasses(tiles, luminosity_threshold, lasso_lambda, device)
    # ---- pass 0: moments -> covariance (ddof = 1) -> the two leading eigenvectors, signs as macenko_stain_extractor.py:24-27
    mom = sp.moments()
    n_px = torch.tensor([sp.n_pixels()], dtype=mom.dtype, device=sp.dev)
    t = _reduce_sums(torch.cat([mom.to(sp.dev), n_px]), group, FIX_MOMENT)
    n, n_all = t[9], int(round(t[10]))
    if n < 1.0:
        raise TissueMaskException("Empty tissue mask computed")
    if n < 2.0:
        raise np.linalg.LinAlgError("Eigenvalues did not converge")       # np.cov of one sample is NaN (reference behaviour)
    s = t[0:3]
    S = np.array([[t[3], t[4], t[5]], [t[4], t[6], t[7]], [t[5], t[7], t[8]]])
    cov = (S - np.outer(s, s) / n) / (n - 1.0)
    _, vec = np.linalg.eigh(cov)
    vec = vec[:, [2, 1]]
    if vec[0, 0] < 0:
        vec[:, 0] *= -1
    if vec[0, 1] < 0:
        vec[:, 1] *= -1
    V = np.concatenate([vec[:, 0], vec[:, 1]])                               # rows = eigenvectors, as the kernels want them
    # ---- passes 1 + 2: exact 1st / 99th angular percentiles over all tissue pixels
    nt = int(round(n))
    lo0, hi0, f0 = percentile_index(nt, 100.0 - angular_percentile)
    lo1, hi1, f1 = percentile_index(nt, angular_percentile)
    ranks = [lo0, hi0, lo1, hi1]
    h1 = _all_reduce(sp.angle_hist(V, 1), group).cpu().numpy()[:1 << L1_BITS]
    loc = [_locate(h1, r) for r in ranks]
    h2 = _all_reduce(sp.angle_hist(V, 2, [b for b, _ in loc]), group).cpu().numpy().reshape(4, 1 << L2_BITS)
    ang = []
    for q, (b, rem) in enumerate(loc):
        low, _ = _locate(h2[q], rem)
        ang.append(angle_from_key((b << L2_BITS) | low))
    min_phi, max_phi = lerp_np(ang[0], ang[1], f0), lerp_np(ang[2], ang[3], f1)
    v1 = vec @ np.array([math.cos(min_phi), math.sin(min_phi)])
    v2 = vec @ np.array([math.cos(max_phi), math.sin(max_phi)])
    HE = np.array([v1, v2]) if v1[0] > v2[0] else np.array([v2, v1])
    M = HE / np.linalg.norm(HE, axis=1)[:, None]
    return M, _conc_percentiles(sp, M, n_all, conc_percentile, group)


def _conc_percentiles(sp, M, n_all, conc_percentile, group):
    """Passes 3 + 4: exact ``conc_percentile`` of each concentration over ALL pixels of the slide -> maxC (1x2)."""
    lo, hi, fr = percentile_index(n_all, conc_percentile)
    c1 = _all_reduce(sp.conc_hist(M, 1), group).cpu().numpy().reshape(2, 1 << L1_BITS)
    loc = [_locate(c1[0], lo), _locate(c1[0], hi), _locate(c1[1], lo), _locate(c1[1], hi)]
    c2 = _all_reduce(sp.conc_hist(M, 2, [b for b, _ in loc]), group).cpu().numpy().reshape(4, 1 << L2_BITS)
    cv = []
    for q, (b, rem) in enumerate(loc):
        low, _ = _locate(c2[q], rem)
        cv.append(conc_from_key((b << L2_BITS) | low))
    return np.array([[lerp_np(cv[0], cv[1], fr), lerp_np(cv[2], cv[3], fr)]])


# ------------------------------------------------------------------------------------------------- Vahadane, slide level
RUIFROK_HE = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11]], dtype=np.float64)
DL_SAMPLE_TOL, DL_FULL_TOL = 1e-4, 2e-5          # csrc/sb_pipeline.cu


def _dict_update(D, t):
    """One block-coordinate sweep of Mairal et al. 2010, Alg. 2 with non-negativity and the unit ball (thread 0 of the
    tile kernel).  D: 2x3 (rows = atoms); t = the ten sums of a dictionary pass."""
    A = np.array([[t[0], t[1]], [t[1], t[2]]])
    Bm = np.array([t[3:6], t[6:9]])                      # rows = atoms
    D = D.copy()
    for j in range(2):
        if A[j, j] > 1e-12:
            u = (Bm[j] - A[0, j] * D[0] - A[1, j] * D[1]) / A[j, j] + D[j]
            u = np.maximum(u, 0.0)
            D[j] = u / max(np.linalg.norm(u), 1.0)
    return D


class _Anderson(object):
    """Type-II Anderson acceleration of the 6-component map, as csrc/sb_device.cuh: aa_step."""

    def __init__(self, m):
        self.m, self.dx, self.dr, self.px, self.pr, self.last = m, [], [], None, None, -1.0

    def carry(self):
        self.px, self.pr, self.last = None, None, -1.0

    def step(self, D, FD):
        x, r = D.reshape(-1).copy(), (FD - D).reshape(-1)
        rn = float(np.sqrt((r * r).sum()))
        if self.m <= 0:
            return FD.copy()
        if self.last >= 0.0 and rn > self.last:
            self.dx, self.dr, self.px, self.pr = [], [], None, None
        self.last = rn
        if self.px is not None:
            self.dx.append(x - self.px)
            self.dr.append(r - self.pr)
            if len(self.dx) > self.m:
                self.dx.pop(0)
                self.dr.pop(0)
        self.px, self.pr = x, r.copy()
        h = len(self.dx)
        if h < 1:
            return FD.copy()
        G = np.array([[float((self.dr[i] * self.dr[j]).sum()) for j in range(h)] for i in range(h)])
        tr = float(np.trace(G))
        if not (tr > 0.0 and np.isfinite(tr)):
            return FD.copy()
        try:
            gam = np.linalg.solve(G + 1e-10 * tr * np.eye(h), np.array([float((self.dr[i] * r).sum()) for i in range(h)]))
        except np.linalg.LinAlgError:
            return FD.copy()
        xn = x + r - sum(gam[i] * (self.dx[i] + self.dr[i]) for i in range(h))
        if not np.all(np.isfinite(xn)):
            return FD.copy()
        Dn = np.maximum(xn, 0.0).reshape(D.shape)
        return Dn / np.maximum(np.sqrt((Dn * Dn).sum(axis=1)), 1.0)[:, None]


def vahadane_slide_fit(tiles, luminosity_threshold=0.8, dl_lambda=0.1, dl_iters=10, dl_sample_iters=12, dl_anderson=4,
                       lasso_lambda=0.01, conc_percentile=99.0, group=None, device=None, passes=None):
    """Vahadane stain matrix (2x3) and maxC (1x2) of the union of ``tiles`` over all ranks of ``group``: the dictionary
    iteration of the tile kernel (sample warm start, Anderson acceleration, residual stopping rules) with each pass's
    ten sums all-reduced, the sample being the union of the tiles' own 1-in-16 samples."""
    sp = passes if passes is not None else SlidePasses(tiles, luminosity_threshold, lasso_lambda, device)
    n_px = torch.tensor([sp.n_pixels()], dtype=torch.float64, device=sp.dev)
    n_all = int(round(float(_all_reduce(n_px, group).item())))
    D = RUIFROK_HE / np.linalg.norm(RUIFROK_HE, axis=1)[:, None]
    aa = _Anderson(dl_anderson)
    n_tissue = None
    phases = ([(True, dl_sample_iters)] if dl_sample_iters > 0 else []) + [(False, dl_iters)]
    k = 0
    while k < len(phases):
        sample, n_it = phases[k]
        k += 1
        aa.carry()
        for it in range(n_it):
            t = _reduce_sums(sp.dl_sums(D, dl_lambda, sample), group, FIX_DL)
            if sample and it == 0 and t[9] < 1024.0:              # sample too small: four more full passes instead
                phases[k] = (False, dl_iters + 4)
                aa = _Anderson(dl_anderson)
                break
            if not sample:
                n_tissue = t[9]
                if n_tissue < 1.0:
                    raise TissueMaskException("Empty tissue mask computed")
            FD = _dict_update(D, t)
            stop = dl_anderson > 0 and float(np.sqrt(((FD - D) ** 2).sum())) < (DL_SAMPLE_TOL if sample else DL_FULL_TOL)
            D = aa.step(D, FD)
            if stop:
                break
    if D[0, 0] < D[1, 0]:                                          # vahadane_stain_extractor.py:38-43
        D = D[[1, 0]]
    M = D / np.linalg.norm(D, axis=1)[:, None]
    return M, _conc_percentiles(sp, M, n_all, conc_percentile, group)
