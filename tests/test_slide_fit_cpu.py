"""Host logic of the slide-level (multi-tile) Macenko fit on CPU: the five passes are supplied by a numpy stand-in for
the CUDA kernels (same keys, same histograms), the rank selection / eigenvectors / stain matrix and the all-reduces are
the real code of stainlib_b200/normalization/slide_fit.py.  Checked against the oracle's fit of the concatenated tiles,
in one process and sharded over a world_size-2 gloo group (one rank holding 3 tiles, the other 1)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyPasses(object):
    """numpy restatement of slide_pass_kernel<0..4> (csrc/sb_pipeline.cu) over this rank's tiles."""

    def __init__(self, tiles, thr=0.8, lam=0.01):
        from oracle import stain_oracle as so
        self.so = so
        self.dev = torch.device("cpu")
        self.tiles = tiles
        self.lam = lam
        if tiles is not None and len(tiles):
            flat = np.concatenate([t.reshape(-1, 3) for t in tiles])
            self.od = so.od_lut()[flat].astype(np.float32)
            self.mask = np.concatenate([so.get_tissue_mask(t, thr).reshape(-1) if so_has_tissue(so, t, thr) else np.zeros(t.shape[0] * t.shape[1], bool)
                                        for t in tiles])
        else:
            self.od = np.zeros((0, 3), np.float32)
            self.mask = np.zeros(0, bool)

    def n_pixels(self):
        return int(self.od.shape[0])

    def moments(self):
        x = self.od[self.mask].astype(np.float64)
        s, S = x.sum(0), x.T @ x
        return torch.tensor([s[0], s[1], s[2], S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2], float(len(x))], dtype=torch.float64)

    @staticmethod
    def _angle_key(px, py):
        s = np.abs(px) + np.abs(py)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(s > 0, py / s, 0).astype(np.float32)
        d = np.where(px >= 0, r, np.where(py >= 0, 2 - r, -2 - r)).astype(np.float32)
        t = (d * np.float32(0.25) + np.float32(1.5)).astype(np.float32)
        return np.minimum(t.view(np.uint32) - np.uint32(0x3F800000), (1 << 23) - 1).astype(np.int64)

    @staticmethod
    def _conc_key(c):
        t = (np.float32(2.0) - np.float32(2.0) / (c.astype(np.float32) + np.float32(2.0))).astype(np.float32)
        k = np.minimum(t.view(np.uint32) - np.uint32(0x3F800000), (1 << 23) - 1).astype(np.int64)
        return np.where(c > 0, k, 0)

    @staticmethod
    def _hists(keys_list, level, bins):
        out = np.zeros(8192, np.int64)
        if level == 1:
            for j, k in enumerate(keys_list):
                out[j * 4096:(j + 1) * 4096] += np.bincount(k >> 11, minlength=4096)
        else:
            for q in range(4):
                k = keys_list[q * len(keys_list) // 4]
                sel = k[(k >> 11) == bins[q]]
                out[q * 2048:(q + 1) * 2048] += np.bincount(sel & 2047, minlength=2048)
        return torch.from_numpy(out)

    def angle_hist(self, V, level, bins=None):
        x = self.od[self.mask]
        V = np.asarray(V, np.float32)
        px, py = x @ V[:3], x @ V[3:]
        return self._hists([self._angle_key(px.astype(np.float32), py.astype(np.float32))], level, bins)

    def dl_sums(self, D, lam, sample):
        so = self.so
        od = self.od.astype(np.float64)
        sel = self.mask.copy()
        if sample:
            keep = np.zeros(len(od), bool)
            off = 0
            for t in self.tiles if self.tiles is not None else []:
                n = t.shape[0] * t.shape[1]
                keep[off + so.dl_sample_indices(n)] = True
                off += n
            sel &= keep
        X = od[sel].T
        Al = so.lasso_pos2(X, np.asarray(D, np.float64).reshape(2, 3).T, lam) if X.shape[1] else np.zeros((2, 0))
        A, Bm = Al @ Al.T, X @ Al.T
        return torch.tensor([A[0, 0], A[0, 1], A[1, 1], Bm[0, 0], Bm[1, 0], Bm[2, 0], Bm[0, 1], Bm[1, 1], Bm[2, 1], float(X.shape[1])],
                            dtype=torch.float64)

    def conc_hist(self, M, level, bins=None):
        C = self.so.lasso_pos2(self.od.astype(np.float64).T, np.asarray(M, np.float64).reshape(2, 3).T, self.lam) if len(self.od) else np.zeros((2, 0))
        return self._hists([self._conc_key(C[0]), self._conc_key(C[1])], level, bins)


def so_has_tissue(so, t, thr):
    try:
        so.get_tissue_mask(t, thr)
        return True
    except so.TissueMaskException:
        return False


def _tiles():
    from stainlib_b200.synth import synth_tile
    return [synth_tile(40 + i, 96, 112) for i in range(3)] + [np.full((96, 112, 3), 255, np.uint8)]


def _big_tiles():
    from stainlib_b200.synth import synth_tile
    return [synth_tile(60 + i, 160, 176) for i in range(4)]


def _expected(tiles):
    from oracle import stain_oracle as so
    big = np.concatenate(tiles, axis=0)
    n = so.ExtractiveStainNormalizer("macenko")
    n.fit(big)
    return n.stain_matrix_target, n.maxC_target


def test_slide_fit_host_logic_single_process():
    sys.path.insert(0, ROOT)
    from stainlib_b200.normalization.slide_fit import macenko_slide_fit
    tiles = _tiles()
    M, maxC = macenko_slide_fit(None, passes=NumpyPasses(tiles))
    M_ref, C_ref = _expected(tiles)
    np.testing.assert_allclose(M, M_ref, rtol=0, atol=1e-5)
    np.testing.assert_allclose(maxC, C_ref, rtol=1e-5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stainlib_b200.normalization.slide_fit import macenko_slide_fit, vahadane_slide_fit
    tiles = _tiles()
    mine = tiles[:3] if rank == 0 else tiles[3:]          # rank 1 holds only the all-white tile: no tissue of its own
    M, maxC = macenko_slide_fit(None, passes=NumpyPasses(mine))
    big = _big_tiles()
    Mv, Cv = vahadane_slide_fit(None, passes=NumpyPasses(big[rank::world]))        # interleaved shards
    q.put((rank, M.tolist(), maxC.tolist(), Mv.tolist(), Cv.tolist()))
    dist.destroy_process_group()


def test_slide_fit_sharded_world2_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]             # every rank holds the same statistics
    from stainlib_b200.normalization.slide_fit import macenko_slide_fit
    M1, C1 = macenko_slide_fit(None, passes=NumpyPasses(_tiles()))
    np.testing.assert_allclose(np.array(res[0][1]), M1, rtol=0, atol=1e-12)   # sharded == unsharded (integer histograms add exactly)
    np.testing.assert_allclose(np.array(res[0][2]), C1, rtol=1e-12)
    # Vahadane: every dictionary pass all-reduces ten sums; the ranks agree and match the unsharded run (summation order only)
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4]
    from stainlib_b200.normalization.slide_fit import vahadane_slide_fit
    Mv1, Cv1 = vahadane_slide_fit(None, passes=NumpyPasses(_big_tiles()))
    np.testing.assert_allclose(np.array(res[0][3]), Mv1, rtol=0, atol=1e-8)
    np.testing.assert_allclose(np.array(res[0][4]), Cv1, rtol=1e-7)


def test_key_inverses_roundtrip():
    from stainlib_b200.normalization.slide_fit import angle_from_key, conc_from_key
    ang = np.linspace(-3.1, 3.1, 41)
    keys = NumpyPasses._angle_key(np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32))
    back = np.array([angle_from_key(int(k)) for k in keys])
    assert np.abs(back - ang).max() < 2e-6
    c = np.array([0.0, 1e-3, 0.5, 1.5, 4.0, 9.0])
    back = np.array([conc_from_key(int(k)) for k in NumpyPasses._conc_key(c)])
    assert np.abs(back - c).max() < 2e-5


def oracle_vahadane_slide(tiles):
    """What the reference-style learner returns for the union of the tiles: the oracle's accelerated schedule on all
    tissue pixels, warm-started on the union of the tiles' own samples; then maxC on the concatenated image."""
    from oracle import stain_oracle as so
    Xs, X = [], []
    for t in tiles:
        od = so.convert_RGB_to_OD(t).reshape(-1, 3)
        m = so_has_tissue(so, t, 0.8) and so.get_tissue_mask(t).reshape(-1)
        if m is False:
            continue
        X.append(od[m])
        si = so.dl_sample_indices(len(od))
        Xs.append(od[si[m[si]]])
    D = so.train_dl_accel(np.concatenate(X).T, np.concatenate(Xs).T)
    M = so.vahadane_finish(D.T)
    big = np.concatenate(list(tiles), axis=0)
    C = so.get_concentrations(big, M)
    return M, np.percentile(C, 99, axis=0).reshape(1, 2)


def test_vahadane_slide_fit_host_logic():
    sys.path.insert(0, ROOT)
    from stainlib_b200.synth import synth_tile
    from stainlib_b200.normalization.slide_fit import vahadane_slide_fit
    tiles = [synth_tile(40 + i, 160, 176) for i in range(3)] + [np.full((160, 176, 3), 255, np.uint8)]
    M, maxC = vahadane_slide_fit(None, passes=NumpyPasses(tiles))
    M_ref, C_ref = oracle_vahadane_slide(tiles)
    np.testing.assert_allclose(M, M_ref, rtol=0, atol=1e-5)
    np.testing.assert_allclose(maxC, C_ref, rtol=1e-4)
