"""Seeded synthetic H&E tile generator.

Used by the parity tests, the golden-fixture generator and ``bench.py`` so that the CPU oracle and the CUDA path always
see identical inputs.  Pure numpy + cv2 (bicubic resize); no dependency on the oracle or on the CUDA library.

Planted stains are the Ruifrok H/E vectors; ``kind='target'`` uses a second stain pair so that normalisation has
something to do.  About three quarters of each tile is tissue, the rest is white background.
"""
import numpy as np
import cv2

STAINS_SOURCE = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11]], dtype=np.float64)
STAINS_TARGET = np.array([[0.55, 0.76, 0.33], [0.10, 0.95, 0.28]], dtype=np.float64)


def _unit_rows(M):
    return M / np.linalg.norm(M, axis=1)[:, None]


def _field(rng, H, W, s):
    coarse = rng.random((H // s + 2, W // s + 2)).astype(np.float32)
    return cv2.resize(coarse, (W, H), interpolation=cv2.INTER_CUBIC).astype(np.float64)


def synth_tile(seed, H=256, W=None, kind="source", noise=2.0):
    """One uint8 [H,W,3] RGB tile, deterministic in ``seed``."""
    W = H if W is None else W
    rng = np.random.default_rng(seed)
    M = _unit_rows(STAINS_TARGET if kind == "target" else STAINS_SOURCE)
    cH = np.clip(2.0 * _field(rng, H, W, 16) - 0.5, 0, None)
    cE = np.clip(1.5 * _field(rng, H, W, 32) - 0.2, 0, None)
    bg = _field(rng, H, W, 64) > 0.75
    cH[bg] = 0.0
    cE[bg] = 0.0
    C = np.stack([cH, cE], axis=-1).reshape(-1, 2)
    rgb = 255.0 * np.exp(-C @ M) + rng.normal(0.0, noise, size=(H * W, 3))
    return np.clip(rgb, 0, 255).astype(np.uint8).reshape(H, W, 3)


def synth_batch(base_seed, B, H=256, W=None, kind="source", pool=64):
    """uint8 [B,H,W,3]; tile i is ``synth_tile(base_seed + i % pool)`` (a pool of distinct tiles, replicated)."""
    W = H if W is None else W
    n = min(B, pool)
    tiles = np.stack([synth_tile(base_seed + i, H, W, kind) for i in range(n)])
    if B <= n:
        return tiles
    reps = -(-B // n)
    return np.concatenate([tiles] * reps, axis=0)[:B]


def edge_case_tiles(H=64, W=64):
    """Degenerate tiles the reference handles specially. Returns dict name -> uint8 [H,W,3]."""
    out = {}
    out["all_white"] = np.full((H, W, 3), 255, np.uint8)
    one = np.full((H, W, 3), 255, np.uint8)
    one[H // 2, W // 2] = (120, 60, 140)
    out["one_tissue_pixel"] = one
    sat = synth_tile(7, H, W)
    sat[: H // 4] = 0
    sat[-(H // 4):] = 255
    out["saturated_bands"] = sat
    rng = np.random.default_rng(11)
    c = np.clip(1.8 * _field(rng, H, W, 8), 0, None).reshape(-1, 1)
    m = _unit_rows(STAINS_SOURCE)[:1]
    single = 255.0 * np.exp(-c @ m) + rng.normal(0.0, 1.0, size=(H * W, 3))
    out["near_single_stain"] = np.clip(single, 0, 255).astype(np.uint8).reshape(H, W, 3)
    out["dark"] = (synth_tile(13, H, W).astype(np.float64) * 0.35).astype(np.uint8)
    return out
