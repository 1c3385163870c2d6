"""Device-resident timing of every single-operator entry point on 1024 x 512^2 tiles (CUDA events, 3 warm-ups, mean of
10): ms per launch sequence, Gpx/s, and the fraction of the measured HBM copy peak at the operator's algorithmic bytes.
python tools/pointwise_bench.py > gpurun_out/pointwise.txt"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.augmentation.augmenter import GrayscaleAugmentor, HedLightColorAugmenter, StainAugmentor
from stainlib_b200.synth import synth_batch, synth_tile
from stainlib_b200.utils.stain_utils import (LuminosityStandardizer, LuminosityThresholdTissueLocator, get_concentrations, get_mean_std,
                                             standardize_brightness, lab_split, merge_back, convert_RGB_to_OD, convert_OD_to_RGB)
from stainlib_b200 import _native as nv

B, H, W = 1024, 512, 512
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
pool = torch.from_numpy(synth_batch(5000, 64, H, W))
x = pool.repeat(B // 64, 1, 1, 1).contiguous().cuda()
npx = B * H * W


def timed(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


rng = np.random.default_rng(0)
hed = HedLightColorAugmenter()
sig, bia = rng.uniform(-0.1, 0.1, (B, 3)), rng.uniform(-0.1, 0.1, (B, 3))
rein = sb.ReinhardStainNormalizer()
rein.fit(synth_tile(1, H, W, kind="target"))
mac = sb.ExtractiveStainNormalizer("macenko")
mac.fit(synth_tile(1, H, W, kind="target"))
aug = StainAugmentor("macenko")
aug.fit(x)
Msrc = sb.MacenkoStainExtractor.get_stain_matrix(x)
planes = lab_split(x)
od32 = convert_RGB_to_OD(x, dtype=torch.float32)
_m, _s = torch.empty(B, 3, dtype=torch.float64, device="cuda"), torch.empty(B, 3, dtype=torch.float64, device="cuda")


def rein_stats(t):
    h, idx = nv.get_handle(0)
    nv.check(nv.load_library().sb_reinhard_stats(h, nv.ptr(t), B, H, W, nv.ptr(_m), nv.ptr(_s), nv.stream_ptr(idx)))


gray = GrayscaleAugmentor()
gray.fit(x)
rows = [
    ("HedLightColorAugmenter.transform (hed ring)", 6.0, lambda: hed.transform(x, sigmas=sig, biases=bia)),
    ("StainAugmentor.pop (stain-augment ring)", 6.0, lambda: aug.pop()),
    ("GrayscaleAugmentor.pop (gray ring)", 6.0, lambda: gray.pop()),
    ("ReinhardStainNormalizer.transform (3 ring passes)", 6.0, lambda: rein.transform(x)),
    ("ReinhardStainNormalizer.transform mask_background (3 ring passes)", 6.0, lambda: rein.transform(x, mask_background=True)),
    ("LuminosityStandardizer.standardize (2 ring passes)", 6.0, lambda: LuminosityStandardizer.standardize(x)),
    ("get_concentrations (fp32 [B,N,2] out)", 11.0, lambda: get_concentrations(x, Msrc)),
    ("get_mean_std (byte-free: forward LAB + statistics pass)", 3.0, lambda: get_mean_std(x)),
    ("ReinhardStainNormalizer.fit statistics over the batch (sb_reinhard_stats: 2 ring passes)", 3.0, lambda: rein_stats(x)),
    ("standardize_brightness (lab_tile_kernel)", 6.0, lambda: standardize_brightness(x)),
    ("lab_split (3 float planes out)", 15.0, lambda: lab_split(x)),
    ("merge_back (3 float planes in)", 15.0, lambda: merge_back(*planes)),
    ("convert_RGB_to_OD float32", 15.0, lambda: convert_RGB_to_OD(x, dtype=torch.float32)),
    ("convert_RGB_to_OD float64", 27.0, lambda: convert_RGB_to_OD(x[:512])),
    ("convert_OD_to_RGB float32", 15.0, lambda: convert_OD_to_RGB(od32)),
    ("get_tissue_mask (mask ring pass)", 4.0, lambda: LuminosityThresholdTissueLocator.get_tissue_mask(x)),
    ("MacenkoStainExtractor.get_stain_matrix (streaming passes 1-4)", 3.0, lambda: sb.MacenkoStainExtractor.get_stain_matrix(x)),
    ("ExtractiveStainNormalizer('macenko').transform", 6.0, lambda: mac.transform(x)),
]
print(f"# {B} x {H}x{W} tiles, device-resident, ms per call; peak = {peak:.0f} GB/s (measured copy rate)")
for name, bpp, fn in rows:
    try:
        ms = timed(fn)
        n_px = npx // 2 if "float64" in name else npx
        print(f"{name:90s} {ms:8.3f} ms  {n_px / ms / 1e6:8.1f} Gpx/s  {n_px * bpp / ms / 1e6 / peak:6.3f} of peak @ {bpp:.0f} B/px")
    except Exception as e:                                     # keep going: this is a survey
        print(f"{name:90s} failed: {type(e).__name__}: {e}")
