"""Small run of every round-2 kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb                                    # noqa: E402
from stainlib_b200.synth import synth_tile, synth_batch, edge_case_tiles   # noqa: E402
from stainlib_b200.utils.stain_utils import (LuminosityThresholdTissueLocator, convert_RGB_to_OD, convert_OD_to_RGB, lab_split,   # noqa: E402
                                             merge_back, get_mean_std, standardize_brightness, get_concentrations)

tgt = synth_tile(1, 256, kind="target")
e = edge_case_tiles(256, 256)
tiles = torch.from_numpy(np.stack([synth_tile(70, 256), e["all_white"], synth_tile(71, 256), e["dark"]])).cuda()
for method in ("macenko", "vahadane"):
    n = sb.ExtractiveStainNormalizer(method)
    n.fit(tgt)
    out = n.transform(tiles)
    f = sb.ExtractiveStainNormalizer(method, cluster_size=2)
    f.fit(tgt)
    assert torch.equal(out, f.transform(tiles)), method
m = LuminosityThresholdTissueLocator.get_tissue_mask(tiles)
h = sb.HedLightColorAugmenter()
h.transform(tiles)
h.transform(tiles, skimage_version="0.18")
h.transform(tiles.float() / 255.0)
convert_RGB_to_OD(tiles)
P = lab_split(tiles)
merge_back(*P)
get_mean_std(tiles)
standardize_brightness(tiles)
r = sb.ReinhardStainNormalizer()
r.fit(tgt)
r.transform(tiles)                                            # streaming ring passes (sb_reinhard.cu): several tiles per CTA
r.transform(tiles, mask_background=True)
sb.LuminosityStandardizer.standardize(tiles)
big = torch.from_numpy(np.stack([synth_tile(72, 528, 400), synth_tile(73, 528, 400)])).cuda()     # partial last chunk, several CTAs per tile
r.transform(big)
os.environ["SB_REINHARD_TILE_KERNEL"] = "1"
r.transform(tiles)                                            # lab_tile_kernel
os.environ["SB_REINHARD_TILE_KERNEL"] = "0"
M = sb.MacenkoStainExtractor.get_stain_matrix(tiles[[0, 2]])
get_concentrations(tiles[[0, 2]], M)                          # ring pass with coalesced stores
convert_OD_to_RGB(convert_RGB_to_OD(tiles[0]))
torch.cuda.synchronize()
print("sanitize smoke ok", int(m.sum()))
