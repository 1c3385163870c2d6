"""A/B of the two Reinhard paths on 1024 x 512^2 tiles: streaming ring passes (sb_reinhard.cu) vs lab_tile_kernel
(SB_REINHARD_TILE_KERNEL=1): bytes must be equal; ms per call by CUDA events.  python tools/reinhard_ab.py"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.synth import synth_batch, synth_tile
from stainlib_b200.utils.stain_utils import LuminosityStandardizer
B, H, W = 1024, 512, 512
pool = torch.from_numpy(synth_batch(5000, 64, H, W))
x = pool.repeat(B // 64, 1, 1, 1).contiguous().cuda()
rein = sb.ReinhardStainNormalizer(); rein.fit(synth_tile(1, H, W, kind="target"))
def timed(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ops = [("transform", lambda: rein.transform(x)), ("transform mask_background", lambda: rein.transform(x, mask_background=True)),
       ("luminosity standardize", lambda: LuminosityStandardizer.standardize(x))]
for name, fn in ops:
    res = {}
    for mode in ("1", "0"):
        os.environ["SB_REINHARD_TILE_KERNEL"] = mode
        out = fn(); ms = timed(fn); res[mode] = (out, ms)
    eq = torch.equal(res["0"][0], res["1"][0])
    nd = int((res["0"][0] != res["1"][0]).sum().item())
    print(f"{name:28s} tile kernel {res['1'][1]:.3f} ms   ring passes {res['0'][1]:.3f} ms   bytes equal {eq} (differing {nd})")
os.environ["SB_REINHARD_TILE_KERNEL"] = "0"
for mode in ("1", "0"):
    os.environ["SB_REINHARD_TILE_KERNEL"] = mode
    r2 = sb.ReinhardStainNormalizer(); r2.fit(x[3])
    print("fit", "tile kernel" if mode == "1" else "ring passes", r2.target_means, r2.target_stds)
