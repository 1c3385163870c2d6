"""One Reinhard transform over 1024 x 512^2 tiles (for ncu captures).  python tools/reinhard_once.py [mask]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.synth import synth_batch, synth_tile
B, H, W = 1024, 512, 512
x = torch.from_numpy(synth_batch(5000, 64, H, W)).repeat(B // 64, 1, 1, 1).contiguous().cuda()
rein = sb.ReinhardStainNormalizer(); rein.fit(synth_tile(1, H, W, kind="target"))
for _ in range(2):
    y = rein.transform(x, mask_background="mask" in sys.argv)
torch.cuda.synchronize()
