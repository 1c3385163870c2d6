import os, sys, torch
sys.path.insert(0, '/root/repo')
import stainlib_b200 as sb
from stainlib_b200.synth import synth_batch, synth_tile
B, H, W = 1024, 512, 512
x = torch.from_numpy(synth_batch(5000, 64, H, W)).repeat(B // 64, 1, 1, 1).contiguous().cuda()
n = sb.ExtractiveStainNormalizer("macenko"); n.fit(synth_tile(1, H, W, kind="target"))
for _ in range(2):
    y = n.transform(x)
torch.cuda.synchronize()
