import os, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import stainlib_b200 as sb
from stainlib_b200.synth import synth_batch
B,H,W=1024,512,512
x = torch.from_numpy(synth_batch(5000, 64, H, W)).repeat(B // 64, 1, 1, 1).contiguous().cuda()
for method in ("macenko","vahadane"):
    n = sb.ExtractiveStainNormalizer(method)
    n.fit(x, slide=True); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(3): n.fit(x, slide=True)
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/3
    print(method, "slide-level fit over", B, "tiles:", round(dt*1e3,2), "ms", round(B*H*W/dt/1e9,1), "Gpx/s")
