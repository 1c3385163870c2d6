import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import stainlib_b200 as sb
from stainlib_b200 import _native as nv
from stainlib_b200.synth import synth_tile, synth_batch
for H,B in ((256,64),(512,64),(1024,32),(2048,4)):
    n = sb.ExtractiveStainNormalizer('macenko'); n.fit(synth_tile(1,H,kind='target'))
    x = torch.from_numpy(synth_batch(5000, B, H)).cuda()
    nv.stream_fallbacks(reset=True)
    n.transform(x); torch.cuda.synchronize()
    print('macenko', H, B, nv.stream_fallbacks())
    v = sb.ExtractiveStainNormalizer('vahadane'); v.fit(synth_tile(1,H,kind='target'))
    nv.stream_fallbacks(reset=True)
    v.transform(x); torch.cuda.synchronize()
    print('vahadane', H, B, nv.stream_fallbacks())
