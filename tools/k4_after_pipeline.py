"""Diagnostic: duration of K4 (sb_recombine) when it runs right behind the statistics kernel (sb_fit), per iteration,
for a given workload -- CUDA events around each call.  python tools/k4_after_pipeline.py vahadane 512 1024"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200 import _native as nv
from stainlib_b200.synth import synth_tile, synth_batch

method, hw, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
norm = sb.ExtractiveStainNormalizer(method)
norm.fit(synth_tile(1, hw, hw, kind="target"))
pool = torch.from_numpy(synth_batch(5000, min(B, 64), hw, hw))
dev_in = pool.repeat(-(-B // pool.shape[0]), 1, 1, 1)[:B].contiguous().cuda()
out = torch.empty_like(dev_in)
M = torch.empty(B, 2, 3, dtype=torch.float64, device="cuda")
C = torch.empty(B, 2, dtype=torch.float64, device="cuda")
p = norm._params()
h, _ = nv.get_handle(0)
lib = nv.load_library()
Mt = torch.as_tensor(norm.stain_matrix_target, device="cuda").contiguous()
st = nv.stream_ptr(0)
nv.check(lib.sb_fit(h, nv.ptr(dev_in), B, hw, hw, ctypes.byref(p), nv.ptr(M), nv.ptr(C), None, st))
scale = (torch.as_tensor(norm.maxC_target, device="cuda") / C).contiguous()
print("M_src min", float(M.min()), "Mt", norm.stain_matrix_target.round(4).tolist(), "scale range", float(scale.min()), float(scale.max()))
N = 8
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(N)]
torch.cuda.synchronize()
for i in range(N):
    ev[i][0].record()
    nv.check(lib.sb_fit(h, nv.ptr(dev_in), B, hw, hw, ctypes.byref(p), nv.ptr(M), nv.ptr(C), None, st))
    ev[i][1].record()
    nv.check(lib.sb_recombine(h, nv.ptr(dev_in), nv.ptr(out), B, hw, hw, nv.ptr(M), nv.ptr(scale), nv.ptr(Mt), 0.01, st))
    ev[i][2].record()
torch.cuda.synchronize()
print("fit ms     ", [round(e[0].elapsed_time(e[1]), 3) for e in ev])
print("k4 ms after", [round(e[1].elapsed_time(e[2]), 3) for e in ev])
ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
ev2[0].record()
for i in range(N):
    nv.check(lib.sb_recombine(h, nv.ptr(dev_in), nv.ptr(out), B, hw, hw, nv.ptr(M), nv.ptr(scale), nv.ptr(Mt), 0.01, st))
    ev2[i + 1].record()
torch.cuda.synchronize()
print("k4 ms alone", [round(ev2[i].elapsed_time(ev2[i + 1]), 3) for i in range(N)])
