"""config[0]: one 256x256 tile, numpy in -> numpy out through the reference-facing API (fit once, transform repeatedly):
wall-clock latency per call, beside a 512^2 and a 1024^2 tile.  python tools/single_tile_latency.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stainlib_b200 as sb
from stainlib_b200.synth import synth_tile
for method in ("macenko", "vahadane"):
    for H in (256, 512, 1024):
        src, tgt = synth_tile(0, H), synth_tile(1, H, kind="target")
        n = sb.ExtractiveStainNormalizer(method)
        t0 = time.perf_counter(); n.fit(tgt); torch.cuda.synchronize(); tf = time.perf_counter() - t0
        for _ in range(5): n.transform(src)
        t0 = time.perf_counter()
        for _ in range(50): out = n.transform(src)
        dt = (time.perf_counter() - t0) / 50
        x = torch.from_numpy(src).cuda()[None]
        for _ in range(5): n.transform(x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(50): y = n.transform(x)
        torch.cuda.synchronize(); dd = (time.perf_counter() - t0) / 50
        print(f"{method:9s} {H:4d}^2: numpy->numpy {dt * 1e3:7.3f} ms/tile ({H * H / dt / 1e6:7.1f} Mpx/s)   device tensor {dd * 1e3:7.3f} ms   first fit {tf * 1e3:.1f} ms")
r = sb.ReinhardStainNormalizer(); r.fit(synth_tile(1, 256, kind="target"))
src = synth_tile(0, 256)
for _ in range(5): r.transform(src)
t0 = time.perf_counter()
for _ in range(50): r.transform(src)
print(f"reinhard   256^2: numpy->numpy {(time.perf_counter() - t0) / 50 * 1e3:7.3f} ms/tile")
